"""Drop-in for diff_render/diftet_6_subdiv/4_render/vertex2face.py (per-vertex -> per-face gather)."""


def vertex2face(vertex_features_bxpxk, faces_fx3):
    b, k = vertex_features_bxpxk.shape[0], vertex_features_bxpxk.shape[2]
    return vertex_features_bxpxk[:, faces_fx3.view(-1)].view(b, -1, k * 3)
