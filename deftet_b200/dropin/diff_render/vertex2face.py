"""Drop-in for diff_render/diftet_6_subdiv/4_render/vertex2face.py (per-vertex -> per-face gather)."""


def vertex2face(vertex_features_bxpxk, faces_fx3):
    b, k = vertex_features_bxpxk.shape[0], vertex_features_bxpxk.shape[2]
    return vertex_features_bxpxk[:, faces_fx3.view(-1)].view(b, -1, k * 3)


# every other name of the reference module comes from the checkout at DEFTET_REFERENCE_ROOT (see dropin/_fallthrough.py)
from _fallthrough import adopt_reference_module as _adopt, missing_attribute as _missing  # noqa: E402

_REPLACED = ('vertex2face',)
_reference = _adopt(globals(), 'diff_render/diftet_6_subdiv/4_render/vertex2face.py', _REPLACED)


def __getattr__(name):
    raise _missing(__name__, name)
