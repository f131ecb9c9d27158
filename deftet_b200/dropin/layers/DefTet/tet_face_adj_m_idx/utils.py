"""Drop-in for reference layers/DefTet/tet_face_adj_m_idx/utils.py."""
from deftet_b200.surface import face_adjacency_table, tet_face_adj_m_f_idx  # noqa: F401


class _Ext:
    """Stands in for the pybind module ``tet_face_adj_m_idx`` (tet_face_adj_m.cpp:26-34): fills adj_idx (F,30) in place."""

    @staticmethod
    def forward(face_fx3x3, adj_idx):
        if face_fx3x3.shape[0]:
            adj_idx.copy_(face_adjacency_table(face_fx3x3=face_fx3x3))


tet_face_adj_m_idx = _Ext()
