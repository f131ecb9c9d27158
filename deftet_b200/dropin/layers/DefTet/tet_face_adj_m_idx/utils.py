"""Drop-in for reference layers/DefTet/tet_face_adj_m_idx/utils.py."""
from deftet_b200.surface import tet_face_adj_m_f_idx  # noqa: F401
