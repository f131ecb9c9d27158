"""Drop-in for reference layers/DefTet/tet_analytic_distance_batch/utils.py."""
from deftet_b200.surface import _AnalyticDistance as VarianceFunc
from deftet_b200.surface import tet_analytic_distance_f_batch  # noqa: F401
