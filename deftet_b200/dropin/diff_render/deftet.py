"""Drop-in for diff_render/diftet_6_subdiv/3_model/deftet.py: `from deftet import Deftet` (6_optim/optim_with_mask_subdiv_from_gridmov.py:30)
resolves here when deftet_b200/dropin/diff_render precedes the reference's 3_model directory on sys.path."""
from deftet_b200.diffrender import Deftet  # noqa: F401
