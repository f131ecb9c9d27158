"""Drop-in for diff_render/diftet_6_subdiv/5_rendereq/deftetrneder.py (`from deftetrneder import rendermeshcolor, preprocess_save`,
6_optim/optim_with_mask_subdiv_from_gridmov.py:34): the rasterizer + peel2mask run as one fused kernel pair, no Kaolin needed."""
from deftet_b200.diffrender import peel2mask, preprocess_save, rendermeshcolor  # noqa: F401
