"""Drop-in for diff_render/diftet_6_subdiv/3_model/cameraop.py."""
from deftet_b200.diffrender import perspective  # noqa: F401
