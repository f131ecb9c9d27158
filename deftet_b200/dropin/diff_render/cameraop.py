"""Drop-in for diff_render/diftet_6_subdiv/3_model/cameraop.py."""
from deftet_b200.diffrender import perspective  # noqa: F401


# every other name of the reference module comes from the checkout at DEFTET_REFERENCE_ROOT (see dropin/_fallthrough.py)
from _fallthrough import adopt_reference_module as _adopt, missing_attribute as _missing  # noqa: E402

_REPLACED = ('perspective',)
_reference = _adopt(globals(), 'diff_render/diftet_6_subdiv/3_model/cameraop.py', _REPLACED)


def __getattr__(name):
    raise _missing(__name__, name)
