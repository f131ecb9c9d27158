"""Leaf-only variant of the diff_render drop-in (python -m deftet_b200.run --leaf ...): the reference's own Deftet model and
rendermeshcolor stay in charge and only this leaf module is replaced -- same code as ../diff_render/cameraop.py."""
import os as _os

_path = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "diff_render", "cameraop.py")
with open(_path) as _f:
    exec(compile(_f.read(), _path, "exec"))
