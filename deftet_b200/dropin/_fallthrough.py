"""Let a drop-in package fall through to the reference checkout for the sub-modules it does not replace.

``layers`` and ``utils`` of this tree shadow the reference's packages of the same name (put
``deftet_b200/dropin`` first on ``sys.path``).  When ``DEFTET_REFERENCE_ROOT`` points at a DefTet checkout,
its ``layers/`` / ``utils/`` directories are appended to the package ``__path__`` so that e.g.
``layers.pc_model`` or ``utils.experiment`` still import from there, unmodified."""
import os


def extend(pkg_path, name):
    root = os.environ.get("DEFTET_REFERENCE_ROOT")
    if root:
        cand = os.path.join(root, name)
        if os.path.isdir(cand) and cand not in pkg_path:
            pkg_path.append(cand)


def exec_reference_init(namespace, name):
    """Run the reference package's own ``__init__.py`` inside the shadowing package (its imports resolve through the extended
    ``__path__``, this tree first).  Returns False when no checkout is configured."""
    root = os.environ.get("DEFTET_REFERENCE_ROOT")
    if not root:
        return False
    init = os.path.join(root, name, "__init__.py")
    if not os.path.isfile(init):
        return False
    with open(init) as f:
        exec(compile(f.read(), init, "exec"), namespace)
    return True


def adopt_reference_module(namespace, rel_path, overrides):
    """Make a shadowing module complete: load the reference file ``rel_path`` (e.g. 'utils/mesh_utils.py') from the checkout
    under a private module name, REPLACE in it the functions this tree re-implements (``overrides``: names defined in
    ``namespace``), and copy every other public name into ``namespace`` -- so that ``from utils.mesh_utils import save_mesh``
    (reference dataloader.py:15) or ``mesh_utils.save_tet_face`` (eval.py) keep working, and reference helpers that call a
    replaced function by its global name (e.g. ``get_tet_adj`` -> ``tet_adj_share``) run the GPU version.
    Returns the private module, or None when no checkout is configured (then only the replaced names exist)."""
    import importlib.util
    import sys
    root = os.environ.get("DEFTET_REFERENCE_ROOT")
    if not root:
        return None
    path = os.path.join(root, rel_path)
    if not os.path.isfile(path):
        return None
    name = "_deftet_reference_." + rel_path[:-3].replace("/", ".")
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        try:
            spec.loader.exec_module(mod)
        except Exception:
            del sys.modules[name]
            raise
    for k in overrides:
        if k in namespace:
            setattr(mod, k, namespace[k])
    for k, v in vars(mod).items():
        if not k.startswith("__") and k not in namespace:
            namespace[k] = v
    return mod


def missing_attribute(module_name, name):
    return AttributeError("module %r of the deftet_b200 drop-in does not replace %r; set DEFTET_REFERENCE_ROOT to a DefTet "
                          "checkout so that the reference definition is used" % (module_name, name))
