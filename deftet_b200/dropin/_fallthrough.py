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
