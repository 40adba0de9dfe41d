"""Locate the reference's package of the same name further down sys.path and chain to it."""
import os
import sys


def reference_package_dir(name: str, own_dir: str):
    own = os.path.realpath(own_dir)
    for entry in sys.path:
        cand = os.path.join(entry or ".", name)
        if os.path.isdir(cand) and os.path.realpath(cand) != own and os.path.exists(os.path.join(cand, "__init__.py")):
            return cand
    return None


def chain(module_globals: dict, name: str) -> None:
    """Extend `__path__` with the reference package and run its `__init__.py` in this namespace: relative imports
    inside it resolve shadowed submodules here first, everything else in the reference."""
    own_dir = os.path.dirname(os.path.abspath(module_globals["__file__"]))
    ref = reference_package_dir(name, own_dir)
    if ref is None:
        return
    module_globals["__path__"].append(ref)
    init = os.path.join(ref, "__init__.py")
    with open(init) as f:
        src = f.read()
    exec(compile(src, init, "exec"), module_globals)
