"""Oracle package: TEST INFRASTRUCTURE ONLY (see hipr_oracle.py header)."""
import importlib.util
import os
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_CACHE = {}


def load_ref(name):
    """Import the compiled, unmodified reference module `neighbor2d` or `neighbor` from
    oracle/_ref/ under a private module name (so it never shadows the product's drop-in
    modules of the same name).  Returns None when oracle/_ref has not been built."""
    path = os.path.join(REF_DIR, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    # the extension's PyInit symbol is PyInit_<name>, so the spec name must stay <name>
    if name in _CACHE:
        return _CACHE[name]
    import sys
    saved = sys.modules.get(name)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # Cython's module init registers itself in sys.modules under its own name; undo that so
    # `import neighbor2d` keeps resolving to the product's drop-in module.
    if saved is not None:
        sys.modules[name] = saved
    else:
        sys.modules.pop(name, None)
    _CACHE[name] = mod
    return mod
