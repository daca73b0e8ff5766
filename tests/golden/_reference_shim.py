"""Import the unmodified reference (``/root/reference``) in a container that lacks its
GUI / HDF5 dependencies by fabricating inert stand-ins for those packages only.

Used by ``make_golden.py`` (build container only).  Never imported by tests that run
on the GPU box, where /root/reference does not exist.
"""
import importlib.abc
import importlib.machinery
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"
_ABSENT = ("matplotlib", "mpl_toolkits", "h5py", "ipywidgets", "IPython")


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        return _Inert()

    def __iter__(self):
        return iter(())


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _ABSENT:
            try:
                sys.meta_path.remove(self)
                real = importlib.util.find_spec(name)
            except Exception:
                real = None
            finally:
                sys.meta_path.insert(0, self)
            if real is None:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = types.ModuleType(spec.name)
        mod.__path__ = []
        mod.__getattr__ = lambda key: _Inert()
        return mod

    def exec_module(self, module):
        pass


def import_reference():
    import importlib.util  # noqa: F401

    sys.meta_path.insert(0, _Finder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import hmclab  # noqa: F401
        import hmclab.Distributions  # noqa: F401
        import hmclab.MassMatrices  # noqa: F401
        import hmclab.Samplers  # noqa: F401
    return hmclab
