"""Meta-path finder fabricating any `kornia.*` module with dummy classes.

kornia is not installed (and there is no network); the reference imports kornia names at module import time
(batch/intensity.py:9-27, tensors/image_geometric_torch.py:8, neuralnets/modelcomponents.py:11-12) and subclasses
two of them (intensity.py:43,56), so every attribute must be a class.  None of it is executed on the
geometric/label half that the golden vectors pin.
"""
import importlib.abc
import importlib.machinery
import sys
import types


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name == "kornia" or name.startswith("kornia."):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []

        def _getattr(attr, _modname=spec.name):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return type(attr, (), {"__init__": lambda self, *a, **k: None})

        m.__getattr__ = _getattr
        return m

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
