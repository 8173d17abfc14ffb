"""Inert stand-in for matplotlib (imported by reference train.py / tests, never drawn with here)."""
import sys, types

def use(*a, **k):
    pass

for _n in ("pyplot", "patches", "collections", "lines", "gridspec", "figure", "axes", "cm", "colors", "animation", "backends"):
    _m = types.ModuleType(f"matplotlib.{_n}")
    _m.__getattr__ = lambda name: type(name, (), {})
    sys.modules[f"matplotlib.{_n}"] = _m
    globals()[_n] = _m
