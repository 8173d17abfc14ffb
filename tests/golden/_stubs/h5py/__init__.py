"""Inert stand-in for h5py: only import-time names, nothing on the geometric/label path calls it."""


class _Dummy:
    def __init__(self, *a, **k):
        raise RuntimeError("h5py stub: not available in this container")


class File(_Dummy):
    pass


class Group(_Dummy):
    pass


class Dataset(_Dummy):
    pass


class HLObject(_Dummy):
    pass


def special_dtype(**kwargs):
    return object


def string_dtype(*a, **k):
    return object
