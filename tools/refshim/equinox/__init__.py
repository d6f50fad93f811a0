"""equinox stand-in: Module = a dataclass.  Test infrastructure."""
import dataclasses
from . import internal  # noqa: F401


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        dataclasses.dataclass(cls, eq=False)


def filter_jit(f=None, **kw):
    return f if f is not None else (lambda g: g)
