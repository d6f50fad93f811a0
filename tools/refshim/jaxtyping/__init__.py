"""jaxtyping stand-in: every annotation is Any.  Test infrastructure."""
from typing import Any


class _Ann:
    def __getitem__(self, item):
        return Any


Array = Any
ArrayLike = Any
PyTree = _Ann()
Scalar = Any
Bool = Float = Int = Shaped = Real = Num = Complex = Inexact = _Ann()
