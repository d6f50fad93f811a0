"""jax.numpy stand-in on NumPy float64 (see ../../README.md).  Test infrastructure."""
import numpy as _np

pi = _np.pi
inf = _np.inf
nan = _np.nan
float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
complex128 = _np.complex128
bool_ = _np.bool_
newaxis = None
dtype = _np.dtype
finfo = _np.finfo
result_type = _np.result_type
issubdtype = _np.issubdtype
floating = _np.floating
integer = _np.integer


class _AtIndexer:
    __slots__ = ("arr", "idx")

    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _new(self, dtype_of):
        out = _np.array(self.arr, dtype=_np.result_type(self.arr.dtype, _np.asarray(dtype_of).dtype), copy=True).view(ndarray)
        return out

    def set(self, v):
        out = self._new(v); out[self.idx] = v; return out

    def add(self, v):
        out = self._new(v); _np.add.at(out, self.idx, v); return out

    def multiply(self, v):
        out = self._new(v); _np.multiply.at(out, self.idx, v); return out

    def divide(self, v):
        out = self._new(v); out[self.idx] = out[self.idx] / v; return out

    def get(self):
        return self.arr[self.idx]


class _At:
    __slots__ = ("arr",)

    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        if isinstance(idx, ndarray):
            idx = _np.asarray(idx)
        return _AtIndexer(self.arr, idx)


class ndarray(_np.ndarray):
    """np.ndarray + the functional-update property `.at` of jax arrays."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


Array = ndarray


def _wrap(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, ndarray):
        return x.view(ndarray)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _lift(fn):
    def f(*a, **k):
        return _wrap(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "f")
    return f


def array(x, dtype=None, copy=True):
    return _np.array(x, dtype=dtype).view(ndarray)


def asarray(x, dtype=None):
    return _np.asarray(x, dtype=dtype).view(ndarray)


def _re(x):
    x = _np.asarray(x)
    return x.real if _np.iscomplexobj(x) else x


def searchsorted(a, v, side="left"):
    # decisions follow real parts (complex-step differentiation passes complex arguments)
    return _wrap(_np.asarray(_np.searchsorted(_re(a), _re(v), side=side)))


def where(c, a=None, b=None):
    if a is None:
        return _wrap(_np.where(c))
    return _wrap(_np.where(c, a, b))


def clip(x, a_min=None, a_max=None, min=None, max=None):
    lo = a_min if a_min is not None else min
    hi = a_max if a_max is not None else max
    return _wrap(_np.clip(x, lo, hi))


def maximum(a, b):
    if _np.iscomplexobj(a) or _np.iscomplexobj(b):
        return where(_re(a) >= _re(b), a, b)
    return _wrap(_np.maximum(a, b))


def minimum(a, b):
    if _np.iscomplexobj(a) or _np.iscomplexobj(b):
        return where(_re(a) <= _re(b), a, b)
    return _wrap(_np.minimum(a, b))


def abs(x):
    if _np.iscomplexobj(x):
        return _wrap(_np.asarray(x) * _np.sign(_re(x)))
    return _wrap(_np.abs(x))


absolute = abs


def sign(x):
    return _wrap(_np.sign(_re(x)))


def trapz(y, x=None, dx=1.0, axis=-1):
    return _wrap(_np.trapezoid(y, x=x, dx=dx, axis=axis))


trapezoid = trapz


def atleast_1d(*a):
    r = _np.atleast_1d(*a)
    return _wrap(r)


def append(a, v, axis=None):
    return _wrap(_np.append(a, v, axis=axis))


class _Linalg:
    eigh = staticmethod(_lift(_np.linalg.eigh))
    lstsq = staticmethod(lambda a, b, rcond=None: _wrap(_np.linalg.lstsq(a, b, rcond=rcond)))
    solve = staticmethod(_lift(_np.linalg.solve))
    inv = staticmethod(_lift(_np.linalg.inv))
    norm = staticmethod(_lift(_np.linalg.norm))


linalg = _Linalg()


class _FFT:
    rfft = staticmethod(_lift(_np.fft.rfft))
    irfft = staticmethod(_lift(_np.fft.irfft))
    fft = staticmethod(_lift(_np.fft.fft))
    ifft = staticmethod(_lift(_np.fft.ifft))


fft = _FFT()


def __getattr__(name):
    fn = getattr(_np, name)
    if callable(fn) and not isinstance(fn, type):
        return _lift(fn)
    return fn
