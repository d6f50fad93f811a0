"""jax.scipy -> SciPy.  Test infrastructure."""
import numpy as _np
import scipy.linalg as _sl
import scipy.special as _sp
from ..numpy import _wrap


class _Linalg:
    @staticmethod
    def lu_factor(a, overwrite_a=False, check_finite=True):
        lu, piv = _sl.lu_factor(_np.array(a), check_finite=False)
        return _wrap(lu), piv

    @staticmethod
    def lu_solve(lu_and_piv, b, trans=0, overwrite_b=False, check_finite=True):
        lu, piv = lu_and_piv
        return _wrap(_sl.lu_solve((_np.asarray(lu), piv), _np.asarray(b), trans=trans, check_finite=False))

    solve = staticmethod(lambda a, b: _wrap(_sl.solve(a, b)))


class _Special:
    gamma = staticmethod(lambda x: _wrap(_np.asarray(_sp.gamma(x))))
    factorial = staticmethod(lambda x: _wrap(_np.asarray(_sp.factorial(x))))
    gammaln = staticmethod(lambda x: _wrap(_np.asarray(_sp.gammaln(x))))


class _Integrate:
    trapezoid = staticmethod(lambda y, x=None, dx=1.0, axis=-1: _wrap(_np.asarray(_np.trapezoid(y, x=x, dx=dx, axis=axis))))


linalg = _Linalg()
special = _Special()
integrate = _Integrate()
