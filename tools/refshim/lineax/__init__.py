"""lineax stand-in: the Jacobian operator GRKT4.step builds (ode_integrators_stiff.py), as a dense complex-step
Jacobian.  Test infrastructure."""
import numpy as _np
import jax as _jax
from jax.numpy import _wrap


class AbstractLinearSolver:
    pass


class LU(AbstractLinearSolver):
    pass


class JacobianLinearOperator:
    def __init__(self, fn, x, args=None, tags=()):
        self.fn, self.x, self.args = fn, x, args
        self._mat = None

    def as_matrix(self):
        if self._mat is None:
            self._mat = _np.asarray(_jax.jacfwd(lambda y: self.fn(y, self.args))(self.x))
        return _wrap(self._mat)

    def mv(self, v):
        return _wrap(_np.asarray(self.as_matrix()) @ _np.asarray(v))


def linearise(op):
    op.as_matrix()
    return op
