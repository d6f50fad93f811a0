"""lineax stand-in (only names the reference's class bodies mention).  Test infrastructure."""


class AbstractLinearSolver:
    pass


class LU(AbstractLinearSolver):
    pass
