from helmholtz_x_b200.eigenvectors import *  # noqa: F401,F403
from helmholtz_x_b200.eigenvectors import normalize_adjoint, normalize_eigenvector  # noqa: F401


def velocity_eigenvector(*args, **kwargs):
    raise NotImplementedError("velocity_eigenvector is post-processing outside the accelerated path (SURVEY section 2.1 #5)")
