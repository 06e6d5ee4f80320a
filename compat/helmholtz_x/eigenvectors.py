from helmholtz_x_b200.eigenvectors import *  # noqa: F401,F403
from helmholtz_x_b200.eigenvectors import normalize_adjoint, normalize_eigenvector, velocity_eigenvector  # noqa: F401
