from helmholtz_x_b200.shape_derivatives import *  # noqa: F401,F403
