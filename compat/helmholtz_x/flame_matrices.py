from helmholtz_x_b200.flame_matrices import *  # noqa: F401,F403
