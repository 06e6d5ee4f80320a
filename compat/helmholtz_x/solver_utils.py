from helmholtz_x_b200.solver_utils import *  # noqa: F401,F403
