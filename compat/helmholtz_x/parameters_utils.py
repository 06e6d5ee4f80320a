from helmholtz_x_b200.parameters_utils import *  # noqa: F401,F403
