from helmholtz_x_b200.dolfinx_utils import *  # noqa: F401,F403
