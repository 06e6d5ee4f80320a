from helmholtz_x_b200.io_utils import *  # noqa: F401,F403
