from helmholtz_x_b200.flame_transfer_function import *  # noqa: F401,F403
