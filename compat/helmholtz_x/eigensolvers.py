from helmholtz_x_b200.eigensolvers import *  # noqa: F401,F403
