from helmholtz_x_b200.acoustic_matrices import *  # noqa: F401,F403
