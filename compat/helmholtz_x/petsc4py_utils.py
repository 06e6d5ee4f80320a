from helmholtz_x_b200.petsc4py_utils import *  # noqa: F401,F403
