"""helmholtz_x/dolfinx_utils.py helpers that drivers import."""
import numpy as np

from .fem import Function
from .parameters_utils import normalize  # noqa: F401  (re-export, dolfinx_utils.py:32)


def cyl2cart(rho, phi, zeta):
    return rho * np.cos(phi), rho * np.sin(phi), zeta


def cart2cyl(x, y, z):
    return np.sqrt(x ** 2 + y ** 2), np.arctan2(y, x), z


def absolute(func):
    abs_temp = abs(func.x.array)
    out = Function(func.function_space)
    out.x.array[:] = abs_temp / np.amax(abs_temp)
    return out


def phase(func, deg=True):
    out = Function(func.function_space, name="P_angle")
    out.x.array[:] = np.angle(func.x.array, deg=deg)
    return out


def distribute_vector_as_chunks(vector):
    """dolfinx_utils.py:187-198.  One process owns every (dof, value) pair of the sparse
    flame vectors here (each GPU keeps its slice), so this is the identity."""
    return vector


def broadcast_vector(vector):
    """dolfinx_utils.py:200-207 -- identity for the same reason."""
    return vector
