"""helmholtz_x/petsc4py_utils.py with the same names and the same (petsc4py) dot
convention: Vec.dot conjugates its argument."""
import numpy as np


def multiply(x0, z):
    x1 = x0.copy()
    x1.scale(z)
    return x1


def conjugate(y0):
    y1 = y0.copy()
    y1.conjugate()
    return y1


def conjugate_function(p):
    p.x.array[:] = np.conjugate(p.x.array)
    return p


def vector_vector(y0, x0):
    """y0.dot(x0) = sum_i y0_i conj(x0_i)  (petsc4py_utils.py:42-64)."""
    return y0.dot(x0)


def vector_matrix_vector(y0, A, x0):
    """y0.dot(A x0)  (petsc4py_utils.py:67-89)."""
    x1 = x0.copy()
    A.mult(x0, x1)
    return vector_vector(y0, x1)


def matrix_vector(Mat, x):
    dummy, vector = Mat.createVecs()
    Mat.mult(x, vector)
    return vector


def FixSign(x):
    """x /= x[0]/|x[0]|  (petsc4py_utils.py:100-111)."""
    x0 = x[0]
    sign = x0 / abs(x0)
    x.scale(1.0 / sign)
