"""helmholtz_x/flame_transfer_function.py: FTF(omega) scalars on the host (a6)."""
import cmath
from math import factorial

import numpy as np


class nTau:
    """n * exp(i omega tau)  (flame_transfer_function.py:5-14)."""

    def __init__(self, n, tau):
        self.n = n
        self.tau = tau

    def __call__(self, omega):
        return self.n * cmath.exp(1j * omega * self.tau)

    def derivative(self, omega):
        return self.n * (1j * self.tau) * cmath.exp(1j * omega * self.tau)


class stateSpace:
    """conj(c (i conj(omega) I - A)^-1 b + d)  (flame_transfer_function.py:16-42)."""

    def __init__(self, S1, s2, s3, s4):
        self.A = np.asarray(S1)
        self.b = np.asarray(s2)
        self.c = np.asarray(s3)
        self.d = np.asarray(s4)
        self.Id = np.eye(*self.A.shape)

    def _H(self, omega, k):
        omega = np.conj(omega)
        Mat = (- 1j) ** k * factorial(k) * np.linalg.matrix_power(1j * omega * self.Id - self.A, - (k + 1))
        H = np.dot(np.dot(self.c, Mat), self.b)
        if k == 0:
            H = H + self.d
        return np.conj(H[0][0])

    def __call__(self, omega):
        return self._H(omega, 0)

    def derivative(self, omega):
        return self._H(omega, 1)
