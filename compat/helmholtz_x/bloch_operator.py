from helmholtz_x_b200.bloch_operator import Blochifier  # noqa: F401
