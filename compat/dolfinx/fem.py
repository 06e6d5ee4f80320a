from helmholtz_x_b200.fem import DG0Space, Function, FunctionSpace, functionspace  # noqa: F401
