"""Minimal dolfinx stand-in: drivers import dolfinx.fem.{Function, functionspace}."""
__version__ = "0.9.0-hx_b200-shim"
