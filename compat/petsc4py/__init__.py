"""Minimal petsc4py stand-in: drivers only touch PETSc.ScalarType."""
