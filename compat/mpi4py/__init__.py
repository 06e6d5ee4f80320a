"""Minimal mpi4py stand-in for helmholtz-x drivers: rank/size come from torchrun's env."""
