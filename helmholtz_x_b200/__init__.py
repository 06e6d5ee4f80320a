"""helmholtz_x_b200 -- B200-native (sm_100a) implementation of helmholtz-x's hot path
behind helmholtz-x's own Python API.  Importing the package does not need a GPU;
every compute call does (there is no CPU fallback)."""
__version__ = "0.1.0"
