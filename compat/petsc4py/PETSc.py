import numpy as np

ScalarType = np.complex128
IntType = np.int32
