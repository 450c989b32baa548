"""Constants shared by host modules that must not import the CUDA binding."""
INIT_STRIDE = 8
SCALAR_STRIDE = 16
N_BASES = 5
CODE_N = 255
