"""ctypes argument signatures of every int-returning entry point declared in include/b200dino.h."""
import ctypes as C

P = C.c_void_p
I = C.c_int
L = C.c_longlong
F = C.c_float

SIGNATURES = {
    "b200_gemm": [P, P],
}
