import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
import opz_import
opz = opz_import.load()
opz.default_context()
import ctypes as _C, os as _os
_probe = _C.CDLL(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests", "libopz_probe.so"))
fn = _probe.opz_debug_umma_rate
fn.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_double)]
fn.restype = C.c_int
for mode, mname in ((1, "TS alone"), (3, "TS + 4 warps tcgen05.ld"), (5, "TS + 4 warps tcgen05.st"), (7, "TS + ld + st")):
    for N in (64, 128, 256):
        out = (C.c_double * 2)()
        opz.check(fn(N, 100, 200, mode, 21, 1, out))
        print(f"{mname:28s} N={N:3d}: {out[0]:6.1f} cycles/MMA (floor {128 * N / 256:.0f})", flush=True)
