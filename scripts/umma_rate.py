"""tcgen05.mma / tcgen05.commit issue-cost probe (development aid): M=128 x N x K=16 f16 MMAs from one thread."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opz_import
opz = opz_import.load()
opz.default_context()
import ctypes as _C, os as _os
_probe = _C.CDLL(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests", "libopz_probe.so"))
fn = _probe.opz_debug_umma_rate
fn.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_double)]
fn.restype = C.c_int
for mode, mname in ((1, "TS"), (0, "SS")):
    for N in (32, 64, 96, 128, 256):
        for per in (6, 25, 100):
            for sd, dname in ((10, "same D"), (20, "alternating D")):
                out = (C.c_double * 2)()
                opz.check(fn(N, per, 200, mode, sd + 1, 1, out))
                print(f"{mname} N={N:3d} {dname:14s} MMAs/batch={per:3d}: total {out[0]:6.1f} cycles/MMA, issue loop {out[1]:6.1f}/MMA; floor {128 * N / 256:.0f}", flush=True)
