"""Static-issue tcgen05.mma probe (development aid): 25 immediate-offset TS MMAs per batch, one thread."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opz_import
opz = opz_import.load()
opz.default_context()
fn = opz.lib.opz_debug_umma_static
fn.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
fn.restype = C.c_int
for N in (8, 16, 32, 64, 128, 256):
    out = (C.c_double * 2)()
    opz.check(fn(N, 400, out))
    print(f"TS static N={N:3d}: {out[0]:6.1f} cycles/MMA overall, {out[1]:6.1f} in the issue sequence (math floor {128 * N / 256:.0f})", flush=True)
