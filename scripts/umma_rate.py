"""tcgen05.mma / tcgen05.commit issue-cost probe (development aid): M=128 x N x K=16 f16 MMAs from one thread."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opz_import
opz = opz_import.load()
opz.default_context()
fn = opz.lib.opz_debug_umma_rate
fn.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_double)]
fn.restype = C.c_int
for N in (64, 128):
    for per in (1, 3, 6, 12, 25):
        for commits in (0, 1, 2):
            out = (C.c_double * 2)()
            opz.check(fn(N, per, 200, 1, 10 + commits, 1, out))
            print(f"TS N={N:3d} MMAs/batch={per:3d} commits/batch={commits}: {out[1] * per:8.1f} cycles per batch in the issue loop ({out[1]:6.1f}/MMA; floor {128 * N / 256:.0f})", flush=True)
