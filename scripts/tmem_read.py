"""Tensor-memory read throughput probe (development aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opz_import
opz = opz_import.load()
opz.default_context()
import ctypes as _C, os as _os
_probe = _C.CDLL(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests", "libopz_probe.so"))
fn = _probe.opz_debug_tmem_read
fn.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_double)]
fn.restype = C.c_int
for x32 in (0, 1):
    for mma in (0, 1):
        for W in (1, 4, 8, 16):
            out = (C.c_double * 2)()
            opz.check(fn(W, 2000, mma, x32, out))
            bpc = W * 32 * 64 * 4 / out[0]
            print(f"{W:2d} warps, {'x32' if x32 else 'x16'} loads, MMAs {'on ' if mma else 'off'}: {out[0]:7.1f} cycles per 64-column read per warp "
                  f"=> {bpc:6.1f} B/clk/SM" + (f"; {out[1]:5.1f} cycles/MMA" if mma else ""), flush=True)
