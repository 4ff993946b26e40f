"""Operand-precision study of the tensor-core MLP (CPU emulation with the NumPy oracle; TEST/STUDY infrastructure).

Which rounding makes the reduced-precision rollout drift from the Float32 solve over the full 8-day horizon of BASELINE
configs[2]?  The oracle's right-hand side is run with the GEMM operands of chosen layers rounded to a 16-bit format, or
split into hi + lo 16-bit parts (the split modes multiply hi and lo parts separately and accumulate in FP32, as the
tensor path does), for a few columns of the sweep, and compared with the committed Float32 / double solutions.

  python scripts/precision_study.py MODE [n_columns] [out.json]
MODE: comma-separated key=value, keys x (state fed to layer 1), h (hidden activations), w (weights):
      values f16 | bf16 | f16x2 | bf16x2 | bf16x3 | f32      e.g.  x=f16x2,h=f16,w=f16
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nde_oracle as O  # noqa: E402


def rounders(kind):
    q = {"f16": O.quant_fp16, "bf16": O.quant_bf16}
    if kind == "f32":
        return None
    if kind in q:
        return lambda a: [q[kind](a)]
    base = q[kind[:-2]]
    n = int(kind[-1])

    def split(a):
        parts, r = [], np.asarray(a, dtype=np.float32)
        for _ in range(n):
            p = base(r)
            parts.append(p)
            r = (r - p).astype(np.float32)
        return parts
    return split


def dither_tapes(W, n_tapes, rng):
    """Ordered temporal dither of the FP16 rounding: tape t holds, per weight, the FP16 neighbour below or above so that the
    MEAN over the n_tapes tapes equals the weight to log2(n_tapes) more bits; thresholds in bit-reversed order, rotated per
    weight so that the weights do not all switch in the same phase."""
    W = np.asarray(W, dtype=np.float32)
    q = W.astype(np.float16)
    qf = q.astype(np.float32)
    lo = np.where(qf > W, np.nextafter(q, np.float16(-np.inf)), q).astype(np.float16)
    hi = np.nextafter(lo, np.float16(np.inf)).astype(np.float16)
    lof, hif = lo.astype(np.float32), hi.astype(np.float32)
    frac = np.where(hif > lof, (W - lof) / np.maximum(hif - lof, 1e-30), 0.0)
    bits = int(np.log2(n_tapes))
    rev = np.array([int(format(t, f"0{bits}b")[::-1], 2) if bits else 0 for t in range(n_tapes)])
    rot = rng.integers(0, n_tapes, size=W.shape)
    tapes = []
    for t in range(n_tapes):
        thr = (rev[(t + rot) % n_tapes] + 0.5) / n_tapes
        tapes.append(np.where(frac > thr, hif, lof).astype(np.float32))
    return tapes


def weyl_fp16(a, t):
    """FP16 rounding with a temporal (Weyl-sequence) dither: magnitude bits + threshold(t, element) then truncation of the 13
    dropped mantissa bits; the mean over consecutive t approaches the FP32 value."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).copy()
    idx = np.arange(a.shape[-1], dtype=np.uint32)
    thr = (np.uint32(t) * np.uint32(5061) + idx * np.uint32(2897)) & np.uint32(8191)
    u = (u + thr) & np.uint32(0xFFFFE000)
    return u.view(np.float32).astype(np.float16).astype(np.float32)


def make_mlp(mode):
    xd, hd = mode["x"] == "f16w", mode["h"] == "f16w"
    rx = None if xd else rounders(mode["x"])
    rh = None if hd else rounders(mode["h"])
    dn = int(mode["w"][4:]) if mode["w"].startswith("f16d") else 0
    rw = None if dn else rounders(mode["w"])
    cache = {}
    calls = [0]
    rng = np.random.default_rng(1234)

    def mlp_forward(x, flat, sizes, act, quant=None):
        key = id(flat)
        if key not in cache:
            layers = O.unpack_weights(flat, sizes)
            if dn:
                cache[key] = [([[w.T.copy()] for w in dither_tapes(W, dn, rng)], b) for W, b in layers]
            else:
                cache[key] = [([W.T.copy()] if rw is None else [p.T.copy() for p in rw(W)], b) for W, b in layers]
        r = calls[0] // 3          # right-hand-side counter (three nets per right-hand side)
        calls[0] += 1
        tape = ((r // 4) + (r % 4)) % dn if dn else 0     # tape(n, s) = (n + s) mod N: every tape visits every RK4 stage slot
        h = x
        for li, (Wp, b) in enumerate(cache[key]):
            if dn:
                Wp = Wp[tape]
            ra = rx if li == 0 else rh
            if (li == 0 and xd) or (li > 0 and hd):
                ap = [weyl_fp16(h, r)]
            else:
                ap = [h] if ra is None else ra(h)
            z = None
            # hi*hi first, then the cross terms whose order sum (i + j) stays below the number of parts kept (bf16x3: hi*lo, lo*hi, ...)
            nmax = max(len(ap), len(Wp))
            for i, a in enumerate(ap):
                for j, w in enumerate(Wp):
                    if i + j < nmax:
                        t = a @ w
                        z = t if z is None else z + t
            z = z + b
            h = O.activation(z, act) if li < len(cache[key]) - 1 else z
        return h
    return mlp_forward


def main():
    mode = {"x": "f16", "h": "f16", "w": "f16"}
    for kv in sys.argv[1].split(","):
        k, v = kv.split("=")
        mode[k] = v
    ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    g = np.load(os.path.join(ROOT, "tests", "golden", "config3_rk4_8day.npz"))
    p = O.Params(tau=float(g["tau"]), flags=int(g["flags"]))
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    sel = np.linspace(0, 63, ncol).astype(int)
    O.mlp_forward = make_mlp(mode)
    t = time.time()
    out = O.rollout(g["x0"][sel], g["bc"][sel], m, p, O.SCHEME_RK4, float(g["dt_nd"]), int(g["n_steps"]), int(g["save_every"]), quant=True)
    el = time.time() - t
    e32 = [float(O.profile_error(out[:, k:k + 1], g["sol32"][sel][:, k:k + 1], 32)) for k in range(out.shape[1])]
    e64 = [float(O.profile_error(out[:, k:k + 1], g["sol64"][sel][:, k:k + 1], 32)) for k in range(out.shape[1])]
    res = {"mode": mode, "columns": sel.tolist(), "seconds": el, "err_vs_f32": e32, "err_vs_f64": e64, "max_vs_f32": max(e32)}
    print(json.dumps({"mode": mode, "max_vs_f32": max(e32), "max_vs_f64": max(e64), "day2": e32[6], "day4": e32[12], "day8": e32[24], "s": round(el)}))
    if len(sys.argv) > 3:
        json.dump(res, open(sys.argv[3], "w"))


if __name__ == "__main__":
    main()
