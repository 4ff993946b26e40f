"""Times the layered tensor path (csrc/opz_gemm.cu) on BASELINE configs[4] (Nz = 128, 384-1024-1024-127 relu x3) and on the
headline net in the split modes: development aid and the command the layer_gemm_kernel ncu capture replays
(scripts/ncu_capture.sh).   python scripts/layered_time.py [columns]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import opz_import  # noqa: E402
from helpers import product_model, product_params  # noqa: E402
from oracle import nde_oracle as O  # noqa: E402

opz = opz_import.load()
ctx = opz.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16


def run(name, model, pp, x0, bc, prec, dt, steps=10, warm=3, flop_rhs=0):
    xa = torch.from_numpy(x0).cuda()
    xb = torch.empty_like(xa)
    bcd = torch.from_numpy(bc).cuda()
    bufs = [xa, xb]
    for i in range(warm):
        opz.rollout_forward_dev(model, pp, bufs[i & 1].data_ptr(), bcd.data_ptr(), x0.shape[0], opz.SCHEME_RK4, dt, 1, 0, bufs[(i + 1) & 1].data_ptr(), precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(warm, warm + steps):
        opz.rollout_forward_dev(model, pp, bufs[i & 1].data_ptr(), bcd.data_ptr(), x0.shape[0], opz.SCHEME_RK4, dt, 1, 0, bufs[(i + 1) & 1].data_ptr(), precision=prec)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name}: {x0.shape[0]} columns, {ms:.2f} ms per RK4 step, {x0.shape[0] / ms * 1e3:.3e} column-timesteps/s, "
          f"{x0.shape[0] * 4 * flop_rhs / ms / 1e9:.1f} algorithmic TFLOP/s", flush=True)


p5 = O.Params(Nz=128, tau=172800.0, flags=O.FLAG_MPP | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL | O.FLAG_DIURNAL_MINUS_SCALE0)
s5 = (384, 1024, 1024, 127)
m5 = O.make_model(s5, O.ACT_RELU, seed=0, out_scale=0.03, bias_scale=0.05)
x5s, b5s = O.synthetic_columns(4096, p5, diurnal=True)
reps = (cols + 4095) // 4096
x5, b5 = np.tile(x5s, (reps, 1))[:cols].copy(), np.tile(b5s, (reps, 1))[:cols].copy()
model5 = product_model(opz, m5, Nz=128, ctx=ctx)
f5 = 2 * 3 * sum(s5[i] * s5[i + 1] for i in range(3))
for name, prec in (("configs[4] fp16", opz.PREC_FP16), ("configs[4] fp16x3", opz.PREC_FP16X3)):
    run(name, model5, product_params(opz, p5), x5, b5, prec, 10.0 / p5.tau, flop_rhs=f5)
model5.close()
if os.environ.get("OPZ_LAYERED_ALL"):
    p = O.Params(tau=691200.0)
    s1 = (96, 400, 400, 31)
    m1 = O.make_model(s1, O.ACT_RELU, seed=0, out_scale=0.1)
    x1s, b1s = O.synthetic_columns(4096, p)
    c1 = cols * 4
    r1 = (c1 + 4095) // 4096
    x1, b1 = np.tile(x1s, (r1, 1))[:c1].copy(), np.tile(b1s, (r1, 1))[:c1].copy()
    model1 = product_model(opz, m1, ctx=ctx)
    f1 = 2 * 3 * sum(s1[i] * s1[i + 1] for i in range(3))
    os.environ["OPZ_FORCE_LAYERED"] = "1"
    for name, prec in (("construct_NN fp16 (layered)", opz.PREC_FP16), ("construct_NN bf16x3", opz.PREC_BF16X3), ("construct_NN fp16x3", opz.PREC_FP16X3),
                       ("construct_NN bf16x3 again", opz.PREC_BF16X3), ("construct_NN bf16 (layered)", opz.PREC_BF16)):
        run(name, model1, product_params(opz, p), x1, b1, prec, 30.0 / p.tau, flop_rhs=f1)
    model1.close()
