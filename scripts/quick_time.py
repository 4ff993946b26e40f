"""Quick device-resident timing of opz_rollout_forward_dev (development aid, not the bench)."""
import argparse
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opz_import
from oracle import nde_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import product_model, product_params

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--prec", type=int, default=0)
ap.add_argument("--arch", default="construct")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
opz = opz_import.load()
p = O.Params()
sizes, act = ((96, 400, 400, 31), O.ACT_RELU) if a.arch == "construct" else ((96, 50, 20, 31), O.ACT_MISH)
m = O.make_model(sizes, act, seed=0, out_scale=0.1)
ctx = opz.default_context()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
model = product_model(opz, m)
pp = product_params(opz, p)
x0, bc = O.synthetic_columns(a.B, p)
x0d, bcd = torch.from_numpy(x0).cuda(), torch.from_numpy(bc).cuda()
outd = torch.empty((a.B, 1, 96), device="cuda")
dt = 30.0 / p.tau
flop = a.B * a.steps * 4 * (2 * 3 * sum(sizes[i] * sizes[i + 1] for i in range(len(sizes) - 1)))
for r in range(a.reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opz.rollout_forward_dev(model, pp, x0d.data_ptr(), bcd.data_ptr(), a.B, opz.SCHEME_RK4, dt, a.steps, 0, outd.data_ptr(), precision=a.prec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"B={a.B} steps={a.steps} prec={a.prec} arch={a.arch}: {ms:.3f} ms  {a.B*a.steps/ms*1e3:.3e} col-steps/s  {flop/ms/1e9:.2f} TFLOP/s", flush=True)
print("finite:", bool(torch.isfinite(outd).all()))
