"""Times opz_loss_grad on BASELINE configs[3] (18 cases, 8 days, construct_NN, RK4) — development aid."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import opz_import
from helpers import product_model, product_params
from oracle import nde_oracle as O

opz = opz_import.load()
ctx = opz.default_context()
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 11520
B = int(sys.argv[2]) if len(sys.argv) > 2 else 18
sizes = tuple(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else (96, 400, 400, 31)
act = int(sys.argv[4]) if len(sys.argv) > 4 else 1
save_every = 90
p = O.Params(tau=691200.0, flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
x0, bc = O.synthetic_columns(B, p)
m = O.make_model(sizes, act, seed=0, out_scale=0.1)
dt = 60.0 / p.tau
model = product_model(opz, m)
pp = product_params(opz, p)
n_save = n_steps // save_every + 1
target = np.repeat(x0[:, None, :], n_save, axis=1).copy()
lw = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
for i in range(3):
    n0 = ctx.launch_count
    t = time.perf_counter()
    tot, comps, g = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, save_every)
    el = time.perf_counter() - t
    print(f"call {i}: {el:.3f} s  loss {tot:.6g} |g| {np.linalg.norm(g):.6g} finite {np.isfinite(g).all()} launches {ctx.launch_count - n0}", flush=True)
