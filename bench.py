#!/usr/bin/env python
"""bench.py — NDE column-timesteps/s of the forward rollout on N B200s (one process per GPU).

Workload (BASELINE.json configs[2], the configuration the metric "1/2/4/8 B200" is quoted on):
an ensemble of 2^20 columns PER GPU (Nz=32; construct_NN 96-400-400-31 relu x3, random init;
mPP diffusivity + convective adjustment; swept wind stress / buoyancy flux), sharded by column
with no inter-GPU communication ("scaling": "weak").  One bench "step" advances every column by a
window of W_t fixed-dt RK4 timesteps of the 8-day rollout (dt = 30 s), state resident on chip
within the window and in HBM between windows; successive steps continue the same rollout.

  value  : column-timesteps/s with inputs resident in HBM (device pointers, CUDA events on the
           launching stream, max over ranks).
  e2e    : the same window through the host-buffer C ABI call opz_rollout_forward (pinned host
           memory, H2D of profiles+forcing and D2H of the new profiles inside the timed region).
  roofline : the dominant (only) kernel is the rollout kernel; bound = tensor pipe for the
           batched MLP (default precision fp16 = tcgen05 kind::f16 with IEEE-half operands and FP32
           accumulation; --precision fp32 times the CUDA-core parity path instead);
           achieved = algorithmic FLOPs per launch / measured launch time.
  cpu_baseline : oracle/nde_oracle.c (a C port of the reference's per-column solve; Julia is not
           installed) on the host cores, bounded sample of the same workload.

`--impl reference` times that CPU port with all host threads on the same config instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIZES = (96, 400, 400, 31)
FLOP_PER_RHS = 2 * 3 * sum(SIZES[i] * SIZES[i + 1] for i in range(len(SIZES) - 1)) + 1800
STAGES = 4
DT_SECONDS = 30.0
TAU = 691200.0  # 8-day suite


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(columns, seed=0):
    from oracle import nde_oracle as O  # synthetic-input generator only (bench harness, not product path)
    p = O.Params(tau=TAU)
    m = O.make_model(SIZES, O.ACT_RELU, seed=seed, out_scale=0.1)
    # the forcing sweep repeats every 4096 columns (64 x 64 grid of (Q_u, Q_b), configs[1])
    x0s, bcs = O.synthetic_columns(4096, p)
    reps = (columns + 4095) // 4096
    x0 = np.tile(x0s, (reps, 1))[:columns]
    bc = np.tile(bcs, (reps, 1))[:columns]
    return O, p, m, x0, bc


def host_cores(CO):
    """Threads the CPU legs use: all cores this process may run on (not OMP_NUM_THREADS, which torchrun pins to 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, CO.num_procs()))


def run_reference(args):
    """CPU arm: the C port of the reference's per-column solve on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle as CO
    # every core of the box, whatever OMP_NUM_THREADS says: torchrun exports OMP_NUM_THREADS=1 to its workers, which in
    # round 1 made this arm a 1-core number at N > 1
    cores = host_cores(CO)
    cols = args.ref_columns or 64 * cores
    O, p, m, x0, bc = workload(cols)
    dt = DT_SECONDS / p.tau
    x = x0
    CO.rollout(x[:cores], bc[:cores], m, p, 2, dt, 1, 0, n_threads=cores)
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        out = CO.rollout(x, bc, m, p, 2, dt, args.window, 0, t0=i * args.window * dt, n_threads=cores)
        el = time.perf_counter() - t
        x = out[:, 0]
        if i >= args.warmup:
            times.append(el)
    tot = sum(times)
    value = cols * args.window * args.steps / tot
    sample = f"{cols} columns x {args.window} RK4 steps per bench step (C port of solve_NDE_mutating, one column per thread)"
    line = {
        "impl": "reference", "metric": "NDE column-timesteps/sec at Nz=32", "value": value, "unit": "column-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, cols),
        "cpu_baseline": {"value": value, "unit": "column-timesteps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "column-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def config_dict(args, columns_per_gpu):
    return {
        "workload": "BASELINE configs[2]: ensemble of 2^20 columns " + ("per GPU" if args.scaling == "weak" else "in total") +
                    " sharded by column, 8-day rollout "
                    "(tau=691200 s), construct_NN 96-400-400-31 relu x3 random init (last layer x0.1), mPP + convective "
                    "adjustment (kappa=1), RK4 dt=30 s; one step = a window of RK4 timesteps of that rollout",
        "columns_per_gpu": columns_per_gpu, "nz": 32, "scheme": "RK4", "dt_seconds": DT_SECONDS,
        "window_timesteps": args.window, "precision": args.precision,
        "l2_policy": "inputs larger than L2 (profiles alone are 384 B x columns)",
        "scaling": args.scaling,
        "parallelism": f"columns sharded over {args.gpus} GPU(s), no communication",
    }


BPTT_CASES = 18
BPTT_DT_SECONDS = 60.0
BPTT_SAVE_EVERY = 90   # 5400 s between training snapshots = `1:9:1153` of the 600 s data (train_NDE_args.jl:195)
BPTT_STEPS = 11520     # 8 days
BPTT_CPU_STEPS = 540   # sample of the CPU baseline of the gradient step


def run_bptt(args, opz, ctx, torch, dist, rank, world, stream):
    """BPTT grad-steps/s: loss + exact discrete-adjoint gradient over the 18 synthetic forcing cases
    (construct_NN, training-flavour RHS: mPP with eps, zero_weights boundary form, RK4 dt = 60 s, 8 days,
    129 snapshots, profile + gradient loss), cases sharded over the ranks, gradient + loss all-reduced (sum)."""
    from helpers import product_model, product_params
    from oracle import nde_oracle as O  # synthetic inputs only
    p = O.Params(tau=TAU, flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
    x0, bc = O.synthetic_columns(BPTT_CASES, p)
    m = O.make_model(SIZES, O.ACT_RELU, seed=0, out_scale=0.1)
    model = product_model(opz, m, ctx=ctx)
    pp = product_params(opz, p)
    dt = BPTT_DT_SECONDS / p.tau
    n_save = BPTT_STEPS // BPTT_SAVE_EVERY + 1
    # targets: the initial profiles relaxed towards a mixed state (the LES data is not available offline);
    # their values do not change the work done
    target = np.repeat(x0[:, None, :], n_save, axis=1).astype(np.float32)
    lo, hi = opz.shard_range(BPTT_CASES, rank, world)
    nb = hi - lo
    dev = torch.device("cuda", torch.cuda.current_device())
    xd = torch.from_numpy(x0[lo:hi].copy()).to(dev) if nb else torch.zeros((1, 96), device=dev)
    bd = torch.from_numpy(bc[lo:hi].copy()).to(dev) if nb else torch.zeros((1, 8), device=dev)
    td = torch.from_numpy(target[lo:hi].copy()).to(dev) if nb else torch.zeros((1, n_save, 96), device=dev)
    P3 = 3 * model.P
    buf = torch.zeros(P3 + 8, device=dev)
    lw = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)

    if dist is not None and ctx.comm_info[0] == 1:
        opz.init_library_comm(ctx, dist)   # libopz's own NCCL communicator: the all-reduce happens inside opz_loss_grad_dev

    def step():
        opz.loss_grad_dev(model, pp, xd.data_ptr(), bd.data_ptr(), nb, BPTT_CASES, td.data_ptr(), lw, opz.SCHEME_RK4, dt,
                          BPTT_STEPS, BPTT_SAVE_EVERY, buf.data_ptr() + 4 * P3, buf.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    step()  # warm-up: allocates the record buffers and builds the graphs
    barrier()
    n0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.bptt_steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.bptt_steps
    launches = (ctx.launch_count - n0) // args.bptt_steps
    if dist is not None:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    loss = float(buf[P3].item())
    gnorm = float(buf[:P3].norm().item())
    model.close()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # (N = 1 only, as for the forward cpu_baseline: under torchrun the twin ran 315 s on rank 0 — OMP_NUM_THREADS=1 is in force
        # before torch initialises its thread pool — while every other rank waited)
        # CPU baseline of the gradient step (Julia is absent; the reference differentiates with Zygote through an adjoint ODE
        # solve): the float32 PyTorch-autograd twin of the same discrete adjoint on the host cores, on a bounded sample
        # (all 18 cases, the first BPTT_CPU_STEPS time steps), scaled to the full rollout
        import time as _time
        import torch as _torch
        from oracle import nde_oracle_torch as OT
        ns = BPTT_CPU_STEPS
        from oracle import c_oracle as _CO
        _torch.set_num_threads(host_cores(_CO))   # not torchrun's OMP_NUM_THREADS=1
        tg = target[:, :ns // BPTT_SAVE_EVERY + 1]
        t0 = _time.perf_counter()
        OT.loss_and_grad(x0, bc, m, p, O.SCHEME_RK4, dt, ns, BPTT_SAVE_EVERY, tg, lw, dtype=_torch.float32)
        el = _time.perf_counter() - t0
        cpu = {"value": 1.0 / (el * BPTT_STEPS / ns), "unit": "grad-steps/s", "cores": int(_torch.get_num_threads()), "kind": "port",
               "sample": f"PyTorch-autograd twin (oracle/nde_oracle_torch.py, float32), 18 cases x first {ns} of {BPTT_STEPS} RK4 steps, "
                         f"{el:.1f} s, scaled linearly"}
    # algorithmic FLOPs: forward sweep + backward (dX and dW) = 3 x stages x F_rhs per column-step
    flops = 3.0 * STAGES * FLOP_PER_RHS * BPTT_CASES * BPTT_STEPS
    return {"metric": "BPTT grad-steps/sec", "value": 1e3 / ms, "unit": "grad-steps/s", "ms_per_grad_step": ms,
            "scaling": "strong", "cases": BPTT_CASES, "cases_this_rank": nb, "timesteps": BPTT_STEPS, "scheme": "RK4",
            "dt_seconds": BPTT_DT_SECONDS, "snapshots": n_save, "precision": "fp32", "timed_steps": args.bptt_steps,
            "gpu_launches_per_step": int(launches), "loss": loss, "grad_norm": gnorm,
            "finite": bool(np.isfinite(loss) and np.isfinite(gnorm)),
            "achieved_tflops": flops / (ms * 1e-3) / 1e12,
            "collective": "none (1 GPU)" if dist is None else
                          f"NCCL all-reduce(sum) of {P3} + 7 fp32 per step, enqueued INSIDE opz_loss_grad_dev (libopz binds NCCL; "
                          f"{ctx.comm_info[2]} groups so far)",
            "cpu_baseline": cpu,
            "note": "latency-bound: 18 columns x 46080 sequential right-hand sides; see DESIGN.md section 6"}


_RESULT_FD = None


def quiet_stdout():
    """Everything any library writes to file descriptor 1 during the run (NCCL prints its version banner there when
    NCCL_DEBUG is set on the box) goes to stderr; the one JSON line is written to the original stdout by emit()."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


def time_windows(opz, torch, dist, stream, ctx, model, pp, x0, bc_dev, cols, window_steps, prec, dt, steps, warmup,
                 sampler=None):
    """`warmup` untimed + `steps` timed windows of `window_steps` RK4 timesteps over `cols` device-resident columns,
    continuing one rollout (ping-pong between two profile buffers).  CUDA events on the launching stream, barrier +
    synchronize on both sides.  Returns (total_ms, per-launch ms, final profiles (device), launches, clocks)."""
    xa = torch.from_numpy(x0[:cols]).cuda()
    xb = torch.empty_like(xa)
    bufs = [xa, xb]

    def window(i):
        opz.rollout_forward_dev(model, pp, bufs[i & 1].data_ptr(), bc_dev.data_ptr(), cols, opz.SCHEME_RK4, dt, window_steps, 0,
                                bufs[(i + 1) & 1].data_ptr(), t0=i * window_steps * dt, precision=prec)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        window(i)
    barrier()
    if sampler is not None:
        sampler.start()
    n0 = ctx.launch_count
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record(stream)
    for k in range(steps):
        window(warmup + k)
        evs[k + 1].record(stream)
    barrier()
    launches = ctx.launch_count - n0
    clocks = sampler.stop() if sampler is not None else None
    total_ms = evs[0].elapsed_time(evs[-1])
    kernel_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]
    return total_ms, kernel_ms, bufs[(warmup + steps) & 1], launches, clocks


def max_over_ranks(torch, dist, *vals):
    if dist is None:
        return vals if len(vals) > 1 else vals[0]
    tt = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out = tuple(float(v) for v in tt)
    return out if len(vals) > 1 else out[0]


def traffic_table():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from `ncu --set full` captures of this
    command committed under profiles/ (profiles/traffic.json: key "<columns>x<window>:<precision>" -> bytes, source file).
    The figure belongs to the kernel build and shape it was captured on; absent key -> null."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return {}


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--columns", type=int, default=1 << 20, help="columns per GPU (weak) or in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --columns per GPU (the headline line); strong: --columns in total, sharded over the GPUs "
                         "(BASELINE configs[2]'s own wording).  At N > 1 the other mode is measured too and reported in the same line.")
    ap.add_argument("--window", type=int, default=0, help="RK4 timesteps per bench step (0 = per-precision default)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "fp16", "bf16"])
    ap.add_argument("--ref-columns", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bptt", action="store_true", help="skip the BPTT grad-steps/s leg")
    ap.add_argument("--no-small", action="store_true", help="skip the informational small-net leg")
    ap.add_argument("--no-fp32", action="store_true", help="skip the FP32 (CUDA-core parity path) leg")
    ap.add_argument("--no-layered", action="store_true", help="skip the layered tensor path legs (near-FP32 split mode; configs[4] Nz=128)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed result")
    ap.add_argument("--bptt-steps", type=int, default=2, help="timed gradient steps of the BPTT leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        # the CPU arm is rank 0's alone: the other ranks of a torchrun launch exit without work
        if int(os.environ.get("RANK", "0")) != 0:
            return
        if args.window == 0:
            args.window = 64
        run_reference(args)
        return

    import torch
    import opz_import
    from helpers import product_model, product_params

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    opz = opz_import.load()
    ctx = opz.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    prec_name = args.precision
    if prec_name == "auto":
        prec_name = os.environ.get("OPZ_BENCH_PRECISION", "fp16")
    prec = {"fp32": opz.PREC_FP32, "bf16": opz.PREC_BF16, "fp16": opz.PREC_FP16}[prec_name]
    args.precision = prec_name
    if args.window == 0:
        args.window = 1 if prec_name == "fp32" else 16
    # columns this rank steps: weak = --columns each; strong = its contiguous shard of --columns (no communication either way)
    cols_weak = args.columns
    lo, hi = opz.shard_range(args.columns, rank, world)
    cols_strong = hi - lo
    cols = cols_weak if args.scaling == "weak" else cols_strong
    O, p, m, x0, bc = workload(cols_weak, seed=0)
    model = product_model(opz, m, ctx=ctx)
    pp = product_params(opz, p)
    dt = DT_SECONDS / p.tau
    bcd = torch.from_numpy(bc).cuda()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (the headline `value`) ----
    sampler = ClockSampler(local_rank)
    total_ms, kernel_ms, final, launches, clocks = time_windows(opz, torch, dist, stream, ctx, model, pp, x0, bcd, cols, args.window,
                                                                prec, dt, args.steps, args.warmup, sampler)
    finite = bool(torch.isfinite(final).all().item())

    # ---- parity of the TIMED result: the columns of one period of the forcing sweep that the C oracle re-solves on the host
    #      over exactly the (warmup + steps) x window timesteps the GPU just took (outside every timed region) ----
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import c_oracle as CO
        n_done = (args.warmup + args.steps) * args.window
        pick = np.arange(0, min(cols, 4096), 64)
        got = final[torch.from_numpy(pick).cuda()].cpu().numpy()[:, None, :]
        t = time.perf_counter()
        ref = CO.rollout(x0[pick], bc[pick], m, p, 2, dt, n_done, 0, n_threads=host_cores(CO))
        parity = {"err": float(O.profile_error(got, ref, 32)), "columns": int(pick.size), "timesteps": int(n_done),
                  "norm": "max over columns and variables of max_k|x - ref| / max_k|ref| on the scaled final state (SURVEY 8d)",
                  "against": "oracle/nde_oracle.c Float32 solve of the same columns (every 64th of the 4096-column forcing sweep)",
                  "tolerance": 1e-4 if prec_name == "fp32" else 1e-2, "oracle_seconds": round(time.perf_counter() - t, 2),
                  "full_rollout": "profiles/r02_parity_8day.json: the same comparison over all 23 040 steps (tests/test_gpu_long.py)"}

    # ---- the other scaling mode (N > 1 only; at N = 1 the two are the same run) ----
    other = None
    if world > 1:
        ocols = cols_strong if args.scaling == "weak" else cols_weak
        o_ms, _, ofinal, _, _ = time_windows(opz, torch, dist, stream, ctx, model, pp, x0, bcd, ocols, args.window, prec, dt,
                                             max(3, args.steps // 2), 3)
        o_ms = max_over_ranks(torch, dist, o_ms / max(3, args.steps // 2))
        tot_cols = args.columns if args.scaling == "weak" else args.columns * world
        other = {"scaling": "strong" if args.scaling == "weak" else "weak", "columns_total": tot_cols,
                 "columns_this_rank": ocols, "value": tot_cols * args.window / (o_ms * 1e-3), "unit": "column-timesteps/s",
                 "ms_per_step": o_ms, "finite": bool(torch.isfinite(ofinal).all().item()),
                 "note": "BASELINE configs[2] wording (2^20 columns in total sharded over the GPUs)" if args.scaling == "weak" else
                         "fixed columns per GPU"}
        del ofinal

    # ---- FP32 leg: the CUDA-core parity path (<= 1e-4 contract) on the same workload, FP32-FMA roofline ----
    fp32 = None
    if not args.no_fp32 and prec_name != "fp32":
        fcols = min(cols, 1 << 19)   # 201 MB of profiles: larger than L2
        f_steps = 3
        f_ms, f_k, ffinal, _, _ = time_windows(opz, torch, dist, stream, ctx, model, pp, x0, bcd, fcols, 1, opz.PREC_FP32, dt, f_steps, 1)
        f_ms = max_over_ranks(torch, dist, f_ms / f_steps)
        f_flops = fcols * STAGES * FLOP_PER_RHS
        sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        f_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        fp32 = {"kernel": "rollout_fp32_kernel (FP32 FMA on CUDA cores)", "value": fcols * world / (f_ms * 1e-3),
                "unit": "column-timesteps/s", "columns_per_gpu": fcols, "window_timesteps": 1, "ms_per_step": f_ms,
                "roofline": {"bound": "fp32_fma", "achieved": f_flops / (f_ms * 1e-3) / 1e12, "peak": f_peak, "unit": "TFLOP/s",
                             "frac": f_flops / (f_ms * 1e-3) / 1e12 / f_peak,
                             "peak_source": f"148 SMs x 128 FMA lanes x 2 x {sm_mhz:.0f} MHz (max SM clock)"},
                "finite": bool(torch.isfinite(ffinal).all().item())}
        del ffinal

    # ---- layered tensor path (csrc/opz_gemm.cu): (i) the near-FP32 split-operand mode on the headline workload, to be read
    #      against the FP32 leg (same <= 1e-4 contract); (ii) BASELINE configs[4]: Nz = 128 columns, 384-1024-1024-127 nets ----
    layered = None
    if not args.no_layered:
        layered = {}
        lcols = min(cols, 1 << 18)            # 100 MB of profiles: larger than L2 together with the operand images
        l_steps = 3
        l_ms, _, lfinal, _, _ = time_windows(opz, torch, dist, stream, ctx, model, pp, x0, bcd, lcols, 1, opz.PREC_FP16X3, dt, l_steps, 2)
        l_ms = max_over_ranks(torch, dist, l_ms / l_steps)
        l_flops = lcols * STAGES * FLOP_PER_RHS
        peak_t = float(measured_peaks()[0].get("bf16_tflops_sustained", 1379.5))
        layered["near_fp32"] = {
            "kernel": "layer_gemm_kernel x 3 layers + pack_state_kernel + physics_rk_kernel per right-hand side (tcgen05, A and B from shared memory)",
            "precision": "fp16x3 (both operands split into two IEEE halves, 3 MMAs per k-step, FP32 accumulate)",
            "value": lcols * world / (l_ms * 1e-3), "unit": "column-timesteps/s", "columns_per_gpu": lcols, "window_timesteps": 1,
            "ms_per_step": l_ms, "algorithmic_tflops": l_flops / (l_ms * 1e-3) / 1e12,
            "tensor_tflops_issued": 3 * l_flops * (448 * 448 + 96 * 448 + 448 * 64) / (400 * 400 + 96 * 400 + 400 * 31) / (l_ms * 1e-3) / 1e12,
            "frac_of_tensor_peak_issued": 3 * l_flops * (448 * 448 + 96 * 448 + 448 * 64) / (400 * 400 + 96 * 400 + 400 * 31) / (l_ms * 1e-3) / 1e12 / peak_t,
            "finite": bool(torch.isfinite(lfinal).all().item())}
        if fp32:
            layered["near_fp32"]["speedup_over_fp32_path"] = layered["near_fp32"]["value"] / fp32["value"]
        if rank == 0 and not args.no_parity:
            from oracle import c_oracle as CO
            pick = np.arange(0, min(lcols, 4096), 64)
            got = lfinal[torch.from_numpy(pick).cuda()].cpu().numpy()[:, None, :]
            ref = CO.rollout(x0[pick], bc[pick], m, p, 2, dt, 2 + l_steps, 0, n_threads=host_cores(CO))
            layered["near_fp32"]["parity_err"] = float(O.profile_error(got, ref, 32))
        del lfinal
        # configs[4]
        p5 = O.Params(Nz=128, tau=172800.0, flags=O.FLAG_MPP | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL | O.FLAG_DIURNAL_MINUS_SCALE0)
        sizes5 = (384, 1024, 1024, 127)
        m5 = O.make_model(sizes5, O.ACT_RELU, seed=0, out_scale=0.03, bias_scale=0.05)
        model5 = product_model(opz, m5, Nz=128, ctx=ctx)
        pp5 = product_params(opz, p5)
        c5 = min(cols, 1 << 16)               # 100 MB of profiles (1 536 B per column)
        x5s, b5s = O.synthetic_columns(4096, p5, diurnal=True)
        x5 = np.tile(x5s, ((c5 + 4095) // 4096, 1))[:c5]
        b5 = torch.from_numpy(np.tile(b5s, ((c5 + 4095) // 4096, 1))[:c5].copy()).cuda()
        dt5 = 10.0 / p5.tau
        f5 = 2 * 3 * sum(sizes5[i] * sizes5[i + 1] for i in range(3)) + 7000
        for pname, pcode in (("fp16", opz.PREC_FP16), ("fp16x3", opz.PREC_FP16X3)):
            q_ms, _, qfinal, _, _ = time_windows(opz, torch, dist, stream, ctx, model5, pp5, x5, b5, c5, 1, pcode, dt5, 3, 2)
            q_ms = max_over_ranks(torch, dist, q_ms / 3)
            nmma = 3 if pname == "fp16x3" else 1
            layered["config5_" + pname] = {
                "workload": "BASELINE configs[4]: Nz=128, 384-1024-1024-127 relu x3, mPP, diurnal forcing, RK4 dt=10 s, one timestep per bench step",
                "value": c5 * world / (q_ms * 1e-3), "unit": "column-timesteps/s", "columns_per_gpu": c5, "ms_per_step": q_ms,
                "algorithmic_tflops": c5 * STAGES * f5 / (q_ms * 1e-3) / 1e12,
                "frac_of_tensor_peak_issued": nmma * c5 * STAGES * (f5 - 7000) * (128.0 / 127.0) / (q_ms * 1e-3) / 1e12 / peak_t,
                "finite": bool(torch.isfinite(qfinal).all().item())}
            del qfinal
        model5.close()

    # ---- the reference's production net on the same ensemble (96-50-20-31 mish, train_NDE.jl:103) ----
    small = None
    if not args.no_small:
        ms_small = O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=0, out_scale=0.1)
        model_s = product_model(opz, ms_small, ctx=ctx)
        s_steps = max(3, min(args.steps, 10))
        s_ms, _, sfinal, _, _ = time_windows(opz, torch, dist, stream, ctx, model_s, pp, x0, bcd, cols, args.window, prec, dt, s_steps, 3)
        sms = max_over_ranks(torch, dist, s_ms / s_steps)
        fl_small = 2 * 3 * (96 * 50 + 50 * 20 + 20 * 31) + 1800
        tot = cols * world if args.scaling == "weak" else args.columns
        small = {"net": "96-50-20-31 mish x3 (train_NDE.jl:103)", "value": tot * args.window / (sms * 1e-3),
                 "unit": "column-timesteps/s", "ms_per_step": sms, "precision": prec_name,
                 "achieved_tflops": cols * args.window * STAGES * fl_small / (sms * 1e-3) / 1e12,
                 "finite": bool(torch.isfinite(sfinal).all().item()),
                 "note": "0.16 MFLOP per column-timestep: bound by the activation / physics instruction issue, not by the tensor pipe"}
        if rank == 0 and not args.no_parity:
            from oracle import c_oracle as CO
            n_done = (3 + s_steps) * args.window
            pick = np.arange(0, min(cols, 4096), 64)
            got = sfinal[torch.from_numpy(pick).cuda()].cpu().numpy()[:, None, :]
            ref = CO.rollout(x0[pick], bc[pick], ms_small, p, 2, dt, n_done, 0, n_threads=host_cores(CO))
            small["parity_err"] = float(O.profile_error(got, ref, 32))
            small["parity_timesteps"] = int(n_done)
        model_s.close()
        del sfinal

    # ---- end-to-end through the host-buffer C ABI ----
    e2e = None
    if not args.no_e2e:
        # two pinned host profile buffers used alternately: the caller's next window starts from the profiles the previous
        # call returned, so they swap roles instead of being copied on the host (torchrun pins OMP_NUM_THREADS=1, which made
        # that 403 MB host memcpy cost more than the host<->device copies it sat between)
        hp = [torch.from_numpy(x0[:cols]).pin_memory(), torch.empty((cols, 96), dtype=torch.float32).pin_memory()]
        hb = torch.from_numpy(bc[:cols]).pin_memory()
        import ctypes as C

        def e2e_step(i):
            hx, ho = hp[i % 2], hp[(i + 1) % 2]
            opz.check(opz.lib.opz_rollout_forward(ctx._h, model._h, C.byref(pp), C.c_void_p(hx.data_ptr()),
                                                  C.c_void_p(hb.data_ptr()), cols, opz.SCHEME_RK4, dt, i * args.window * dt,
                                                  args.window, 0, prec, C.c_void_p(ho.data_ptr())))

        e2e_step(0)
        barrier()
        t = time.perf_counter()
        e_steps = max(2, min(args.steps, 5))
        for k in range(e_steps):
            e2e_step(1 + k)
        barrier()
        e2e_ms = (time.perf_counter() - t) * 1e3 / e_steps
        e2e = {"ms": e2e_ms, "h2d": int(x0[:cols].nbytes + bc[:cols].nbytes), "d2h": int(x0[:cols].nbytes)}

    # ---- BPTT leg: BASELINE configs[3], 18 cases sharded over the ranks, one NCCL all-reduce per step ----
    bptt = None
    if not args.no_bptt:
        bptt = run_bptt(args, opz, ctx, torch, dist, rank, world, stream)

    # ---- max over ranks ----
    t_ms, e_ms = max_over_ranks(torch, dist, total_ms, e2e["ms"] if e2e else 0.0)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    total_cols = cols * world if args.scaling == "weak" else args.columns
    colsteps_per_step = total_cols * args.window
    value = colsteps_per_step * args.steps / (t_ms * 1e-3)
    flop_per_launch = cols * args.window * STAGES * FLOP_PER_RHS
    avg_kernel_ms = float(np.mean(kernel_ms))
    achieved = flop_per_launch / (avg_kernel_ms * 1e-3) / 1e12
    if prec_name == "fp32":
        sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        bound, peak_source = "fp32_fma", f"148 SMs x 128 FMA lanes x 2 x {sm_mhz:.0f} MHz (max SM clock)"
        kernel_name = "rollout_fp32_kernel (FP32 FMA on CUDA cores)"
    else:
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        bound = "tensor"
        peak_source = (f"{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)" if peak_kind == "measured" else
                       "of fallback: MEASURED_PEAKS.json absent, B200_PROFILING.md's stated fallback")
        kernel_name = "rollout_tc2_kernel (tcgen05 cta_group::2, " + prec_name + " operands, fp32 accumulate)"
    tr = traffic_table().get(f"{cols}x{args.window}:{prec_name}")
    line = {
        "metric": "NDE column-timesteps/sec at Nz=32", "value": value, "unit": "column-timesteps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[prec_name], "data": "synthetic",
        "config": config_dict(args, cols),
        "clocks": clocks, "gpu_launches": int(launches), "finite": finite,
        "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": tr["bytes"] if tr else None, "traffic_unit": "bytes of DRAM traffic per launch (ncu)",
                     "traffic_source": tr["source"] if tr else None, "peak_source": peak_source, "kernel": kernel_name,
                     "flop_per_launch": flop_per_launch, "avg_launch_ms": avg_kernel_ms},
    }
    if parity:
        line["parity_err"] = parity["err"]
        line["parity"] = parity
    if e2e:
        line["e2e"] = {"value": colsteps_per_step / (e_ms * 1e-3), "unit": "column-timesteps/s",
                       "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e_ms}
    if other:
        line["other_scaling"] = other
    if fp32:
        line["fp32"] = fp32
    if layered:
        line["layered_path"] = layered
    if bptt:
        line["bptt"] = bptt
    if small:
        line["small_net"] = small
    if world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle as CO
        cores = host_cores(CO)
        cb_cols = 64 * cores
        t = time.perf_counter()
        CO.rollout(x0[:cb_cols], bc[:cb_cols], m, p, 2, dt, 4, 0, n_threads=cores)  # pilot: sizes the sample to ~12 s
        pilot = time.perf_counter() - t
        cb_steps = int(min(1024, max(8, 4 * 12.0 / max(pilot, 1e-3))))
        t = time.perf_counter()
        CO.rollout(x0[:cb_cols], bc[:cb_cols], m, p, 2, dt, cb_steps, 0, n_threads=cores)
        el = time.perf_counter() - t
        line["cpu_baseline"] = {"value": cb_cols * cb_steps / el, "unit": "column-timesteps/s", "cores": cores, "kind": "port",
                                "sample": f"first {cb_cols} columns x {cb_steps} RK4 steps, oracle/nde_oracle.c (C port; Julia absent), "
                                          f"{el:.1f} s"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
