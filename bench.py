#!/usr/bin/env python
"""bench.py -- RK trajectory-steps/sec (loss + adjoint gradient, FP64) of the UDE hot path on N B200s.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N > 1 under torchrun, one rank per GPU) prints ONE
JSON line.  A "step" is one loss+gradient evaluation of the workload BASELINE.json's metric is quoted on: config C3,
MultiNODE with one covariate, d = 4, hidden 32, 10^5 synthetic series x 30 points per GPU, multiple-shooting loss
(pred_length 5), RK4 with 4 substeps per observation interval => 1.16e7 RK steps per evaluation per GPU.
Weak scaling: every rank holds its own 10^5 series; the process-model gradient (+ loss) is sum-allreduced over
NCCL by the library on the launching stream once per step.

  value     : whole-job steps/s, theta and all inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       : same metric through the public host-buffer C-ABI call (ude_loss_grad): pinned host theta -> H2D ->
              kernels -> D2H(loss, grad) inside the timed region
  roofline  : the segment kernel against the MEASURED FP64 FMA peak (ude_measure_fp64_peak, run live; MEASURED_PEAKS.json
              has no FP64 entry) -- the path is FP64-issue bound, not HBM bound; HBM figures are added
  cpu_baseline : the CPU oracle (a port: the Julia reference cannot run here) on a bounded sample, all host cores
`--impl reference` times that CPU oracle alone (rank 0 only under torchrun).
"""
from __future__ import annotations

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

METRIC = "RK trajectory-steps/sec (loss+adjoint grad, FP64)"
UNIT = "steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c3mse", "c4", "c5"])
    ap.add_argument("--series", type=int, default=0, help="series (c2/c3), members (c4) or segments (c5) PER GPU; 0 = BASELINE size")
    ap.add_argument("--solver", default="rk4", choices=["rk4", "tsit5"])
    ap.add_argument("--substeps", type=int, default=4)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--hpl", type=int, default=0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: fused finalize + peer-memory exchange kernel (default) or finalize + ncclAllReduce + scatter")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="f32 = the optional FP32 mode (1e-4 parity bar)")
    ap.add_argument("--ws", type=int, default=0, help="0 auto, 1 register-resident weights, 2 shared-memory weights")
    ap.add_argument("--cpu-sample", type=int, default=0, help="series in the CPU-oracle sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dense-e2e", action="store_true", help="e2e through dense copies of theta and gradient (no UDE_OPT_SPARSE_UHAT_IO)")
    ap.add_argument("--total-series", type=int, default=0, help="N>1: series IN TOTAL of the strong-scaling C3 measurement (default 100000)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short C2/C4/C5 side measurements (N = 1 only)")
    return ap.parse_args()


def build_problem(ude, args, rank, n_override=None):
    P = ude.problem
    solver = P.RK4 if args.solver == "rk4" else P.TSIT5
    seed = 1000 + rank
    cfg = args.config
    n = n_override if n_override else args.series
    if cfg == "c2":
        prob, th = ude.synthetic.config_c2(n_series=n or 10_000, seed=seed, solver=solver, substeps=args.substeps)
        name = "C2 NODE d=2 H=10, multiple shooting pl=5, %d LV series x 60 pts" % prob.n_series
    elif cfg in ("c3", "c3mse"):
        prob, th = ude.synthetic.config_c3(n_series=n or 100_000, seed=seed, solver=solver, substeps=args.substeps,
                                           loss="multiple shooting" if cfg == "c3" else "mse")
        name = "C3 MultiNODE+1 covariate d=4 NN 5->32->4, %s, %d series x 30 pts per GPU" % (
            "multiple shooting pl=5" if cfg == "c3" else "default MSE state-space loss", prob.n_series)
    elif cfg == "c4":
        prob, th = ude.synthetic.config_c4(n_members=n or 1000, seed=seed, solver=solver, substeps=args.substeps)
        name = "C4 coral UDE d=3 NN 4->10->2, conditional likelihood, %d members x 10 LFO folds per GPU" % (prob.n_models // 10)
    else:
        prob, th = ude.synthetic.config_c5(n_segments=n or 1_000_000, seed=seed, solver=solver, substeps=args.substeps)
        name = "C5 wide NODE d=8 H=128, %d one-interval segments per GPU" % prob.n_series
    # every rank must start from the SAME process model (shared parameters): take rank 0's initialisation
    if prob.n_models == 1:
        pm0 = ude.synthetic.init_pm(prob, np.random.default_rng(4242))
        prob.pm_view(th, 0)[:] = pm0
    return prob, th, name


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a
    thread of this process every 10 ms (the same counters nvidia-smi prints; a piped `nvidia-smi -lms` delivers its lines
    in 4 KB bursts, too late to attribute them to a 100 ms window); nvidia-smi once at the end is the fallback."""

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                index = int(vis.split(",")[index])
            except Exception:
                pass
        self.index = index
        self.samples = []          # (sm_mhz, sm_max_mhz, power_w, reason_mask)
        self.run = False
        self.nv = None
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.run = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((sm, self.mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def mark(self):
        """index of the next sample: call at the start of the timed region (the sampler starts before warm-up)"""
        return len(self.samples)

    def stop(self, first=0):
        if self.nv is None:
            return self._smi_once()
        self.run = False
        self.t.join(timeout=1)
        nv = self.nv
        rows = self.samples[first:] or self.samples[-1:]
        names = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        reasons = sorted({nm for r in rows for nm, bit in names if r[3] & bit})
        sm = [r[0] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": rows[0][1] if rows else None,
                "power_w_max": max(r[2] for r in rows) if rows else None, "samples": len(sm), "reasons": reasons, "source": "nvml"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "power_w_max": float(f[2]), "samples": 1,
                    "reasons": sorted(nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")),
                    "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def cpu_oracle_rate(ude, args, n_sample, threads, min_seconds=12.0, max_reps=400):
    """steps/s of the CPU oracle (loss + adjoint gradient) on an n_sample-series instance of the same workload."""
    from oracle import oracle
    prob, th, _ = build_problem(ude, args, 0, n_override=n_sample)
    steps = prob.count_steps()
    oracle.loss_grad(prob, th, threads=threads)          # warm-up (page-in, thread pool)
    t0 = time.perf_counter(); reps = 0
    while True:
        oracle.loss_grad(prob, th, threads=threads)
        reps += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or reps >= max_reps:
            break
    return steps * reps / el, prob, steps, reps, el


def host_threads():
    """Host threads the CPU arm may use: the affinity mask of this process.  torchrun exports OMP_NUM_THREADS=1 to every
    rank, which must not shrink the reference arm to one core (the N>1 ratios compared N GPUs with ONE core)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Julia and cannot run in this
    image (no julia binary, no depot, no network), so this arm times the CPU oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ude_b200 as ude
    from oracle import oracle
    threads = host_threads()
    # one "step" = one loss+grad of a bounded sample of the workload (sized for ~0.3 s of wall time per step)
    n_sample = args.cpu_sample or {"c2": 6000, "c3": 2500, "c3mse": 2500, "c4": 16, "c5": 6000}[args.config]
    prob, th, name = build_problem(ude, args, 0, n_override=n_sample)
    steps = prob.count_steps()
    for _ in range(args.warmup):
        oracle.loss_grad(prob, th, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.loss_grad(prob, th, threads=threads)
    el = time.perf_counter() - t0
    value = steps * args.steps / el
    sample = f"{prob.n_series} series ({steps} RK steps) of the workload per step, oracle/libude_oracle.so, {threads} OpenMP threads"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": name + " (bounded CPU sample)", "solver": args.solver, "substeps": args.substeps,
                      "same_config": False,
                      "sample": f"{prob.n_series} of the GPU arm's 100000 series per GPU: a RATE metric (steps/s), so the sample size does "
                                f"not bias it; the arm runs on rank 0 only and does not grow with --gpus",
                      "host_threads": threads, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                      "note": "Julia reference not runnable here (parity unpinned); CPU oracle port timed instead"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout.  Point fd 1 at stderr for
    the duration of the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode()); sys.stdout.flush()


class Ctx:
    """torch / distributed plumbing shared by every measurement of a run"""

    def __init__(self, args):
        import torch
        import ude_b200 as ude
        self.torch, self.ude, self.args = torch, ude, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: the CUDA library is the only implementation (no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = _bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        # a dedicated non-default stream: the C ABI treats a NULL stream as "the handle's own stream", and CUDA events
        # only see the stream they are recorded on
        self.tstream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def engine(self, prob, dtype=0):
        """Engine for this rank's shard + the cross-rank exchange of [loss | shared process-model gradient] (single-model
        problems only; folds x members shard by model and exchange nothing).  Collective: every rank calls it."""
        torch, dist, ude, args = self.torch, self.dist, self.ude, self.args
        eng = ude.Engine(prob, device=self.local, lanes_per_segment=args.lanes, hidden_per_lane=args.hpl, weights_in_smem=args.ws, dtype=dtype)
        exchange = "none"
        if self.world > 1 and prob.n_models > 1:
            exchange = "none: the models (folds x members) are independent, ranks only own different ones"
        elif self.world > 1:
            if args.exchange == "peer":
                # fused finalize + allreduce over peer memory: the ranks swap the IPC descriptions of their exchange buffers
                mine = torch.frombuffer(bytearray(eng.peer_export(self.world)), dtype=torch.uint8).to(self.dev)
                allb = [torch.zeros(128, dtype=torch.uint8, device=self.dev) for _ in range(self.world)]
                dist.all_gather(allb, mine)
                eng.peer_attach([b.cpu().numpy().tobytes() for b in allb], self.world, self.rank)
                exchange = "fused finalize + peer-memory allreduce (ude_finalize_exchange)"
            else:
                idt = torch.zeros(128, dtype=torch.uint8, device=self.dev)
                if self.rank == 0:
                    idt.copy_(torch.frombuffer(bytearray(ude.get_unique_id()), dtype=torch.uint8))
                dist.broadcast(idt, 0)
                eng.comm_init(bytes(idt.cpu().numpy().tobytes()), self.world, self.rank)
                exchange = "ncclAllReduce"
            if self.rank != 0:   # the regulariser is a function of the shared parameters: count it once (rank 0)
                prob.reg_weight = np.zeros_like(prob.reg_weight)
                eng._check(eng.lib.ude_set_model_weights(eng.handle, prob.obs_scale.ctypes.data, prob.proc_scale.ctypes.data,
                                                         prob.shoot_weight.ctypes.data, prob.reg_weight.ctypes.data))
        return eng, exchange


def time_resident(cx, eng, prob, theta0, steps, warmup, clocks=None):
    """K evaluations with theta / loss / grad resident in HBM: CUDA events on the launching stream, barrier +
    synchronize on both sides, max over ranks.  Returns a dict (ms_step, kernel_ms, loss, grad_l1, clocks)."""
    torch = cx.torch
    d_theta = torch.from_numpy(theta0).to(cx.dev)
    d_grad = torch.empty(prob.n_theta, dtype=torch.float64, device=cx.dev)
    d_loss = torch.empty(prob.n_models, dtype=torch.float64, device=cx.dev)

    def step():
        eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), cx.stream)

    cx.barrier()     # ranks finish building their shards at different times; the fused exchange waits for peers (bounded)
    for _ in range(max(warmup, 3)):
        step()
    cx.barrier()
    c_first = clocks.mark() if clocks else 0
    eng.profile_begin(steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    cx.barrier()
    ms_total = e0.elapsed_time(e1)
    kern_ms = eng.profile_read(steps)
    clk = clocks.stop(c_first) if clocks else None
    ms_step = cx.max_over_ranks(ms_total) / steps
    out = {"ms_step": ms_step, "kernel_ms": float(np.mean(kern_ms)) if len(kern_ms) else None, "clocks": clk,
           "loss": float(d_loss.sum().item()),
           "grad_l1": float(d_grad.abs().sum().item())}    # cheap checksum: lets two builds of the kernel be compared at full size
    del d_theta, d_grad, d_loss
    return out


def time_e2e(cx, eng, prob, theta0, steps, sparse=True):
    """The same metric through the public host-buffer C-ABI call (ude_loss_grad): page-locked host theta in, loss + gradient
    out, every step.  For the (multiple-)shooting losses UDE_OPT_SPARSE_UHAT_IO is on: only the block-start columns of uhat
    are ever read / have a gradient (LFC.jl:281-296, :665-680), so the kernels fetch and store exactly those columns over PCIe
    (zero-copy) and the caller's gradient buffer is zeroed ONCE, before the first call."""
    torch = cx.torch
    n = prob.n_theta
    h_theta = torch.from_numpy(theta0).pin_memory()
    h_grad = torch.zeros(n, dtype=torch.float64).pin_memory()
    h_loss = torch.empty(prob.n_models, dtype=torch.float64).pin_memory()
    if sparse:
        eng.set_option(cx.ude.OPT_SPARSE_UHAT_IO, 1)
    cx.barrier()
    for _ in range(3):
        eng.loss_grad_ptr(h_theta.data_ptr(), n, h_loss.data_ptr(), h_grad.data_ptr())
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.loss_grad_ptr(h_theta.data_ptr(), n, h_loss.data_ptr(), h_grad.data_ptr())
    torch.cuda.synchronize()
    el = cx.max_over_ranks(time.perf_counter() - t0)
    h2d, d2h = eng.last_io_bytes()
    if sparse:
        eng.set_option(cx.ude.OPT_SPARSE_UHAT_IO, 0)
    dense = h2d == 8 * n
    return {"value": eng.steps_per_eval() * cx.world * steps / el, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": 1e3 * el / steps,
            "api": (("ude_loss_grad (page-locked host theta in, per-model losses + gradient out; dense copies of theta and gradient, "
                     "pipelined in runs of whole models: H2D / kernel + finalize / D2H overlap)") if dense and prob.n_models > 1 else
                    "ude_loss_grad (page-locked host theta in, loss + gradient out; dense copies of theta and gradient)" if dense else
                    "ude_loss_grad with UDE_OPT_SPARSE_UHAT_IO (page-locked host theta in, loss + gradient out; the block-start uhat "
                    "columns are read / their gradient written zero-copy by the kernels, process model copied; gradient buffer "
                    "zeroed once by the caller, structurally-zero entries never rewritten)"),
            "loss_check": float(h_loss.sum().item()), "grad_l1_check": float(h_grad.abs().sum().item())}


def side_config(cx, key, cfg, n_total, solver, dtype_k, steps=10, e2e=False):
    """One of the other BASELINE configs at its BASELINE size, split evenly over the ranks (strong scaling: by series /
    by (member, fold) / by segment), same protocol as the headline."""
    args = cx.args
    a2 = argparse.Namespace(**vars(args)); a2.config = cfg; a2.series = max(n_total // cx.world, 1); a2.solver = solver
    p2, th2, nm2 = build_problem(cx.ude, a2, cx.rank)
    eng, exch = cx.engine(p2, dtype=dtype_k)
    try:
        r = time_resident(cx, eng, p2, th2, steps, 3)
        st = eng.steps_per_eval() * cx.world
        fp = cx.fp64_peak
        out = {"workload": nm2 + (f" -- BASELINE size split over {cx.world} GPUs" if cx.world > 1 else ""), "steps_per_s": st / (r["ms_step"] * 1e-3),
               "ms_per_step": r["ms_step"], "kernel": eng.kernel_name(), "kernel_ms": r["kernel_ms"], "n_gpus": cx.world,
               "fp64_frac": (st * p2.flops_per_step() / (r["ms_step"] * 1e-3) * 1e-12 / (fp * cx.world)) if dtype_k == 0 else None}
        if e2e:
            out["e2e"] = time_e2e(cx, eng, p2, th2, steps)
        return out
    finally:
        cx.barrier()
        eng.close()


def main():
    args = parse()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    cx = Ctx(args)
    torch, ude, dist, world, rank, local = cx.torch, cx.ude, cx.dist, cx.world, cx.rank, cx.local

    prob, theta0, name = build_problem(ude, args, rank)
    eng, exchange = cx.engine(prob, dtype=1 if args.dtype == "f32" else 0)
    steps_per_eval = eng.steps_per_eval()
    n_theta = prob.n_theta
    fp64_peak = ude.measure_fp64_peak(local, f32=(args.dtype == "f32"))     # FP32 FMA peak in the optional FP32 mode
    cx.fp64_peak = ude.measure_fp64_peak(local) if args.dtype == "f32" else fp64_peak

    # ---- device-resident timing -------------------------------------------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()
    res = time_resident(cx, eng, prob, theta0, args.steps, args.warmup, clocks)
    ms_step, clk = res["ms_step"], res["clocks"]
    value = steps_per_eval * world / (ms_step * 1e-3)
    loss_val, grad_l1 = res["loss"], res["grad_l1"]

    # ---- end-to-end through the public host-buffer call ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = time_e2e(cx, eng, prob, theta0, args.steps, sparse=not args.dense_e2e)
        if rank == 0 and world == 1 and not args.dense_e2e and e2e["h2d_bytes_per_step"] != 8 * n_theta:
            d = time_e2e(cx, eng, prob, theta0, args.steps, sparse=False)      # context: the dense-copy path (round 1's e2e)
            e2e["dense_copy_path"] = {k: d[k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step")}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        n_sample = args.cpu_sample or {"c2": 6000, "c3": 2500, "c3mse": 2500, "c4": 16, "c5": 6000}[args.config]
        rate, sp, ssteps, reps, el = cpu_oracle_rate(ude, args, n_sample, threads)
        rate1 = cpu_oracle_rate(ude, args, max(n_sample // 8, 8), 1, min_seconds=4.0)[0] if threads > 1 else rate
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "one_core_value": rate1,
               "sample": f"{sp.n_series} series = {ssteps} RK steps per evaluation x {reps} evaluations in {el:.1f} s "
                         f"(oracle/libude_oracle.so, OpenMP over series); Julia reference not runnable here"}

    # ---- strong scaling of the configuration the north star names: C3 with 10^5 series IN TOTAL, sharded by series ----
    strong = None
    if args.config == "c3" and world > 1 and not args.no_other_configs:
        try:
            a2 = argparse.Namespace(**vars(args)); a2.series = max((args.total_series or 100_000) // world, 1)
            p2, th2, nm2 = build_problem(ude, a2, rank)
            e2, _ = cx.engine(p2)
            k2 = max(args.steps, 50)
            r2 = time_resident(cx, e2, p2, th2, k2, 5)
            st2 = e2.steps_per_eval() * world
            strong = {"scaling": "strong", "workload": f"C3, {a2.series * world} series x 30 pts IN TOTAL, {a2.series} per GPU (sharded by series)",
                      "value": st2 / (r2["ms_step"] * 1e-3), "unit": UNIT, "ms_per_step": r2["ms_step"], "steps": k2,
                      "segment_kernel_ms": r2["kernel_ms"],
                      "outside_the_segment_kernel_ms": (r2["ms_step"] - r2["kernel_ms"]) if r2["kernel_ms"] else None,
                      "limiter": "per step: one persistent segment-kernel launch + the fused finalize/exchange launch; what does not shrink with "
                                 "1/N is launch latency, the block-partial reduction + peer exchange (~10-20 us) and the tail of the persistent "
                                 "grid (tasks per warp no longer a large integer)",
                      "e2e": None if args.no_e2e else time_e2e(cx, e2, p2, th2, k2)}
            cx.barrier()
            e2.close()
        except Exception as ex:
            strong = {"error": str(ex)[:300]}

    # ---- the other BASELINE configs, same protocol (resident inputs, CUDA events), a few steps each: context, not the headline
    others = None
    if args.config == "c3" and not args.no_other_configs:
        others = {}
        todo = [("c2", "c2", 10_000, args.solver, 0), ("c4", "c4", 1000, args.solver, 0), ("c5", "c5", 1_000_000, args.solver, 0),
                ("c5_f32_mode", "c5", 1_000_000, args.solver, 1)]
        if world == 1:
            todo += [("c3_tsit5", "c3", 100_000, "tsit5", 0), ("c3_f32_mode", "c3", 100_000, args.solver, 1)]
        for key, cfg, n, solver2, dt2k in todo:
            try:
                others[key] = side_config(cx, key, cfg, n, solver2, dt2k, e2e=(key in ("c2", "c4", "c5") and not args.no_e2e))
            except Exception as ex:        # a side measurement must never take the headline down
                others[key] = {"error": str(ex)[:200]}
                cx.barrier()
    if others is not None and rank == 0 and world == 1:
        try:       # UKF marginal likelihood on the C1 tutorial model (host buffers in and out, loss + gradient)
            p3, th3 = ude.synthetic.config_c1()
            with ude.Engine(p3, device=local) as e3:
                F3, Pe3 = 0.3 * np.eye(p3.d), 0.1 * np.eye(p3.d)
                for _ in range(3):
                    e3.ukf_loss_grad(th3, F3, Pe3)
                t0 = time.perf_counter()
                for _ in range(20):
                    e3.ukf_loss_grad(th3, F3, Pe3)
                others["ukf_c1"] = {"workload": "C1 LV tutorial UDE, UKF marginal likelihood (ude_ukf_loss_grad), 59 filter steps x 5 sigma points",
                                    "ms_per_eval": (time.perf_counter() - t0) / 20 * 1e3}
        except Exception as ex:
            others["ukf_c1"] = {"error": str(ex)[:200]}
        try:       # a right-hand side written as expressions (generated trait -> nvcc plugin) against the hand-written kernel of the same model
            P = ude.problem
            rngu = np.random.default_rng(0)
            Yu, tsu = ude.synthetic.lotka_volterra_data(rngu, 20000, 30, random_u0=True)
            t0 = time.perf_counter()
            urhs = ude.expression_rhs(d=2, nn_out=1, scalars=("r", "m", "theta"), du=["r*u1 - y1", "theta*y1 - m*u2"])
            ude.user_rhs.ensure_kernels(urhs, 10, 0)
            ures = {"workload": "LV UDE (test/UDETests.jl:20-28) NN 2->10->1, multiple shooting pl=5, 20000 series x 30 pts",
                   "generate_and_load_s": time.perf_counter() - t0}
            for label, kind, ur in (("hand_written", P.RHS_LV_UDE, None), ("generated", urhs.kind, urhs)):
                pu = P.Problem(d=2, nn_in=2, nn_hidden=10, nn_out=1, rhs_kind=kind, n_scalars=3, solver=P.RK4, substeps=4,
                               loss_kind=P.LOSS_MULTI_SHOOT, pred_length=5).default_layout()
                pu.user_rhs = ur
                ude.synthetic._tile_series(pu, Yu, tsu)
                pu.finalize()
                thu = ude.synthetic.theta_from(pu, ude.synthetic.init_pm(pu, np.random.default_rng(1), scalars=(1.0, 0.5, 0.5)), np.asarray(pu.data))
                eu, _ = cx.engine(pu)
                try:
                    ru = time_resident(cx, eu, pu, thu, 10, 3)
                    ures[label] = {"kernel": eu.kernel_name(), "steps_per_s": eu.steps_per_eval() / (ru["ms_step"] * 1e-3), "ms_per_step": ru["ms_step"]}
                finally:
                    eu.close()
            others["user_rhs_lv"] = ures
        except Exception as ex:
            others["user_rhs_lv"] = {"error": str(ex)[:200]}
        try:       # the same loss on a many-model handle: 100 members x 10 leave-future-out folds of the coral UDE in one launch chain
            p4, th4 = ude.synthetic.config_c4(n_members=100, n_folds=10)
            with ude.Engine(p4, device=local) as e4:
                F4, Pe4 = 0.3 * np.eye(p4.d), 0.1 * np.eye(p4.d)
                for _ in range(3):
                    e4.ukf_loss_grad(th4, F4, Pe4)
                t0 = time.perf_counter()
                for _ in range(10):
                    e4.ukf_loss_grad(th4, F4, Pe4)
                ms = (time.perf_counter() - t0) / 10 * 1e3
                n_upd = int(np.sum(np.asarray(p4.lengths) - 1))
                others["ukf_c4_ensemble"] = {"workload": "C4 coral UDE, UKF marginal likelihood, 100 members x 10 folds = 1000 models in one handle "
                                                         "(per-model process model and P_nu factor), 7 sigma points",
                                             "ms_per_eval": ms, "filter_updates_per_s": n_upd / (ms * 1e-3),
                                             "rk_steps_per_s": 2 * n_upd * (2 * p4.d + 1) * p4.substeps / (ms * 1e-3)}
        except Exception as ex:
            others["ukf_c4_ensemble"] = {"error": str(ex)[:200]}
        try:       # the tutorial workflow itself: 500 Adam iterations of the C1 conditional-likelihood fit, device resident
            p4, th4 = ude.synthetic.config_c1()
            with ude.Engine(p4, device=local) as e4:
                e4.adam(th4, 0.025, 50)
                t0 = time.perf_counter()
                _, tr4 = e4.adam(th4, 0.025, 500)
                el = time.perf_counter() - t0
                others["c1_train_500_adam_iterations"] = {
                    "workload": "C1 LV tutorial UDE (59 segments), conditional likelihood, ude_adam: theta, moments and gradient stay on the GPU",
                    "ms_total": el * 1e3, "us_per_iteration": el / 500 * 1e6, "cuda_graph": bool(e4.adam_used_graph()),
                    "loss_first": float(tr4.ravel()[0]), "loss_last": float(tr4.ravel()[-1])}
        except Exception as ex:
            others["c1_train_500_adam_iterations"] = {"error": str(ex)[:200]}

    if rank == 0:
        kms = res["kernel_ms"]
        flops_step = prob.flops_per_step()
        achieved = steps_per_eval * flops_step / (kms * 1e-3) * 1e-12 if kms else None
        traffic = None      # dram__bytes_read+write of the dominant kernel per launch, from the committed ncu capture (per RK step x steps)
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                per_step = json.load(open(tf)).get(eng.kernel_name(), {}).get("dram_bytes_per_rk_step")
                traffic = per_step * steps_per_eval if per_step else None
            except Exception:
                traffic = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = prob.hbm_bytes_per_eval()
        is_mma = "mma" in eng.kernel_name()
        roofline = {"bound": ("tensor" if is_mma else "fp64") if args.dtype == "f64" else "fp32",
                    "pipe": ("FP64 tensor-core DMMA m8n8k4 + DFMA: ONE shared FP64 datapath (profiles/microbench_r01b.json), so the FP64 "
                             "FMA peak is the denominator, not the bf16 tensor peak") if is_mma else "FP64 FMA" if args.dtype == "f64" else "FP32 FMA",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (achieved / fp64_peak) if achieved else None, "traffic": traffic,
                    "peak_source": ("measured live: ude_measure_fp64_peak (DFMA chain, 8-way ILP, 64 warps/SM); vendor 37.2" if args.dtype == "f64"
                                    else "measured live: ude_measure_fp32_peak (FFMA chain, 8-way ILP, 64 warps/SM)"),
                    "flops_per_step": flops_step, "kernel": eng_kernel_name(eng), "kernel_ms": kms,
                    "kernel_share_of_step": (kms / ms_step) if kms else None,
                    "hbm_algorithmic_bytes": hbm_bytes, "hbm_achieved_gbs": (hbm_bytes / (kms * 1e-3) * 1e-9) if kms else None,
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
               "data": "synthetic",
               "config": {"workload": name, "solver": args.solver, "substeps": args.substeps,
                          "rk_steps_per_eval_per_gpu": steps_per_eval, "segments_per_gpu": prob.count_segments(),
                          "n_theta_per_gpu": n_theta, "loss_value": loss_val, "grad_l1": grad_l1,
                          "l2": "inputs+stash larger than L2 (theta, grad, data = %.0f MB per step)" % (3 * 8 * prob.d * prob.n_cols / 1e6),
                          "parallelism": ("single GPU" if world == 1 else (f"dp{world} by model, no exchange ({exchange})" if prob.n_models > 1 else
                                          f"dp{world} by series, {prob.n_pm + 1} doubles exchanged per step: {exchange}")),
                          "cpu_affinity": cx.numa},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "strong_scaling": strong,
               "gpu_launches": eng.launches_per_eval() * args.steps, "clocks": clk, "other_configs": others}
        emit(out)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _bind_to_gpu_numa_node(index):
    """One rank per GPU: pin this process (and the pinned host buffers it first-touches) to the CPUs NVML reports as local to
    its GPU, so the 192 MB per step of the host-buffer path do not cross the socket interconnect.  Best effort."""
    if os.environ.get("UDE_BENCH_NO_AFFINITY"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[index]) if vis else index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cpus {allowed[0]}-{allowed[-1]}"
    except Exception as ex:       # noqa: BLE001
        return f"unavailable ({type(ex).__name__})"
    return None


def eng_kernel_name(eng):
    nm = eng.kernel_name()
    fam = "ude_mma_kernel (FP64 DMMA m8n8k4)" if "mma" in nm else ("ude_seg_kernel (FFMA + warp shuffles)" if nm.startswith("f32") else "ude_seg_kernel (DFMA + warp shuffles)")
    return f"{fam} <{nm}>: fused RK forward + loss + discrete adjoint"


if __name__ == "__main__":
    main()
