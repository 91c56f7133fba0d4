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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Julia and cannot run in this
    image (no julia binary, no depot, no network), so this arm times the CPU oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ude_b200 as ude
    from oracle import oracle
    threads = oracle.max_threads()
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


def main():
    args = parse()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import ude_b200 as ude

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA library is the only implementation (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = _bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    prob, theta0, name = build_problem(ude, args, rank)
    eng = ude.Engine(prob, device=local, lanes_per_segment=args.lanes, hidden_per_lane=args.hpl, weights_in_smem=args.ws,
                     dtype=1 if args.dtype == "f32" else 0)
    exchange = "none"
    if world > 1 and prob.n_models > 1:
        exchange = "none: the models (folds x members) are independent, ranks only own different ones"
    elif world > 1:
        if args.exchange == "peer":
            # fused finalize + allreduce over peer memory: the ranks swap the IPC descriptions of their exchange buffers
            mine = torch.frombuffer(bytearray(eng.peer_export(world)), dtype=torch.uint8).to(dev)
            allb = [torch.zeros(128, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(allb, mine)
            eng.peer_attach([b.cpu().numpy().tobytes() for b in allb], world, rank)
            exchange = "fused finalize + peer-memory allreduce (ude_finalize_exchange)"
        else:
            idt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(ude.get_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            eng.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)
            exchange = "ncclAllReduce"
        if rank != 0:   # the regulariser is a function of the shared parameters: count it once (rank 0)
            prob.reg_weight = np.zeros_like(prob.reg_weight)
            eng._check(eng.lib.ude_set_model_weights(eng.handle, prob.obs_scale.ctypes.data, prob.proc_scale.ctypes.data,
                                                     prob.shoot_weight.ctypes.data, prob.reg_weight.ctypes.data))
    steps_per_eval = eng.steps_per_eval()
    n_theta = prob.n_theta

    d_theta = torch.from_numpy(theta0).to(dev)
    d_grad = torch.empty(n_theta, dtype=torch.float64, device=dev)
    d_loss = torch.empty(prob.n_models, dtype=torch.float64, device=dev)
    # a dedicated non-default stream: the C ABI treats a NULL stream as "the handle's own stream", and CUDA events
    # only see the stream they are recorded on
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step():
        eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = ude.measure_fp64_peak(local, f32=(args.dtype == "f32"))     # FP32 FMA peak in the optional FP32 mode

    # ---- device-resident timing -------------------------------------------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()
    barrier()          # ranks finish building their shards at different times; the fused exchange waits for peers (bounded)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    c_first = clocks.mark()
    eng.profile_begin(args.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    kern_ms = eng.profile_read(args.steps)
    clk = clocks.stop(c_first)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = steps_per_eval * world / (ms_step * 1e-3)
    loss_val = float(d_loss.sum().item())
    grad_l1 = float(d_grad.abs().sum().item())          # cheap checksum: lets two builds of the kernel be compared at full size

    # ---- end-to-end through the public host-buffer call ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_theta = torch.from_numpy(theta0).pin_memory()
        h_grad = torch.empty(n_theta, dtype=torch.float64).pin_memory()
        h_loss = torch.empty(prob.n_models, dtype=torch.float64).pin_memory()
        barrier()
        for _ in range(3):
            eng.loss_grad_ptr(h_theta.data_ptr(), n_theta, h_loss.data_ptr(), h_grad.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.loss_grad_ptr(h_theta.data_ptr(), n_theta, h_loss.data_ptr(), h_grad.data_ptr())
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el = float(te.item())
        e2e = {"value": steps_per_eval * world * args.steps / el, "unit": UNIT, "h2d_bytes_per_step": 8 * n_theta,
               "d2h_bytes_per_step": 8 * (n_theta + prob.n_models), "ms_per_step": 1e3 * el / args.steps,
               "api": "ude_loss_grad (host pinned theta in, loss+grad out)", "loss_check": float(h_loss.sum().item())}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        threads = oracle.max_threads()
        n_sample = args.cpu_sample or {"c2": 6000, "c3": 2500, "c3mse": 2500, "c4": 16, "c5": 6000}[args.config]
        rate, sp, ssteps, reps, el = cpu_oracle_rate(ude, args, n_sample, threads)
        rate1 = cpu_oracle_rate(ude, args, max(n_sample // 8, 8), 1, min_seconds=4.0)[0] if threads > 1 else rate
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "one_core_value": rate1,
               "sample": f"{sp.n_series} series = {ssteps} RK steps per evaluation x {reps} evaluations in {el:.1f} s "
                         f"(oracle/libude_oracle.so, OpenMP over series); Julia reference not runnable here"}

    # ---- the other BASELINE configs, same protocol (resident inputs, CUDA events), a few steps each: context, not the headline
    others = None
    if rank == 0 and world == 1 and args.config == "c3" and not args.no_other_configs:
        others = {}
        for key, cfg, n, solver2, dt2k in (("c2", "c2", 10_000, args.solver, 0), ("c4", "c4", 1000, args.solver, 0),
                                           ("c5", "c5", 1_000_000, args.solver, 0), ("c3_tsit5", "c3", 100_000, "tsit5", 0),
                                           ("c3_f32_mode", "c3", 100_000, args.solver, 1)):
            try:
                a2 = argparse.Namespace(**vars(args)); a2.config = cfg; a2.series = n; a2.solver = solver2
                p2, th2, nm2 = build_problem(ude, a2, 0)
                cfg = key
                with ude.Engine(p2, device=local, dtype=dt2k) as e2:
                    dt2 = torch.from_numpy(th2).to(dev); dg2 = torch.empty_like(dt2)
                    dl2 = torch.empty(p2.n_models, dtype=torch.float64, device=dev)
                    for _ in range(3):
                        e2.loss_grad_device(dt2.data_ptr(), dl2.data_ptr(), dg2.data_ptr(), stream)
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(10):
                        e2.loss_grad_device(dt2.data_ptr(), dl2.data_ptr(), dg2.data_ptr(), stream)
                    b.record(); torch.cuda.synchronize()
                    ms2 = a.elapsed_time(b) / 10
                    st2 = e2.steps_per_eval()
                    others[cfg] = {"workload": nm2, "steps_per_s": st2 / (ms2 * 1e-3), "ms_per_step": ms2, "kernel": e2.kernel_name(),
                                   "fp64_frac": (st2 * p2.flops_per_step() / (ms2 * 1e-3) * 1e-12 / fp64_peak) if dt2k == 0 else None}
                del dt2, dg2, dl2
            except Exception as ex:        # a side measurement must never take the headline down
                others[cfg] = {"error": str(ex)[:200]}

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
        kms = float(np.mean(kern_ms)) if len(kern_ms) else None
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
                          "cpu_affinity": numa},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
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
