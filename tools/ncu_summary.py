"""Summarise an .ncu-rep (read here, no GPU) into profiles/<name>.txt: key raw metrics, stall mix, dynamic opcode mix,
hottest instructions.  usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/ncu_X.txt [units_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None     # RK steps in the profiled launch


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, unit, val = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
want = ["Kernel Name", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
lines = [f"# {rep}"]
for w in want:
    for h in hdr:
        if h == w:
            lines.append(f"{h} = {m[h][0]} [{m[h][1]}]")


def num(name):
    try:
        return float(m[name][0].replace(",", ""))
    except Exception:
        return None


rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
if rd is not None and wr is not None:
    tot = rd * scale.get(m["dram__bytes_read.sum"][1], 1) + wr * scale.get(m["dram__bytes_write.sum"][1], 1)
    lines.append(f"traffic (dram read + write per launch) = {tot:.6g} bytes")
    if units:
        lines.append(f"traffic per RK step = {tot / units:.1f} bytes")
t = num("gpu__time_duration.sum")
tms = None
if t:
    tms = t * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(m["gpu__time_duration.sum"][1], 1.0)
    if units:
        lines.append(f"warp instructions per RK step = {num('smsp__inst_executed.sum') / units:.1f}")

src = list(csv.reader(io.StringIO(ncu("source"))))
h2, data = src[1], src[2:]
ix = {h: i for i, h in enumerate(h2)}
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
# executed FP64 flops from the instruction-level counts: DFMA = 2 per thread, DADD / DMUL = 1, DMMA.8x8x4 = 512 per warp instruction
ex_thr = collections.Counter(); ex_warp = collections.Counter()
for r in data:
    sp = r[ix["Source"]].strip().split()
    op = (sp[1] if sp[0].startswith("@") else sp[0]).split(".")[0]
    ex_warp[op] += int(r[ix["Instructions Executed"]])
    ex_thr[op] += int(r[ix["Predicated-On Thread Instructions Executed"]])
flops = 2 * ex_thr["DFMA"] + ex_thr["DADD"] + ex_thr["DMUL"] + 512 * ex_warp["DMMA"]
if t:
    lines.append(f"executed FP64 flops (source page: 2 x DFMA + DADD + DMUL thread-level, 512 per DMMA.8x8x4) = {flops:.4g} per launch "
                 f"= {flops / (tms * 1e-3) * 1e-12:.2f} TFLOP/s under ncu clocks; DMMA share {512 * ex_warp['DMMA'] / max(flops, 1):.2f}")
    if units:
        lines.append(f"executed FP64 flops per RK step = {flops / units:.0f} (algorithmic: SURVEY 8d count, see bench.py roofline.flops_per_step); "
                     f"DMMA warp instructions per RK step = {ex_warp['DMMA'] / units:.2f}")
lines.append("\nstall mix (% of warp samples): " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
c = collections.Counter(); ti = 0
for r in data:
    sp = r[ix["Source"]].strip().split()
    op = (sp[1] if sp[0].startswith("@") else sp[0]).split(".")[0]
    n = int(r[ix["Instructions Executed"]]); c[op] += n; ti += n
lines.append("dynamic opcode mix (% of executed warp instructions): " + ", ".join(f"{k} {100 * v / ti:.1f}" for k, v in c.most_common(16)))
lines.append("\nhottest instructions (samples, executed, dominant stall):")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:25]):
    r = data[i]
    st = {s: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0}
    lines.append(f"  {i:5d} {r[ix['Source']].strip()[:64]:64s} {r[ix['# Samples']]:>7s} {r[ix['Instructions Executed']]:>10s} {max(st, key=st.get)[6:] if st else ''}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
