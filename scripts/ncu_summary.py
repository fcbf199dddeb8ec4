"""Writes profiles/<tag>_<kernel>_summary.txt and profiles/traffic.json from the .ncu-rep files a
scripts/gpu_profile.sh run left in gpurun_out/."""
import csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]

def unit_bytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(u, 1)

traffic = {}
for name, rep in (("das_solve_kernel", "prof_das.ncu-rep"), ("pdip_solve_kernel", "prof_solve.ncu-rep"),
                  ("lsc_prune_kernel", "prof_asm.ncu-rep"), ("lsc_pairs_kernel", "prof_pairs.ncu-rep")):
    path = os.path.join(root, "gpurun_out", rep)
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = [f"# ncu --set full --clock-control none, one launch of {d['Kernel Name'][0]}", f"# bench command: python bench.py --steps 3 --warmup 3 --no-cpu (4096 agents, K=40, M=5, D=3)"]
    for k in KEYS:
        if k in d:
            out.append(f"{k:90s} {d[k][0]:>16s} {d[k][1]}")
    rd = unit_bytes(*d["dram__bytes_read.sum"]); wr = unit_bytes(*d["dram__bytes_write.sum"])
    traffic[name + "_bytes_per_launch"] = rd + wr
    out.append(f"dram traffic per launch (read+write): {rd + wr:.0f} bytes")
    try:
        per_cyc = {op: float(d[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"][0]) for op in ("dfma", "dadd", "dmul")}
        cycles = float(d["smsp__cycles_elapsed.max"][0])
        flops = (2 * per_cyc["dfma"] + per_cyc["dadd"] + per_cyc["dmul"]) * cycles
        peak = 2 * float(d["sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained"][0])
        traffic[name + "_fp64_flops_per_launch"] = flops
        out.append(f"fp64 flops per launch (2 dfma + dadd + dmul thread instructions): {flops:.4g}  "
                   f"= {100 * flops / cycles / peak:.2f} % of the FP64 issue peak ({peak:.0f} flop/cycle)")
    except KeyError:
        pass
    open(os.path.join(root, "profiles", f"{tag}_{name}_summary.txt"), "w").write("\n".join(out) + "\n")
src = os.path.join(root, "gpurun_out", "launches.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    open(os.path.join(root, "profiles", f"{tag}_launches.csv"), "w").writelines(lines)
json.dump(traffic, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(traffic)
