"""Summarise an ncu report of the solve kernel by source line / phase (reads a .ncu-rep through `ncu -i`)."""
import csv, subprocess, sys, re, os
rep = sys.argv[1]
src_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "lsc_dr_planner_b200", "csrc", "pdip_kernel.cuh")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = [r for r in rows if len(r) > 3 and r[0] == 'Line No'][0]
ib = hdr.index('stall_barrier')
agg = {}
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = os.path.basename(r[1]); continue
    if len(r) > 8 and r[0].isdigit() and cur == os.path.basename(src_path):
        agg[int(r[0])] = (int(r[7]), int(r[6]), int(r[ib] or 0), r[1])
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total warp-inst", tot, "samples", tots)
src = open(src_path).read().split('\n')
marks = [(i + 1, m.group(1)) for i, l in enumerate(src) for m in [re.search(r'//\s*@phase\s+(\S+)', l)] if m]
marks.append((len(src) + 1, 'end'))
for (a, name), (b, _) in zip(marks, marks[1:]):
    inst = sum(v[0] for k, v in agg.items() if a <= k < b); smp = sum(v[1] for k, v in agg.items() if a <= k < b)
    bar = sum(v[2] for k, v in agg.items() if a <= k < b)
    print("%-22s %4d-%4d inst %5.1f%%  samples %5.1f%%  (barrier %5.1f%%)" % (name, a, b - 1, 100 * inst / tot, 100 * smp / tots, 100 * bar / tots))
top = sorted(agg.items(), key=lambda kv: -(kv[1][1] - kv[1][2]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
for k, v in top:
    print("%5.1f%% inst %5.1f%% smp(non-barrier) L%d %s" % (100 * v[0] / tot, 100 * (v[1] - v[2]) / tots, k, v[3][:110]))
