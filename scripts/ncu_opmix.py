"""Instruction mix of a kernel from its .ncu-rep (SASS page): executed warp instructions by opcode family.
python scripts/ncu_opmix.py gpurun_out/prof_das.ncu-rep [out.txt]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
name = rows[0][1]
hdr = rows[1]
i_src, i_inst, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
FAM = [("fp64 (DFMA DADD DMUL DSETP DMNMX MUFU.*64)", ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX", "MUFU")),
       ("shared memory (LDS STS)", ("LDS", "STS")), ("global / local / constant memory (LDG STG LDL STL LDC LDCU)", ("LDG", "STG", "LDL", "STL", "LDC", "LDCU")),
       ("integer / logic (IMAD IADD3 LOP3 LEA SHF ISETP IABS ...)", ("IMAD", "IADD", "LOP3", "LEA", "SHF", "ISETP", "IABS", "IMNMX", "POPC", "FLO", "PRMT", "SGXT", "VIMNMX", "VIADD", "LOP", "BREV")),
       ("uniform datapath (U*)", ("U",)), ("select / move / convert (SEL FSEL MOV F2F I2F F2I ...)", ("SEL", "FSEL", "MOV", "F2F", "I2F", "F2I", "I2FP", "F2FP", "CS2R", "S2R", "S2UR", "R2UR", "PLOP3", "P2R", "R2P")),
       ("warp exchange (SHFL REDUX VOTE MATCH)", ("SHFL", "REDUX", "VOTE", "MATCH")),
       ("control (BRA BSSY BSYNC EXIT WARPSYNC NOP CALL RET ...)", ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP", "CALL", "RET", "BRX", "JMP", "BREAK", "YIELD", "BAR", "DEPBAR", "ERRBAR", "MEMBAR")),
       ("fp32 (FFMA FADD FMUL FSETP FMNMX)", ("FFMA", "FADD", "FMUL", "FSETP", "FMNMX", "FCHK"))]
fam = collections.Counter(); ops = collections.Counter(); smp = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= i_inst or not r[i_inst].isdigit():
        continue
    op = r[i_src].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    base = o.split(".")[0]
    n = int(r[i_inst]); tot += n; ops[base] += n; smp[base] += int(r[i_smp] or 0)
    for label, keys in FAM:
        if any(base == k or (k == "U" and base.startswith("U") and base not in ("UNPACK",)) for k in keys):
            fam[label] += n
            break
    else:
        fam["other"] += n
out = [f"# instruction mix of {name}", f"# {rep}: {tot} warp instructions executed"]
for label, n in fam.most_common():
    out.append(f"{100 * n / tot:5.1f} %  {label}")
out.append("# top opcodes (share of instructions, share of stall samples)")
ts = sum(smp.values()) or 1
for o, n in ops.most_common(24):
    out.append(f"{100 * n / tot:5.1f} %  {100 * smp[o] / ts:5.1f} %  {o}")
s = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(s)
print(s)
