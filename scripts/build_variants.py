"""Builds liblscqp_<tag>.so variants (light-instance shapes) for scripts/ab_test.sh.  Development aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

VARIANTS = {
    "full3": ("-DLSCQP_FULL_MINCTAS=3",), "full4": ("-DLSCQP_FULL_MINCTAS=4",),
    "asm2": ("-DLSCQP_ASM_MINBLOCKS=2",), "asm4": ("-DLSCQP_ASM_MINBLOCKS=4",), "asm5": ("-DLSCQP_ASM_MINBLOCKS=5",),
    "asm6": ("-DLSCQP_ASM_MINBLOCKS=6",),
    "g1k8c8": ("-DLSCQP_LIGHT_G=1", "-DLSCQP_LIGHT_KPT=8", "-DLSCQP_LIGHT_MINCTAS=8"),
    "g1k8c10": ("-DLSCQP_LIGHT_G=1", "-DLSCQP_LIGHT_KPT=8", "-DLSCQP_LIGHT_MINCTAS=10"),
    "g1k8c9": ("-DLSCQP_LIGHT_G=1", "-DLSCQP_LIGHT_KPT=8", "-DLSCQP_LIGHT_MINCTAS=9"),
    "g1k8c12": ("-DLSCQP_LIGHT_G=1", "-DLSCQP_LIGHT_KPT=8", "-DLSCQP_LIGHT_MINCTAS=12"),
    "g2k4c6": ("-DLSCQP_LIGHT_G=2", "-DLSCQP_LIGHT_KPT=4", "-DLSCQP_LIGHT_MINCTAS=6"),
    # occupancy sensitivity of the light instance: pad its shared memory so that only 7 / 6 / 5 / 4 CTAs fit per SM
    "occ7": ("-DLSCQP_LIGHT_EXTRA_SMEM=1300",), "occ6": ("-DLSCQP_LIGHT_EXTRA_SMEM=1950",), "occ5": ("-DLSCQP_LIGHT_EXTRA_SMEM=2900",),
    "occ4": ("-DLSCQP_LIGHT_EXTRA_SMEM=4300",),
    "g2k4c4": ("-DLSCQP_LIGHT_G=2", "-DLSCQP_LIGHT_KPT=4", "-DLSCQP_LIGHT_MINCTAS=4"),
}
VARIANTS.update({"kpt8": ("-DLSCQP_DAS_KPT=8",), "nockpt": ("-DLSCQP_DAS_NO_CKPT",), "das14": ("-DLSCQP_DAS_MINCTAS=14", "-DLSCQP_DAS_QMAX=28", "-DLSCQP_DAS_KPT=8"), "das13": ("-DLSCQP_DAS_MINCTAS=13", "-DLSCQP_DAS_QMAX=30", "-DLSCQP_DAS_KPT=8"),
                 "c12": ("-DLSCQP_LIGHT_MINCTAS=12",), "c16": ("-DLSCQP_LIGHT_MINCTAS=16",)})
import shutil
for tag in sys.argv[1:]:
    out = os.path.join(g.ROOT, "lsc_dr_planner_b200", f"liblscqp_{tag}.so")
    # only inst_0.cu (the M = 5 instances the bench runs) is rebuilt with the variant's defines; the other translation
    # units are taken from the default build
    objdir = os.path.join(g.ROOT, "lsc_dr_planner_b200", "_obj", os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    base = os.path.join(g.ROOT, "lsc_dr_planner_b200", "_obj", "liblscqp.so")
    for u in g.UNITS:
        if u != "inst_0.cu":
            shutil.copy2(os.path.join(base, u.replace(".cu", ".o")), objdir)
            os.utime(os.path.join(objdir, u.replace(".cu", ".o")))
    if os.path.exists(os.path.join(objdir, "inst_0.o")):
        os.unlink(os.path.join(objdir, "inst_0.o"))
    g.build_library(out, VARIANTS[tag])
    print("built", out)
