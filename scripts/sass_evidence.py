"""profiles/*_sass_evidence.txt: instruction counts per kernel from `cuobjdump -sass` of the shipped library (runs without a GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "seal-3d_b200", "libseal3d_b200.so")
WANT = ["k_ffmlp_", "k_wide_", "k_adam_tables", "k_peer_", "k_ngp_", "k_grid_backward<float, 3u, 2u, true>", "k_grid_backward<__half, 3u, 2u, true>",
        "k_grid_forward<float, 3u, 2u, true, 2, true, 512>", "k_distill_rays", "k_composite_train", "k_march_", "k_occ_bbox", "k_vm_", "k_head_encode",
        "k_fixed_to_float", "k_v_stats"]
PAT = [("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("SYNCS", r"\bSYNCS"), ("REDG", r"\bREDG"),
       ("LDG.256", r"\bLDG\.[A-Z0-9.]*256"), ("STG.256", r"\bSTG\.[A-Z0-9.]*256"), ("LDGSTS", r"\bLDGSTS"), ("ATOMG", r"\bATOMG")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names, counts, instr = [], collections.defaultdict(collections.Counter), collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = cur.split("(")[0].replace("void ", "")
            names.append(cur)
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", line):
            instr[cur] += 1
            for k, p in PAT:
                if re.search(p, line):
                    counts[cur][k] += 1
    out = ["# SASS evidence (cuobjdump -sass of the shipped seal-3d_b200/libseal3d_b200.so, sm_100a): instruction counts per kernel (scripts/sass_evidence.py)",
           "#   UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTCBAR = tcgen05.commit,",
           "#   REDG = fire-and-forget global reductions (scatters, weight-gradient flushes), SYNCS = mbarrier try_wait / arrive,",
           "#   LDG.256 / STG.256 = 256-bit row accesses, LDGSTS = cp.async (the split-K weight-gradient pipeline), ATOMG = returning atomics", ""]
    for n in names:
        if any(n.startswith(w) or w in n for w in WANT):
            out.append("%-72s %6d instr  %s" % (n[:72], instr[n], "  ".join("%s=%d" % (k, counts[n][k]) for k, _ in PAT if counts[n][k])))
    # a few raw lines: TS-form MMA (A operand from tensor memory), SS-form MMA (both from shared-memory descriptors), tensor-memory stores / loads, 256-bit rows
    out += ["", "# excerpts"]
    shown = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if not cur:
            continue
        for tag, pat, fn, limit in (("TS", r"UTCHMMA tmem\[\w+\], gdesc", "k_ngp_mlp_fwd_ts", 2), ("SS", r"UTCHMMA gdesc\[\w+\], gdesc", "k_ngp_mlp_bwd", 2),
                                    ("STTM", r"\bSTTM", "k_ngp_mlp_fwd_ts", 1), ("LDTM", r"\bLDTM", "k_ngp_mlp_fwd_ts", 1), ("STG256", r"STG\.[A-Z0-9.]*256", "k_ngp_mlp_bwd", 1),
                                    ("LDG256", r"LDG\.[A-Z0-9.]*256", "k_ngp_scatter", 1), ("LDGSTS", r"LDGSTS", "k_wide_wgrad_pipe", 1), ("RED128", r"REDG\.E\.ADD\.F32x4", "k_ngp_scatter", 1)):
            if fn in cur and shown[tag] < limit and re.search(pat, line):
                shown[tag] += 1
                out.append("%-22s %s" % (fn, line.strip()))
    txt = "\n".join(out) + "\n"
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_evidence.txt")
    open(dst, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
