#!/usr/bin/env python
"""Which data-movement / reduction instructions each hot kernel of libstereo_b200.so actually contains
(cuobjdump -sass, sm_100a): profiles/<tag>_sass.md.   python scripts/sass_summary.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stereo_b200", "lib", "libstereo_b200.so")
MNEMONICS = ["UBLKCP", "UBLKPF", "UTMALDG", "UTMAPF", "LDGSTS", "SYNCS", "CREDUX", "REDUX", "SHFL", "IDP", "BAR", "ATOM", "ATOMG", "RED",
             "LDG", "STG", "LDS", "STS", "LDL", "STL", "NANOSLEEP", "MUFU"]
HOT = ["gsweep_kernel<float, 6", "gsweep_kernel<float, 8", "gsweep_kernel<float, 2", "sweep_kernel<float, 2", "ncc_levels_kernel<4>",
       "ncc_levels_kernel<2>", "ncc_volume_kernel<float>", "window_stats_kernel", "bfs_tile_kernel", "push_kernel<false>",
       "collect_relabel_kernel<false>", "pairwise_tables_kernel", "gpair_tables_kernel<float, 6", "gnode_tables_kernel<float, 6"]


def main(tag):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    names = subprocess.run(["cu++filt"], input="\n".join(f.split("\n", 1)[0].strip() for f in funcs), capture_output=True, text=True).stdout.split("\n")
    rows = []
    for f, name in zip(funcs, names):
        name = name.replace("(int)", "")
        if not any(h in name for h in HOT):
            continue
        cnt = collections.Counter()
        n = 0
        for line in f.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                n += 1
                cnt[m.group(1)] += 1
        rows.append((name, n, cnt))
    rows.sort(key=lambda r: r[0])
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass.md"), "w") as out:
        out.write(f"# SASS mnemonic counts of the hot kernels ({tag}; cuobjdump -sass of libstereo_b200.so, sm_100a)\n\n"
                  "UBLKCP = cp.async.bulk (TMA bulk copy), UBLKPF = cp.async.bulk.prefetch.L2, UTMALDG = cp.async.bulk.tensor (tensor-map copy), "
                  "SYNCS = mbarrier, CREDUX / REDUX = redux.sync, IDP = dp4a, LDL / STL = local memory (spills).\n\n")
        out.write("| kernel | instr | " + " | ".join(MNEMONICS) + " |\n|---|---|" + "---|" * len(MNEMONICS) + "\n")
        for name, n, cnt in rows:
            short = re.sub(r"\(.*$", "", name).replace("sb::", "").replace("(anonymous namespace)::", "")
            out.write(f"| `{short}` | {n} | " + " | ".join(str(cnt.get(m, 0)) for m in MNEMONICS) + " |\n")
    print("wrote", f"profiles/{tag}_sass.md", len(rows), "kernels")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "rX")
