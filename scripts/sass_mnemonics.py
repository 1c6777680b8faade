#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libkrylov_b200.so (cuobjdump -sass), the static evidence kept in
profiles/r01_sass_mnemonics.txt: DMMA in the Gram kernel, 128-bit read-only gathers + L2 prefetches in the SpMM."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["spmm_kernel", "gram_dmma", "combine3_norm", "pair_update_kernel", "pair_step_kernel", "slq_quadrature",
        "multi_dot", "multi_axpy", "entries_step", "frechet_step"]
PREFIXES = ("DMMA", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "CCTL", "BAR", "SHFL", "RED", "ATOM", "LDGSTS", "UTMA", "MUFU")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "krylov_robustness_b200", "libkrylov_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            cnt[cur][m.group(1)] += 1
    print("# SASS mnemonic counts of the hot kernels in libkrylov_b200.so (cuobjdump -sass, sm_100a)")
    print("# DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64); CCTL.E.PF2 = prefetch.global.L2;")
    print("# LDG.E.128.CONSTANT = 128-bit read-only gathers; DADD / DFMA = FP64 pipe")
    for fn, c in sorted(cnt.items()):
        if not any(k in fn for k in KEYS):
            continue
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:120]
        agg = collections.Counter()
        for k, v in c.items():
            if k.startswith(PREFIXES):
                agg[k if k.startswith(("LDG", "CCTL", "DMMA")) else k.split(".")[0]] += v
        print("%s\n    instructions %d | %s" % (name, sum(c.values()), ", ".join("%s %d" % kv for kv in sorted(agg.items()))))


if __name__ == "__main__":
    main()
