#!/usr/bin/env python3
"""Per-kernel SASS opcode summary of the built C-ABI library: which kernels use the FP64 tensor path (DMMA),
the bulk-copy engine (UBLKCP = cp.async.bulk), mbarriers (SYNCS), cp.async (LDGSTS), and that no tcgen05 (UTCMMA / LDTM) or tensor-map
TMA (UTMALDG) appears -- tcgen05 has no f64 kind and the staged copies are 1-D.
usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "numcosmo_b200/lib/libncm_sd_gpu.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
dem = {}
kern, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        order.append(kern)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
KEYS = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "LDS", "STS", "STG", "SHFL", "BAR", "ATOM", "RED", "UTCMMA", "LDTM", "UTMALDG", "HMMA"]
arch = re.findall(r"arch = (sm_\w+)", out)
print(f"# {lib}: {len(order)} kernels, cubin archs {sorted(set(arch))}")
print("# opcode families counted by prefix; columns: " + " ".join(KEYS) + " | total")
for k, n in zip(order, names):
    c = counts[k]
    fam = {key: sum(v for op, v in c.items() if op.split(".")[0].startswith(key)) for key in KEYS}
    short = re.sub(r"\(anonymous namespace\)::", "", n)
    short = short[: short.index("(")] if "(" in short else short
    print(f"{short:70s} " + " ".join(f"{fam[key]:5d}" for key in KEYS) + f" | {sum(c.values())}")
tot = collections.Counter()
for k in order:
    tot.update(counts[k])
print("# library totals: " + ", ".join(f"{key}={sum(v for op, v in tot.items() if op.split('.')[0].startswith(key))}" for key in KEYS))
dm = sorted({op for op in tot if op.startswith("DMMA")})
print("# DMMA shapes: " + ", ".join(dm))
