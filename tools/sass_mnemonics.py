"""tools/sass_mnemonics.py — static SASS mnemonic counts (cuobjdump -sass) of the product kernels in libtsb200.so.
usage: python tools/sass_mnemonics.py [out.json]   (writes profiles/r2_sass_mnemonics_final.json by default)"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tinyspec-cling_b200", "libtsb200.so")
WANT = ("chain_kernel<double, 10, 0, false, false>", "chain_kernel<float, 10, 0, false, false>",
        "spectral_cols_kernel<8, 8, tsb::GenHash>", "spectral_rows_kernel<8, 8>", "granular_kernel", "timepar_scan")
KEYS = ["SHFL", "VOTE", "BAR", "WARPSYNC", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "LDG", "STG", "LD", "ST", "ATOM", "RED",
        "DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "FFMA2", "FADD2", "FMUL2", "MUFU", "LDL", "STL"]
ALWAYS = ("SHFL", "BAR", "UBLKCP", "LDL", "STL")

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
kern, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
out = {}
for mangled, c in counts.items():
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    if any(w in name for w in WANT):
        short = name.split("(")[0]
        out[short] = {k: c.get(k, 0) for k in KEYS if c.get(k, 0) or k in ALWAYS}
        out[short]["total"] = sum(c.values())
doc = {"what": "static SASS mnemonic counts (cuobjdump -sass libtsb200.so) of the product kernels, sm_100a", "kernels": out}
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_mnemonics_final.json")
json.dump(doc, open(path, "w"), indent=1)
for k, v in out.items():
    print(k, v)
