"""SASS opcode histogram per kernel of libbqa_pointnet2.so -> profiles/<name>.json (CPU only: cuobjdump).

    python tools/sass_histogram.py profiles/r2_sass_histogram.json

For every kernel: instruction count and the counts of the opcodes the design claims rest on
(tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, cp.async.bulk = UBLKCP, tensor-map TMA = UTMALDG / UTMASTG,
cp.async = LDGSTS, st.async = STAS, mbarrier = SYNCS, barrier.cluster = UCGABAR, redux.sync = REDUX /
CREDUX), plus the ten most frequent opcodes.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEY = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS",
       "STAS", "SYNCS", "UCGABAR", "REDUX", "CREDUX", "VOTE", "SHFL", "ATOMS", "ATOMG", "RED", "HMMA", "FFMA", "LDG", "STG",
       "LDS", "STS", "BAR"]


def histogram(so):
    out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return sorted(arch), kernels


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True)
    return r.stdout.splitlines()


def main():
    from bridgeqa_b200 import build
    so = build.build()
    arch, kernels = histogram(so)
    pretty = demangle(list(kernels))
    rows = []
    for (name, cnt), nice in zip(kernels.items(), pretty):
        nice = re.sub(r"bqa::\(anonymous namespace\)::", "", nice)
        rows.append({"kernel": nice.split("(")[0][:120], "instructions": sum(cnt.values()),
                     "key_opcodes": {k: cnt[k] for k in KEY if cnt[k]},
                     "top10": dict(cnt.most_common(10))})
    total = collections.Counter()
    for cnt in kernels.values():
        total.update(cnt)
    doc = {"library": os.path.relpath(so, ROOT), "arch": arch, "kernels": len(rows),
           "totals": {k: total[k] for k in KEY if total[k]}, "per_kernel": rows}
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_histogram.json")
    with open(dst, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc["totals"]))


if __name__ == "__main__":
    main()
