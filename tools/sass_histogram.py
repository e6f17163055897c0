"""SASS opcode histogram of the built library (per kernel family and in total):  python tools/sass_histogram.py > profiles/..."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "interiorpointmethod_b200/libipm_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, per, tot = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        fn = fn.replace("void ", "").replace("ipm::", "").replace("(anonymous namespace)::", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and fn:
        op = m.group(1)
        per[fn][op.split(".")[0]] += 1
        tot[op if op.startswith(("DMMA", "UTMALDG", "UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "UTMA")) else op.split(".")[0]] += 1
print("# cuobjdump -sass %s : opcode counts (whole library)" % lib)
keys = ["DMMA.8x8x4", "DFMA", "DADD", "DMUL", "MUFU", "UTMALDG.3D", "UBLKCP.S.G", "UBLKPF", "LDGSTS", "SYNCS", "LDG", "STG",
        "LDS", "STS", "LDL", "STL", "BAR", "SHFL", "ATOMG", "RED"]
for k in keys:
    n = sum(v for o, v in tot.items() if o == k or o.startswith(k + "."))
    print("%-14s %8d" % (k, n))
print("tcgen05 (UTC*MMA / LDTM): %d  - FP64 has no tcgen05 kind (ptxas rejects tcgen05.mma.kind::f64); the tensor pipe is DMMA"
      % sum(v for o, v in tot.items() if o.startswith(("UTCMMA", "UTCHMMA", "LDTM", "UTCQMMA"))))
print("\n# per kernel: DMMA / DFMA / TMA+bulk / LDL+STL (spills)")
for f, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
    tma = sum(v for o, v in c.items() if o.startswith(("UTMALDG", "UBLKCP", "UBLKPF")))
    print("%-60s insts %6d  DMMA %5d  DFMA %5d  TMA/bulk %3d  local ld/st %4d" % (f[:60], sum(c.values()), c.get("DMMA", 0),
                                                                                c.get("DFMA", 0), tma, c.get("LDL", 0) + c.get("STL", 0)))
