"""Instruction histogram by source line of one captured kernel: python tools/ncu_inst_hist.py report.ncu-rep [top]
(ncu --set full --import-source on, -lineinfo).  Lines that execute many instructions for little work are the
cheap wins in issue-bound kernels."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, lines, cur = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0] not in ("", "-") and len(r) >= len(hdr) - 2:
        lines.append((cur, r[0], r[1], dict(zip(hdr, r))))


def f(v):
    try:
        return float(v)
    except Exception:
        return 0.0


tot = sum(f(l[3].get("Instructions Executed")) for l in lines)
smp = sum(f(l[3].get("# Samples")) for l in lines)
print("total warp instructions %.0f, samples %.0f" % (tot, smp))
lines.sort(key=lambda l: -f(l[3].get("Instructions Executed")))
for fl, ln, src, d in lines[:top]:
    print("%s:%-4s %5.1f%% inst %5.1f%% samples  %s" % (fl, ln, 100 * f(d.get("Instructions Executed")) / tot,
                                                     100 * f(d.get("# Samples")) / max(smp, 1), src.strip()[:120]))
