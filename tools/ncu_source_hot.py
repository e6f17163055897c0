"""Hot source lines of one captured kernel:  python tools/ncu_source_hot.py report.ncu-rep [top]
Reads `ncu --page source --print-source cuda,sass --csv` (needs -lineinfo and --import-source on) and prints the
source lines with the most warp-stall samples and their dominant stall reasons."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur_file, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        print("##", r[1])
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0] not in ("", "-") and len(r) >= len(hdr) - 2:
        d = dict(zip(hdr, r))
        lines.append((cur_file, r[0], r[1], d))


def f(v):
    try:
        return float(v)
    except Exception:
        return 0.0


tot = sum(f(l[3].get("# Samples")) for l in lines)
print("total samples", tot)
lines.sort(key=lambda l: -f(l[3].get("# Samples")))
for fl, ln, src, d in lines[:top]:
    st = sorted(((k, f(v)) for k, v in d.items() if k.startswith("stall_") and "(Not" not in k and f(v) > 0), key=lambda kv: -kv[1])[:3]
    print("%s:%-4s %5.1f%%  inst %-9s %-100s %s" % (fl, ln, 100 * f(d.get("# Samples")) / max(tot, 1), d.get("Instructions Executed", ""),
                                                   src.strip()[:100], " ".join("%s=%d" % (k[6:], v) for k, v in st)))
