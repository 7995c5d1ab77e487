"""Aggregate an ncu source page (cuda,sass view) per CUDA source line: stall samples and executed
warp instructions.  usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, out = None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit():
        try:
            out.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
ti, ts = sum(o[1] for o in out), sum(o[0] for o in out)
print("total warp instructions %d, stall samples %d" % (ti, ts))
for o in sorted(out, reverse=True)[:top]:
    print("smp %6d %5.1f%%  inst %9d %5.1f%%  %s:%d  %s" % (o[0], 100.0 * o[0] / max(ts, 1), o[1], 100.0 * o[1] / max(ti, 1), o[2], o[3], o[4]))
