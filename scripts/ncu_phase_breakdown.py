"""Bucket executed warp-instructions of fte_eval_kernel by kernel phase from an ncu report:
the SASS stream (address order) is split at BAR.SYNC instructions (= __syncthreads between phases).
usage: python scripts/ncu_phase_breakdown.py report.ncu-rep [n_frames] [phase names...]"""
import csv, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]; nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 256000
names = sys.argv[3:] or ["P0", "P1a", "P1b", "P2", "P3", "P4a", "P4b", "P5"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines()) if len(r) > 10 and r[0].startswith("0x")]
ph = 0; tot = 0; stats = []
cur = dict(inst=0, smp=0, ops=Counter(), static=0, thr=0)
for r in rows:
    sass = r[1].strip()
    inst = int(r[5]); smp = int(r[4]); thr = int(r[6])
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', sass)
    op = m.group(2).split('.')[0] if m else '?'
    cur["inst"] += inst; cur["smp"] += smp; cur["ops"][op] += inst; cur["static"] += 1; cur["thr"] += thr
    tot += inst
    if op == "BAR":
        stats.append(cur); cur = dict(inst=0, smp=0, ops=Counter(), static=0, thr=0)
stats.append(cur)
tots = sum(s["smp"] for s in stats)
print(f"warp-inst {tot}  per frame {tot / nfr:.1f}")
for i, s in enumerate(stats):
    nm = names[i] if i < len(names) else f"seg{i}"
    lanes = s["thr"] / max(s["inst"], 1)
    print(f"  {nm:5s} static {s['static']:5d}  inst {s['inst'] / tot * 100:5.1f}% ({s['inst'] / nfr:7.1f}/frame, {lanes:4.1f} lanes)  stall samples {s['smp'] / max(tots, 1) * 100:5.1f}%   top: "
          + ", ".join(f"{op} {n / nfr:.0f}" for op, n in s["ops"].most_common(8)))
