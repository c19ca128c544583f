"""Shared-memory wavefronts (actual / ideal) of fte_eval by phase and opcode from an ncu report's source page; the SASS stream is
split at BAR.SYNC.   usage: python scripts/ncu_smem_by_phase.py report.ncu-rep [n_frames] [phase names...]"""
import csv, subprocess, re, sys
from collections import Counter
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 256000
names = sys.argv[3:] or ["pro", "P0", "P1b", "P2", "P3", "P4", "P5"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
acc = []; cur = dict(wf=0, ideal=0, inst=0, smp=0, ops=Counter(), opsi=Counter())
for r in data:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1].strip()); op = m.group(2) if m else '?'
    wf = int(r[ix["L1 Wavefronts Shared"]] or 0); idl = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    cur["wf"] += wf; cur["ideal"] += idl; cur["inst"] += int(r[ix["Instructions Executed"]]); cur["smp"] += int(r[ix["# Samples"]])
    if wf: cur["ops"][op] += wf; cur["opsi"][op] += idl
    if op.startswith("BAR"):
        acc.append(cur); cur = dict(wf=0, ideal=0, inst=0, smp=0, ops=Counter(), opsi=Counter())
acc.append(cur)
tot = sum(a["wf"] for a in acc); ti = sum(a["inst"] for a in acc); ts = sum(a["smp"] for a in acc)
print("shared wavefronts per frame %.1f (ideal %.1f), warp-instructions per frame %.1f" % (tot / n, sum(a["ideal"] for a in acc) / n, ti / n))
for i, a in enumerate(acc):
    nm = names[i] if i < len(names) else "seg%d" % i
    print("%-4s inst/frame %6.1f  samples %5.1f%%  wf/frame %6.1f ideal %6.1f  " % (nm, a["inst"] / n, a["smp"] / ts * 100, a["wf"] / n, a["ideal"] / n)
          + ", ".join("%s %.1f(%.1f)" % (k, v / n, a["opsi"][k] / n) for k, v in a["ops"].most_common(6)))
