"""Top stall sites of a kernel from an ncu report's source page (SASS view): address, samples, dominant stall reasons.
usage: python scripts/ncu_hot.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
tot = sum(int(r[ix["# Samples"]]) for r in data)
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("total samples", tot, " by reason:", ", ".join(f"{h[6:]} {v / tot * 100:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
exc = sum(int(r[ix["L1 Wavefronts Shared Excessive"]] or 0) for r in data); wf = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in data)
print("shared wavefronts", wf, "excessive", exc)
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    s = int(r[ix["# Samples"]])
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{r[0]} {s / tot * 100:5.2f}%  exec {r[ix['Instructions Executed']]:>9s}  wf {r[ix['L1 Wavefronts Shared']]:>8s}/{r[ix['L1 Wavefronts Shared Ideal']]:>8s}  {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}  {r[1].strip()[:90]}")
