"""Summarises `ncu -i X.ncu-rep --page source --csv` output: stall-reason totals and the hottest SASS instructions.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_source_summary.py src.csv [ntop]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
name = rows[0][1]
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_")]
body = rows[2:]
tot = {s: 0 for s in stalls}
samples = 0
for r in body:
    samples += int(r[idx["# Samples"]] or 0)
    for s in stalls:
        tot[s] += int(r[idx[s]] or 0)
print(f"kernel: {name}")
print(f"instructions: {len(body)}, total warp-stall samples: {samples}")
for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {s:24s} {v:9d} {100.0 * v / max(samples, 1):5.1f}%")
exc = sum(int(r[idx["L1 Wavefronts Shared Excessive"]] or 0) for r in body)
wav = sum(int(r[idx["L1 Wavefronts Shared"]] or 0) for r in body)
print(f"shared wavefronts {wav}, excessive {exc}")
print("top instructions (index, samples, top stall reasons, SASS):")
order = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:ntop]
for i in order:
    r = body[i]
    st = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:3]
    print(f"  {i:5d} {int(r[idx['# Samples']]):7d}  " + " ".join(f"{s[6:]}={v}" for v, s in st if v) + "   " + r[idx["Source"]].strip()
          + (f"   [smem excessive {r[idx['L1 Wavefronts Shared Excessive']]}]" if int(r[idx['L1 Wavefronts Shared Excessive']] or 0) else ""))
