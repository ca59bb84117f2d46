#!/usr/bin/env python
"""Digest `ncu -i X.ncu-rep --page source --csv`: stall-reason totals, top instructions by samples, shared-memory
excess wavefronts, and executed-instruction mix.  Usage: ncu_source.py src.csv [topN]"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
num = lambda r, h: float(r[ix[h]]) if r[ix[h]] not in ("", "-") else 0.0
tot_samples = sum(num(r, "# Samples") for r in body)
tot_inst = sum(num(r, "Instructions Executed") for r in body)
print(f"kernel: {rows[0][1]}")
print(f"instructions executed (warp-level): {tot_inst:.0f}   samples: {tot_samples:.0f}")
stalls = [h for h in hdr if h.startswith("stall_")]
st = {h: sum(num(r, h) for r in body) for h in stalls}
print("stall samples (all):", ", ".join(f"{k[6:]}={v / tot_samples * 100:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0))
mix = collections.Counter()
for r in body:
    op = r[ix["Source"]].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    mix[op.split(".")[0]] += num(r, "Instructions Executed")
print("executed mix:", ", ".join(f"{k}={v / tot_inst * 100:.1f}%" for k, v in mix.most_common(14)))
print(f"\ntop {top} instructions by stall samples:")
for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:top]:
    reasons = sorted(((num(r, h), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"  {num(r, '# Samples') / tot_samples * 100:5.2f}%  exec={num(r, 'Instructions Executed'):10.0f}  {r[ix['Source']].strip():70s} {reasons[0][1]}:{reasons[0][0]:.0f} {reasons[1][1]}:{reasons[1][0]:.0f}")
ex = [r for r in body if num(r, "L1 Wavefronts Shared Excessive") > 0]
if ex:
    print("\nshared-memory instructions with excess wavefronts:")
    for r in sorted(ex, key=lambda r: -num(r, "L1 Wavefronts Shared Excessive"))[:12]:
        print(f"  excess={num(r, 'L1 Wavefronts Shared Excessive'):10.0f} of {num(r, 'L1 Wavefronts Shared'):10.0f} (ideal {num(r, 'L1 Wavefronts Shared Ideal'):10.0f}) nway={r[ix['L1 Conflicts Shared N-Way']]}  {r[ix['Source']].strip()}")
