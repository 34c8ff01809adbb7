"""Summarise an `ncu --page source --csv` dump: top SASS lines by stall samples, and the stall mix."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
si, src = ix['# Samples'], ix['Source']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
tot_st = {s: 0 for s in stalls}
for r in rows[2:]:
    try:
        n = int(r[si])
    except Exception:
        continue
    data.append((n, r))
    for s in stalls:
        try: tot_st[s] += int(r[ix[s]])
        except Exception: pass
tot = sum(n for n, _ in data) or 1
print("total samples", tot)
print("stall mix:", ", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for n, r in sorted(data, key=lambda t: -t[0])[:topn]:
    top = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    ws = r[ix['L1 Wavefronts Shared']]
    print(f"{n:8d} {100*n/tot:5.1f}%  {r[src][:80]:80s} {top} wf={ws}")
