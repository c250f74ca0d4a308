"""Top stall sites of a kernel from an ncu report's source page (SASS view): python tools/ncu_hot.py report.ncu-rep [kernel-index] [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
b = blocks[kidx]
print(b[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ix["# Samples"]])
    except: continue
    tot += n
    st = {h: int(r[ix[h]] or 0) for h in stall_cols}
    data.append((n, r[ix["Address"]], r[ix["Source"]], int(r[ix["Instructions Executed"]] or 0), st))
print("total samples", tot)
agg = collections.Counter()
for n, a, s, ie, st in data:
    for h, v in st.items(): agg[h] += v
print("stall mix:", ", ".join(f"{h[6:]}={v}" for h, v in agg.most_common(10)))
for n, a, s, ie, st in sorted(data, key=lambda t: -t[0])[:top]:
    top2 = ", ".join(f"{h[6:]}={v}" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{n:7d} {100.0*n/tot:5.1f}%  exec={ie:9d}  {s[:90]:90s} | {top2}")
