"""Top stall locations of an ncu report's source page: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    try: s = int(r[ix["# Samples"]])
    except ValueError: continue
    data.append((s, r[ix["Source"]], r[ix["Instructions Executed"]]))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
acc = 0
for i, (s, src, ne) in enumerate(data):
    data[i] = (s, i, src, ne)
for s, i, src, ne in sorted(data, reverse=True)[:topn]:
    print("%6d %5.1f%%  #%5d  exec=%8s  %s" % (s, 100.0 * s / tot, i, ne, src[:110]))
