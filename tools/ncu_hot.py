"""Print the hottest SASS instructions (warp-stall samples) of a kernel from an .ncu-rep, with context.
usage: python tools/ncu_hot.py report.ncu-rep [top] [context]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot)
order = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:top]
for i in order:
    r = data[i]
    s = sorted(((k, f(r, k)) for k in stalls), key=lambda kv: -kv[1])[:2]
    print(f"--- {100 * f(r, '# Samples') / tot:5.1f}%  line {i}  {s}")
    for j in range(max(0, i - ctx), min(len(data), i + ctx + 1)):
        rr = data[j]
        print(f"   {'>>' if j == i else '  '} {f(rr, '# Samples'):7.0f} x{f(rr, 'Instructions Executed'):9.0f}  {rr[ix['Source']][:100]}")
