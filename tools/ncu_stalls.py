"""Summarise an ncu report's source page for one kernel: stall reasons + hottest SASS lines.
usage: python tools/ncu_stalls.py REPORT.ncu-rep KERNEL_REGEX [launch_index] [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
# split per kernel instance
blocks, cur = [], []
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = []
    cur.append(line)
if cur: blocks.append(cur)
blk = blocks[idx]
print(blk[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
hdr, data = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return int(r[ix[k]])
    except Exception: return 0
tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, " instructions executed", sum(num(r, "Instructions Executed") for r in data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(num(r, s) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v: print(f"  {s:26s}{v:8d} {100 * v / max(tot, 1):5.1f}%")
print()
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
    st = {s: num(r, s) for s in stalls if num(r, s) > 0}
    m = max(st, key=st.get) if st else ""
    print(str(num(r, "# Samples")).rjust(7), str(num(r, "Instructions Executed")).rjust(10), r[ix["Source"]].strip()[:80].ljust(80), m)
