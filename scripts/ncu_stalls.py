"""Per-instruction stall attribution from an ncu report's source page (needs --import-source on / -lineinfo).
  python scripts/ncu_stalls.py <report.ncu-rep> <kernel regex> [min samples]
"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 150
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows) + 1]
h = hi[0]
hdr = rows[h]
iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[h + 1:hi[1] - 1] if len(r) == len(hdr)]
tot = sum(int(r[iN] or 0) for r in data)
print("kernel:", rows[h - 1][1][:100] if h else "")
print("total samples", tot, "instructions", len(data))
agg = {c: sum(int(r[hdr.index(c)] or 0) for r in data) for c in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in enumerate(data):
    n = int(r[iN] or 0)
    if n > thr:
        st = {c: int(r[hdr.index(c)] or 0) for c in stalls}
        top = max(st, key=st.get)
        print(idx, r[0][-5:], n, r[iX], top, st[top], r[iS][:90])
