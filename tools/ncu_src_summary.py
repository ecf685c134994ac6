"""Summarise an `ncu --page source --csv --print-source sass` export: sample share per SASS region and top stalls."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
end = len(rows)
for i in range(hi + 1, len(rows)):
    if rows[i] and rows[i][0] == "Kernel Name":
        end = i
        break
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ci = hdr.index("# Samples")
si = hdr.index("Source")
ie = hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci]) for r in data)
print("total samples", tot, "instructions", len(data))
KEYS = ['MUFU', 'LDTM', 'STTM', 'UTMA', 'UTCHMMA', 'BAR', 'SYNCS', 'UTCBAR', 'FFMA2', 'FMUL2', 'MEMBAR', 'LDG', 'STG', 'LDS', 'STS']
for b in range(0, len(data), bucket):
    seg = data[b:b + bucket]
    s = sum(int(r[ci]) for r in seg)
    ex = max(int(r[ie]) for r in seg)
    ops = set((r[si].split()[1] if r[si].startswith('@') else r[si].split()[0]) for r in seg if r[si].split())
    key = sorted(o for o in ops if any(k in o for k in KEYS))
    print(f"{b:5d} {s:6d} {100 * s / tot:5.1f}%  ex={ex:9d} {key}")
for r in sorted(data, key=lambda r: -int(r[ci]))[:16]:
    st = {h: int(r[hdr.index(h)]) for h in stalls if int(r[hdr.index(h)]) > 0}
    st = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{int(r[ci]):6d} {100 * int(r[ci]) / tot:5.1f}%  idx={data.index(r):5d} ex={r[ie]:>8s} {r[si][:64]:64s} {st}")
