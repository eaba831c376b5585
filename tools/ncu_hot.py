"""Summarise an `ncu --page source --csv` dump: hottest instruction windows by warp-stall samples."""
import csv
import sys

path = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 48
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
# keep the first kernel instance only
first, seen = [], set()
for r in data:
    if r[0] in seen:
        break
    seen.add(r[0]); first.append(r)
data = first
S = ix["# Samples"]
tot = sum(int(r[S]) for r in data)
print("instructions", len(data), "total samples", tot)
acc = []
for i in range(0, len(data), win):
    chunk = data[i:i + win]
    acc.append((sum(int(r[S]) for r in chunk), i, chunk))
for s, i, chunk in sorted(acc, key=lambda x: -x[0])[:16]:
    ops = {}
    for r in chunk:
        sp = r[ix["Source"]].split()
        op = sp[0] if sp else ""
        if op.startswith("@") and len(sp) > 1:
            op = sp[1]
        ops[op] = ops.get(op, 0) + int(r[S])
    top = sorted(ops.items(), key=lambda x: -x[1])[:6]
    print(f"instr {i:5d}: {s:6d} samples ({100 * s / max(tot, 1):.1f}%)", top)
