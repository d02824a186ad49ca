"""Dynamic SASS opcode mix of a kernel from an .ncu-rep (source page): executed warp-instructions per opcode
and the implied fmaheavy-pipe cycles (IMAD.WIDE / IMAD.HI = 4 cycles per warp, other IMAD forms = 2)."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); si = hdr.index("Source")
mix = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= ci: continue
    src = r[si].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m: continue
    op = m.group(2)
    try: n = int(r[ci])
    except ValueError: continue
    mix[op] += n
tot = sum(mix.values())
heavy = 0
for op, n in mix.items():
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"): heavy += 4 * n
    elif op.startswith("IMAD"): heavy += 2 * n
print(f"total warp-instructions {tot:,}; modelled fmaheavy cycles {heavy:,}")
for op, n in mix.most_common(22):
    print(f"{op:28s} {n:>14,} {100.0 * n / tot:6.2f}%")
