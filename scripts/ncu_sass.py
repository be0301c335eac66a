"""Opcode histogram (weighted by executed warp-instructions) of one kernel of an .ncu-rep, plus a SASS dump with
per-instruction execution counts.  Usage: python scripts/ncu_sass.py rep.ncu-rep kernel_substring [dump.txt]"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
dump = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "sass"], capture_output=True, text=True).stdout
ops = collections.Counter()
rows = []
active = False
seen = False
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        active = kern in r[1] and not seen      # the report lists every kernel twice: keep the first listing
        seen = seen or active
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if not active or hdr is None:
        continue
    try:
        n = int(r[hdr.index("Instructions Executed")])
        samp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    sass = r[1].split()
    op = sass[1] if sass[0].startswith("@") else sass[0]
    ops[op.split(".")[0]] += n
    rows.append((n, samp, r[1].strip()))
tot = sum(ops.values())
print(f"total warp-instructions {tot}")
for op, n in ops.most_common(40):
    print(f"  {op:12s} {n:12d} {100.0 * n / tot:5.1f}%")
if dump:
    with open(dump, "w") as fh:
        for n, samp, s in rows:
            fh.write(f"{n:10d} {samp:6d}  {s}\n")
