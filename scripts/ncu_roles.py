"""Per-role (producer / halo / consumer) instruction and stall breakdown of k3f_pass1/k3f_pass2 from an .ncu-rep.
Roles are told apart by walking the SASS in address order: instructions inlined from other files inherit the
role of the closest preceding instruction that maps to a line of rmi3_fast.cuh."""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
src = open("seghiero_b200/csrc/rmi3_fast.cuh").read().split("\n")
# role boundaries from marker comments inside the kernel
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"// =+ (producers|halo warp|consumers) =+", l)
    if m:
        marks.append((i, m.group(1)))
helpers = {}
for i, l in enumerate(src, 1):
    m = re.match(r"__device__ __forceinline__ void (pp_taps|lp_taps|load_row8|stencil\w*|lpgrad\w*)", l)
    if m:
        helpers[i] = m.group(1)


def role_of(line, kstart):
    r = "setup"
    for ln, name in marks:
        if ln >= kstart and line >= ln:
            r = name
    return r


out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
amap = {}
fn = cur = None
line = 0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        pass
    elif r[0] != "":
        try:
            line = int(r[0])
        except ValueError:
            pass
    elif len(r) > 3 and r[2].startswith("0x") and fn and kern in fn:
        amap[r[2]] = (cur, line)
kstart = next(i for i, l in enumerate(src, 1) if l.startswith(kern + "(") or (kern + "(") in l and "__global__" in src[i - 2] + src[i - 3])
out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "sass"], capture_output=True, text=True).stdout
fn = hdr = None
seen = set()
role = "setup"
inst = collections.Counter()
stall = collections.defaultdict(collections.Counter)
ops = collections.defaultdict(collections.Counter)
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        fn = r[1]
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if not (fn and kern in fn and r[0].startswith("0x")) or r[0] in seen:
        continue
    seen.add(r[0])
    f, ln = amap.get(r[0], ("?", 0))
    if f == "rmi3_fast.cuh":
        if ln >= kstart:
            role = role_of(ln, kstart)
        # helper bodies above the kernel: keep the current role (they are inlined where they are called)
    n = int(r[hdr.index("Instructions Executed")])
    inst[role] += n
    sass = r[1].split()
    op = (sass[1] if sass and sass[0].startswith("@") and len(sass) > 1 else (sass[0] if sass else "")).split(".")[0]
    ops[role][op] += n
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                stall[role][h[6:]] += int(r[i])
            except ValueError:
                pass
elems = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
for role in inst:
    tot = sum(stall[role].values()) or 1
    print(f"{role:10s} {inst[role] / 1e6:8.1f}M warp-inst  ({inst[role] * 32 / elems:.1f}/elem)  samples {tot}")
    print("     stalls: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in stall[role].most_common(7)))
    print("     ops: " + ", ".join(f"{k} {v * 32 / elems:.1f}" for k, v in ops[role].most_common(16)))
