"""Aggregate an `ncu --page source --csv` SASS dump of vilo_solve_kernel per source line.
usage: python tools/ncu_lines.py <ncu_sass.csv> <lib.so> [top]
Joins instruction order with `nvdisasm -g` line info of the same library build (instruction i of the kernel in both)."""
import csv, re, subprocess, sys, tempfile, os, collections
csv_path, lib, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if ".text._ZN4cerb17vilo_solve_kernel" in l)
lines, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]; body = rows[2:]
ci = {n: i for i, n in enumerate(hdr)}
assert len(body) == len(lines), (len(body), len(lines))
agg = collections.defaultdict(lambda: collections.Counter())
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
for r, ln in zip(body, lines):
    a = agg[ln]; n = int(r[ci["# Samples"]]); a["samples"] += n; tot += n; a["inst"] += int(r[ci["Instructions Executed"]])
    for st in stalls: a[st] += int(r[ci[st]])
    a["conf"] += int(r[ci["L1 Wavefronts Shared Excessive"]] or 0)
print(f"total samples {tot}")
allst = collections.Counter()
for a in agg.values():
    for st in stalls: allst[st] += a[st]
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in allst.most_common(9)))
src = {}
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    f, n = ln if ln else ("?", 0)
    if f not in src:
        pth = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cerberus_b200", "csrc", f)
        src[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    main = ", ".join(f"{k[6:]} {v}" for k, v in collections.Counter({st: a[st] for st in stalls}).most_common(3))
    print(f"{100 * a['samples'] / tot:5.1f}%  inst {a['inst']:9d}  bank-excess {a['conf']:9d}  {f}:{n:<4d} [{main}]  {text}")

# ---- optional: share of samples per line range of solve_kernel.cuh (rough phase view) ----
if len(sys.argv) > 4:
    ranges = [(158, 189, "vis setup"), (190, 223, "vis eval+tile"), (224, 258, "vis DMMA"), (259, 277, "vis reduce"), (278, 303, "vis tail"), (306, 316, "imu lin"), (317, 353, "inertial cost / prior res"),
              (354, 377, "scatter_H"), (378, 447, "imu whiten+gram"), (448, 490, "prior"), (491, 519, "plus/ambient"), (520, 557, "setup"), (558, 616, "post-lin"), (617, 655, "cauchy"),
              (656, 720, "hyy chain"), (721, 776, "schur dmma"), (777, 795, "T solve"), (796, 831, "TT^T"), (832, 915, "dense chol"), (916, 942, "backsub"), (943, 964, "y part"), (965, 977, "inv depth"), (978, 1101, "dogleg/accept")]
    acc = collections.Counter(); bar = collections.Counter()
    for ln, a in agg.items():
        f, n = ln if ln else ("?", 0)
        key = f
        if f == "solve_kernel.cuh":
            key = next((nm for lo, hi, nm in ranges if lo <= n <= hi), f"solve:{n}")
        acc[key] += a["samples"]; bar[key] += a["stall_barrier"]
    for k, v in acc.most_common(40): print(f"  {100 * v / tot:5.1f}%  (barrier {100 * bar[k] / tot:4.1f}%)  {k}")
