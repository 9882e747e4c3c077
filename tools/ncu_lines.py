"""Aggregate an `ncu --page source --csv` SASS dump of vilo_solve_kernel per source line.
usage: python tools/ncu_lines.py <ncu_sass.csv> <lib.so> [top]
Joins instruction order with `nvdisasm -g` line info of the same library build (instruction i of the kernel in both)."""
import csv, re, subprocess, sys, tempfile, os, collections
csv_path, lib, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
kernel = next((a[7:] for a in sys.argv[4:] if a.startswith("kernel=")), "vilo_solve_kernel")      # e.g. kernel=marg_schur_kernel
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith("//--------------------- .text._ZN4cerb") and kernel in l)
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
    srcl = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cerberus_b200", "csrc", "solve_kernel.cuh")).read().splitlines()
    def at(pat):
        return next(i + 1 for i, l in enumerate(srcl) if pat in l)
    cuts = [("vision_cost", at("double vision_cost(")), ("vis setup", at("double vision_linearize(")), ("vis eval+tile", at("PH_MARK(20)")), ("vis DMMA", at("PH_MARK(21)")),
            ("vis reduce", at("PH_MARK(22)")), ("vis tail", at("PH_MARK(23)")), ("imu lin", at("CERB_D void imu_lin_all")), ("inertial cost / prior res", at("double inertial_cost(")),
            ("scatter_H", at("CERB_D int imu_col_dest")), ("imu whiten+gram", at("double inertial_linearize(")), ("prior", at("PH_MARK(26)")), ("plus/ambient", at("CERB_D void apply_plus")),
            ("setup", at("CERB_GLOBAL void")), ("post-lin", at("PH_MARK(0)")), ("cauchy", at("PH_MARK(3)")), ("hyy chain", at("PH_MARK(4)")), ("schur dmma", at("PH_MARK(6)")),
            ("T solve", at("PH_MARK_T(7")), ("TT^T", at("PH_MARK(8)")), ("dense chol", at("PH_MARK(9)")), ("backsub", at("PH_MARK(10)")), ("y part", at("PH_MARK(11)")),
            ("inv depth", at("PH_MARK(12)")), ("dogleg/accept", at("PH_MARK(13)")), ("end", len(srcl) + 1)]
    ranges = [(cuts[i][1], cuts[i + 1][1] - 1, cuts[i][0]) for i in range(len(cuts) - 1)]
    acc = collections.Counter(); bar = collections.Counter()
    for ln, a in agg.items():
        f, n = ln if ln else ("?", 0)
        key = f
        if f == "solve_kernel.cuh":
            key = next((nm for lo, hi, nm in ranges if lo <= n <= hi), f"solve:{n}")
        acc[key] += a["samples"]; bar[key] += a["stall_barrier"]
    for k, v in acc.most_common(40): print(f"  {100 * v / tot:5.1f}%  (barrier {100 * bar[k] / tot:4.1f}%)  {k}")
