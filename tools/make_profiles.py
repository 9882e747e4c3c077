"""Turn the outputs of tools/gpu_round_end.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, Pf = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_final"

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); return {h: (rows[2][i], rows[1][i]) for i, h in enumerate(rows[0])}

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__ops_path_tensor_src_fp64.sum', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'sm__cycles_active.avg']
def b(v): return float(v[0]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[v[1]]
for nw in (1024, 148):
    a = raw(os.path.join(G, f"solve_final_{nw}.ncu-rep"))
    rd, wr = b(a['dram__bytes_read.sum']), b(a['dram__bytes_write.sum']); alg = (114160 + 816 * 150) * nw
    with open(os.path.join(Pf, f"solve_kernel_{tag}_{nw}.txt"), "w") as f:
        f.write(f"# ncu --set full summary, vilo_solve_kernel, {nw} windows x 150 features, B200 ({tag})\n# command: ncu --set full --clock-control none --import-source on -k regex:vilo_solve -c 1 python tools/profile_solve.py {nw} 150 1\n\n")
        for k in WANT: f.write(f"{k:92s} {a[k][0]} {a[k][1]}\n")
        f.write(f"\nderived:\n  DRAM traffic per launch      {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB write = {(rd + wr) / 1e6:.1f} MB ({(rd + wr) / nw / 1e6:.2f} MB per window)\n"
                f"  algorithmic bytes per launch {alg / 1e6:.1f} MB (237 KB per window) -> traffic / algorithmic = {(rd + wr) / alg:.1f}x\n")
    if nw == 1024:
        json.dump({"kernel": "vilo_solve_kernel", "windows": 1024, "features": 150, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                   "duration_s_under_ncu": float(a['gpu__time_duration.sum'][0]) * 1e-3, "source": f"profiles/solve_kernel_{tag}_1024.txt (ncu --set full)"},
                  open(os.path.join(Pf, "solve_kernel_traffic.json"), "w"), indent=1)
    print(nw, a['gpu__time_duration.sum'], f"traffic {(rd + wr) / 1e6:.0f} MB", a['sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'], a['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'], a['smsp__issue_active.avg.pct_of_peak_sustained_active'], a['sm__cycles_active.avg'])
# per-line / per-phase aggregation of the one-wave capture
sass = os.path.join(G, "solve_final_148_sass.csv")
open(sass, "w").write(subprocess.run(["ncu", "-i", os.path.join(G, "solve_final_148.ncu-rep"), "--page", "source", "--csv"], capture_output=True, text=True).stdout)
open(os.path.join(Pf, f"solve_kernel_{tag}_lines.txt"), "w").write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), sass, os.path.join(ROOT, "cerberus_b200", "libcerberus_b200.so"), "40", "ranges"], capture_output=True, text=True).stdout)
# launch list
rows = [r for r in csv.reader(open(os.path.join(G, "launches_final.csv"))) if len(r) > 5]
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]; H = rows[hi]; ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    v *= {'nsecond': 1e-3, 'ns': 1e-3, 'usecond': 1.0, 'us': 1.0, 'msecond': 1e3, 'ms': 1e3}.get(r[ui], 1.0)
    n = r[ki].split('(')[0].split('::')[-1]; tot[n] += v; cnt[n] += 1
T = sum(tot.values())
with open(os.path.join(Pf, f"launches_{tag}_summary.csv"), "w") as f:
    f.write(f"# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` ({tag}, B200); gpu__time_duration.sum per kernel,\n# cold-cache / serialised: compare SHARES, not absolutes.  preintegrate / *_eval kernels belong to the synthetic-data set-up.\nkernel,launches,total_us,share\n")
    for k, v in tot.most_common(): f.write(f"{k},{cnt[k]},{v:.1f},{v / T:.4f}\n")
step = {k: v for k, v in tot.items() if k in ("vilo_solve_kernel", "prior_prepare_kernel", "imu_leg_prepare_kernel", "pack_kernel", "unpack_kernel")}
print({k: round(v / sum(step.values()), 4) for k, v in step.items()})
shutil.copy(os.path.join(G, "launches_final.csv"), os.path.join(Pf, f"launches_{tag}.csv"))
shutil.copy(os.path.join(G, "phase_final.txt"), os.path.join(Pf, f"phase_breakdown_{tag}.txt"))
shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(Pf, f"bench_{tag}.json"))
shutil.copy(os.path.join(G, "bench_reference.json"), os.path.join(Pf, f"bench_{tag}_reference.json"))
d = json.load(open(os.path.join(G, "bench_final.json")))
print("bench", round(d["value"]), "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline"))

# ---- round 2: marg_schur_kernel capture, auxiliary kernels, end-to-end trace, replay -------------------------------------------
mrep = os.path.join(G, "marg_schur_final.ncu-rep")
if os.path.exists(mrep):
    a = raw(mrep)
    with open(os.path.join(Pf, f"marg_schur_kernel_{tag}.txt"), "w") as f:
        f.write(f"# ncu --set full summary, marg_schur_kernel, 296 windows of the 150-feature size (m = 169, n = 86; the cerb_marginalize_schur launch of tools/aux_kernels.py), B200 ({tag})\n"
                f"# command: ncu --set full --clock-control none --import-source on -k regex:marg_schur -s 1 -c 1 python tools/aux_kernels.py 1024 150\n\n")
        for k in WANT:
            if k in a: f.write(f"{k:92s} {a[k][0]} {a[k][1]}\n")
aux = os.path.join(G, "aux_launches_final.csv")
if os.path.exists(aux):
    rows = [r for r in csv.reader(open(aux)) if len(r) > 10]
    H = rows[0]; ik, im, iv, iid = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((r[iid], r[ik].split('(')[0]), {})[r[im]] = float(r[iv].replace(',', ''))
    with open(os.path.join(Pf, f"aux_kernels_{tag}.csv"), "w") as f:
        f.write(f"# kernels outside the timed solve step at benchmark scale (python tools/aux_kernels.py 1024 150 under ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum), {tag}\n"
                "# measured copy peak of this pool: 6566 GB/s (MEASURED_PEAKS.json)\nid,kernel,duration_us,dram_MB,dram_GBps,frac_of_hbm_peak\n")
        for (i, k), m in d.items():
            t = m.get('gpu__time_duration.sum', 0.0); b = m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
            f.write(f"{i},{k},{t / 1e3:.1f},{b / 1e6:.2f},{(b / t if t else 0):.1f},{(b / t / 6566.1 if t else 0):.4f}\n")
for src, dst in (("aux_wall_final.txt", f"aux_wall_{tag}.txt"), ("marg_bench_final.txt", f"marg_bench_{tag}.txt"), ("replay_gpu.txt", f"replay_gpu_{tag}.txt"),
                 ("marg_phase_final.txt", f"marg_phase_{tag}.txt"), ("marg_kernel_times_final.txt", f"marg_kernel_times_{tag}.txt"), ("replay_native_gpu.txt", f"replay_native_gpu_{tag}.txt")):
    if os.path.exists(os.path.join(G, src)): shutil.copy(os.path.join(G, src), os.path.join(Pf, dst))
err = os.path.join(G, "bench_final.err")
if os.path.exists(err):
    lines = [l for l in open(err) if "cerb_solve_batch" in l]
    open(os.path.join(Pf, f"e2e_trace_{tag}.txt"), "w").write("# CERB_TRACE=1 python bench.py: host timeline of every cerb_solve_batch call of the end-to-end leg (registered host buffers)\n" + "".join(lines))
