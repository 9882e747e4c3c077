"""How does the CPU oracle scale with threads on this box? (cgroup limits / SMT / allocator contention)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cerberus_b200 import abi, synth
from oracle_lib import OracleBackend
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: pass
ob = OracleBackend()
batch = synth.generate_batch(128, 150, ob, prior_features=16)
saved = batch.copy_states()
for nt in (1, 4, 16, 32, 64, 128):
    n = min(128, max(nt, 4))
    sub = synth.tile_batch(batch, n)
    t0 = time.time(); ob.solve_batch(sub, nthreads=nt); dt = time.time() - t0
    print("threads %3d windows %3d: %.3f s -> %.1f solves/s (%.2f per thread)" % (nt, n, dt, n / dt, n / dt / nt), flush=True)
