"""BASELINE.json configs[4]: 64k synthetic 10-frame x 2000-feature windows sharded across the ranks of one node (one rank per GPU, contiguous
batch split, no data-path collective), each rank working through its share in sub-batches of 1024 windows (host buffers -> cerb_solve_batch).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 tools/stress_64k.py [--windows 65536] [--features 2000]

Prints one JSON line on rank 0: total solves/s over the whole job (max over ranks of the wall time of the share, barrier on both sides), the
device-side rate, bytes moved, and the algorithmic HBM rate (B_alg(F) = 114160 + 816 F bytes per solve)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=65536); ap.add_argument("--features", type=int, default=2000)
    ap.add_argument("--sub", type=int, default=1024); ap.add_argument("--distinct", type=int, default=32)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cerberus_b200 import abi, synth, lib, parallel
    lo, hi = parallel.shard_range(a.windows, rank, world)
    share = hi - lo
    F = a.features
    cfg = abi.default_config(); cfg.device = local
    cfg.max_batch, cfg.max_features, cfg.max_obs = a.sub, ((F + 7) // 8) * 8 + 8, (((F + 7) // 8) * 8 + 8) * abi.NUM_FRAMES
    be = lib.Backend(cfg)
    base = synth.generate_batch(min(a.distinct, a.sub), F, be, cfg=cfg, window0=lo, prior_features=24)
    batch = synth.tile_batch(base, a.sub)
    saved = batch.copy_states()
    regs = be.register_batch(batch)
    be.solve_batch(batch)                      # warm-up
    def barrier():
        if world > 1: dist.barrier()
        be.sync()
    barrier()
    t0 = time.perf_counter(); dev_ms = 0.0; done = 0; iters = []
    while done < share:
        n = min(a.sub, share - done)
        batch.restore_states(saved)
        if n < a.sub: batch.n = n
        rep = be.solve_batch(batch)
        ms, _ = be.last_solve_stats(); dev_ms += ms; done += n; iters.append(float(rep["iterations"][:n].mean()))
    wall = time.perf_counter() - t0
    barrier()
    if world > 1:
        t = torch.tensor([wall, dev_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); wall, dev_ms = t.tolist()
    if rank == 0:
        b_alg = 114160 + 816 * F
        print(json.dumps({"workload": f"{a.windows} windows x {F} features over {world} GPU(s), sub-batches of {a.sub} through cerb_solve_batch (registered host buffers)",
                          "solves_per_s_end_to_end": a.windows / wall, "solves_per_s_device": a.windows / (dev_ms * 1e-3), "wall_s": wall, "device_s": dev_ms * 1e-3,
                          "windows_per_gpu": share, "mean_iterations": float(np.mean(iters)), "algorithmic_GB_per_s_per_gpu": b_alg * share / (dev_ms * 1e-3) / 1e9,
                          "h2d_GB_total": a.windows * (F * 11 * 80 + F * 24 + 10 * 10432 + 86 * 86 * 8 + 3500) / 1e9}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
