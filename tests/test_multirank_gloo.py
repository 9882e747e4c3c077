"""World-size-2 batch split on CPU (gloo): the N > 1 path of bench.py / cerberus_b200.parallel without GPUs.
Each rank solves its contiguous shard of the batch with the kernel simulator; the gathered result must equal the
single-process result bit for bit (no collective touches the data path)."""
import os
import socket
import numpy as np
import torch.multiprocessing as mp
from cerberus_b200 import abi, synth, parallel
from helpers import small_cfg, sim_backend

NW, F, ITERS = 3, 6, 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _solve_shard(lo, hi):
    from oracle_lib import OracleBackend
    cfg = small_cfg(max_features=8, iters=ITERS)
    gen = synth.generate_batch(hi - lo, F, OracleBackend(cfg), window0=lo, prior_features=4)
    sim_backend(cfg).solve_batch(gen)
    st = gen.state_array()
    return np.concatenate([st["para_Pose"].reshape(hi - lo, -1), gen.report_array()["final_cost"][:, None]], axis=1)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(NW, rank, world)
    rows = _solve_shard(lo, hi)
    allrows = parallel.gather_rows(rows, dist)
    tmax = parallel.max_over_ranks(float(rank + 1), dist)
    if rank == 0:
        q.put((allrows, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_the_batch():
    for n in (1, 2, 7, 1024):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_batch_split_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    allrows, tmax = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert tmax == 2.0
    single = _solve_shard(0, NW)
    assert allrows.shape == single.shape
    assert (allrows == single).all()
