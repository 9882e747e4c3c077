"""Wall-clock timing of cerb_marginalize_schur on the GPU (host buffers in / out, so the H2D of A and the D2H of the factor are
included) for the sizes of BASELINE.json's configurations.  Not part of bench.py: the marginalization is outside the timed step.

    python tools/marg_bench.py [--windows 256] [--reps 3]
"""
import argparse
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerberus_b200 import abi, lib      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", type=int, default=-1, help="index of the single size to run (profiling)")
    a = ap.parse_args()
    cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs = 2, 8, 88
    be = lib.Backend(cfg)
    rng = np.random.default_rng(0)
    sizes = ((6, 80), (19 + 16, 86), (19 + 50, 86), (19 + 150, 86))
    for m, n in (sizes if a.only < 0 else sizes[a.only:a.only + 1]):
        pos = m + n
        J = rng.standard_normal((a.windows, 3 * pos, pos)); J[:, :, :m] *= np.exp(rng.uniform(-2, 2, (a.windows, 1, m)))
        A = np.swapaxes(J, 1, 2) @ J; b = (np.swapaxes(J, 1, 2) @ rng.standard_normal((a.windows, 3 * pos, 1)))[..., 0]
        be.marginalize_schur(A[:2], b[:2], m)                       # warm-up
        best, sw = 1e9, None
        for _ in range(a.reps):
            t = time.perf_counter(); _, _, sw = be.marginalize_schur(A, b, m, return_sweeps=True); best = min(best, time.perf_counter() - t)
        print(f"m={m:4d} n={n:3d} windows={a.windows}: {best * 1e3:8.2f} ms  ({best / a.windows * 1e6:8.1f} us / window, "
              f"{A.nbytes / best / 1e9:5.2f} GB/s of A),  sweeps {sw[:, 0].mean():.1f} / {sw[:, 1].mean():.1f}")


if __name__ == "__main__":
    main()
