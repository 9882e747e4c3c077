"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py     (shared-memory hazards: mbarrier-guarded bulk copies, named barriers)

6 windows x 20 features (two pipeline waves on a grid capped at 4 CTAs would need > 8 windows; the chunked path is covered by tests), prior,
2 iterations; host-buffer solve (pack / prepare / solve / unpack), per-feature passes, marginalization (both modes), preintegration, evaluators."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cerberus_b200 import abi, synth, lib

cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs, cfg.max_num_iterations = 8, 24, 24 * 11, int(os.environ.get("ITERS", "2"))
be = lib.Backend(cfg)
batch = synth.generate_batch(6, 20, be, cfg=cfg, prior_features=8, window0=5)       # preintegrate_kernel, evaluators, marg kernels (the prior)
regs = be.register_batch(batch)
rep = be.solve_batch(batch)
be.unregister(regs)
print("solve", rep["iterations"], rep["final_cost"][:2])
err, rem = be.outlier_errors(batch.n); dep = be.triangulate(batch.n); sh = be.shift_depth(batch.n)
flags = np.array([0, 1, 0, 1, 0, 0], dtype=np.int32)
J = np.zeros((6, abi.MAX_PRIOR_DIM ** 2)); r = np.zeros((6, abi.MAX_PRIOR_DIM)); priors = (abi.Prior * 6)()
for w in range(6):
    priors[w].linearized_jacobians = J[w].ctypes.data_as(abi.c_dp); priors[w].linearized_residuals = r[w].ctypes.data_as(abi.c_dp)
sw = be.batch_marginalize(flags, None, priors)
print("marginalize", [priors[w].n for w in range(6)], sw.ravel())
cost, g, d = be.debug_linearize(batch, 1)
print("probe", cost)
# the eigen-Schur kernel in its other memory plans: T split between shared memory and the L2 workspace (m = 169, one CTA of 512 threads per SM),
# everything in the workspace (forced on a small matrix through the plan's test hook)
rng = np.random.default_rng(3)
for m, n, hook in ((169, 86, None), (30, 20, "4096")):
    if hook: os.environ["CERB_TEST_MARG_SMEM"] = hook
    pos = m + n
    Jf = rng.standard_normal((2, 2 * pos, pos)); A = np.swapaxes(Jf, 1, 2) @ Jf; b = (np.swapaxes(Jf, 1, 2) @ rng.standard_normal((2, 2 * pos, 1)))[..., 0]
    Jo, ro, sw2 = be.marginalize_schur(A, b, m, return_sweeps=True)
    H = np.swapaxes(Jo, 1, 2) @ Jo
    Ai = np.linalg.inv(A[:, :m, :m]); Hr = A[:, m:, m:] - A[:, m:, :m] @ Ai @ A[:, :m, m:]
    print("marginalize_schur", m, n, hook, sw2.ravel(), float(np.abs(H - Hr).max() / np.abs(Hr).max()))
    os.environ.pop("CERB_TEST_MARG_SMEM", None)
