"""cerb_marginalize_schur: the dense tail of MarginalizationInfo::marginalize() (marginalization_factor.cpp:281-305) on the device
(parallel Jacobi eigen-solver + pseudo-inverse Schur complement) against the CPU restatement and against LAPACK.

The factor (linearized_jacobians, linearized_residuals) is unique only up to the order / sign of the eigenvectors, so the
comparison is on what MarginalizationFactor::Evaluate consumes: H = J^T J and g = J^T r."""
import numpy as np
import pytest
from cerberus_b200 import abi, lib
from oracle_lib import OracleBackend
from helpers import sim_backend

EPS = 1e-8


def _problem(rng, B, m, n, rank=None, scale=1.0):
    pos = m + n
    R = rank if rank is not None else 3 * pos
    Jf = rng.standard_normal((B, R, pos)) * scale
    Jf[:, :, :m] *= np.exp(rng.uniform(-2, 2, (B, 1, m)))          # inverse-depth-like spread of the dropped columns
    r = rng.standard_normal((B, R))
    Jt = np.swapaxes(Jf, 1, 2)
    return Jt @ Jf, (Jt @ r[..., None])[..., 0]


def _lapack(A, b, m):
    Amm = 0.5 * (A[:, :m, :m] + np.swapaxes(A[:, :m, :m], 1, 2))
    ev, V = np.linalg.eigh(Amm)
    inv = np.where(ev > EPS, 1.0 / np.where(ev > EPS, ev, 1.0), 0.0)
    Ai = (V * inv[:, None, :]) @ np.swapaxes(V, 1, 2)
    Ar = A[:, m:, m:] - A[:, m:, :m] @ Ai @ A[:, :m, m:]
    br = b[:, m:] - (A[:, m:, :m] @ Ai @ b[:, :m, None])[..., 0]
    L = np.tril(Ar); Ar = L + np.swapaxes(np.tril(Ar, -1), 1, 2)
    ev2, V2 = np.linalg.eigh(Ar)
    keep = ev2 > EPS
    H = (V2 * np.where(keep, ev2, 0.0)[:, None, :]) @ np.swapaxes(V2, 1, 2)
    g = (V2 @ (np.where(keep, 1.0, 0.0) * (np.swapaxes(V2, 1, 2) @ br[..., None])[..., 0])[..., None])[..., 0]
    return H, g


def _canon(J, r):
    return np.swapaxes(J, 1, 2) @ J, (np.swapaxes(J, 1, 2) @ r[..., None])[..., 0]


def _check(be, oracle, rng, B, m, n, rank=None, n_oracle=2, tol=1e-9):
    A, b = _problem(rng, B, m, n, rank)
    J, r, sw = be.marginalize_schur(A, b, m, EPS, return_sweeps=True)
    assert np.isfinite(J).all() and np.isfinite(r).all() and (sw >= 0).all() and (sw < 25).all(), sw.max()
    H, g = _canon(J, r)
    Hl, gl = _lapack(A, b, m)
    assert np.abs(H - Hl).max() < tol * np.abs(Hl).max() and np.abs(g - gl).max() < tol * max(1.0, np.abs(gl).max())
    Jo, ro = oracle.marginalize_schur(A[:n_oracle], b[:n_oracle], m, EPS)
    Ho, go = _canon(Jo, ro)
    assert np.abs(H[:n_oracle] - Ho).max() < tol * np.abs(Ho).max() and np.abs(g[:n_oracle] - go).max() < tol * max(1.0, np.abs(go).max())
    # J J^T is diagonal (rows are scaled orthonormal eigenvectors), with the kept eigenvalues on it
    G = J @ np.swapaxes(J, 1, 2)
    off = G - np.einsum("bii->bi", G)[:, :, None] * np.eye(n)
    assert np.abs(off).max() < 1e-9 * np.abs(G).max()


@pytest.mark.parametrize("m,n,rank", [(13, 20, None), (7, 9, None), (12, 10, 15), (1, 6, None), (10, 30, 14)])
def test_marg_schur_sim(m, n, rank):
    """full-rank, odd sizes (bye in the round-robin), rank-deficient Amm and Schur complement (eps clamps), a single dropped coordinate"""
    cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs = 2, 8, 8 * 11
    _check(sim_backend(cfg), OracleBackend(cfg), np.random.default_rng(m * 100 + n), 3, m, n, rank)


@pytest.mark.parametrize("limit,m,n,rank", [(8192, 30, 20, None), (8192, 31, 20, 40), (4096, 30, 20, None), (4096, 33, 12, 40)])
def test_marg_schur_memory_plans_sim(monkeypatch, limit, m, n, rank):
    """the layouts a large dropped block falls into, forced on small matrices by shrinking the shared-memory budget of the plan (test hook):
    8192 B -> M1 in shared memory, T split between shared memory and the global workspace (what m = 169 gets on the device);
    4096 B -> M1 and T in the global workspace (m > ~235)"""
    monkeypatch.setenv("CERB_TEST_MARG_SMEM", str(limit))
    cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs = 2, 8, 8 * 11
    _check(sim_backend(cfg), OracleBackend(cfg), np.random.default_rng(limit + m), 3, m, n, rank)


def test_marg_schur_more_windows_than_ctas_sim():
    """the per-CTA workspace is reused window after window (grid = min(windows, 2 x SMs))"""
    cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs = 2, 8, 8 * 11
    _check(sim_backend(cfg), OracleBackend(cfg), np.random.default_rng(5), 40, 3, 5, n_oracle=40)


def test_marg_schur_bad_arguments_sim():
    cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs = 2, 8, 8 * 11
    be = sim_backend(cfg)
    A, b = _problem(np.random.default_rng(0), 1, 3, 100)
    with pytest.raises(Exception):
        be.marginalize_schur(A, b, 3)                              # n over CERB_MAX_PRIOR_DIM


@pytest.mark.gpu
@pytest.mark.parametrize("B,m,n,rank", [(32, 169, 86, None), (8, 6, 80, None), (4, 40, 86, 60), (300, 35, 86, None), (3, 260, 86, None)])
def test_marg_schur_gpu(B, m, n, rank):
    """MARGIN_OLD at the 150-feature size (m = 19 + 150), MARGIN_SECOND_NEW (m = 6) and a rank-deficient problem"""
    cfg = abi.default_config(); cfg.max_batch, cfg.max_features, cfg.max_obs = 2, 8, 8 * 11
    _check(lib.Backend(cfg), OracleBackend(cfg), np.random.default_rng(B), B, m, n, rank, n_oracle=1)
