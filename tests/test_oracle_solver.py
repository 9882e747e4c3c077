"""Behaviour of the CPU oracle's Ceres-1.14 restatement and of the host-side helpers."""
import numpy as np
import ctypes as C
from cerberus_b200 import abi, synth
from oracle_lib import OracleBackend
from helpers import small_cfg, sim_backend

ob = OracleBackend()


def test_solve_reduces_cost_and_moves_towards_truth():
    batch, truth = synth.generate_batch(2, 40, ob, return_truth=True)
    st = batch.state_array()
    lam0 = batch.para_Feature[:, :40].copy()
    rep = ob.solve_batch(batch)
    assert (rep["final_cost"] < 0.01 * rep["initial_cost"]).all()
    assert (rep["iterations"] <= 12).all() and (rep["status"] == 0).all()
    # 0.5 px pixel noise whitened by 460/1.5 gives ~0.11 per visual residual: the optimum is a statistically sane fit
    n_res = 2 * 40 * 21 + 310 + 86
    assert (rep["final_cost"] < 0.5 * n_res).all() and (rep["final_cost"] > 0.01 * n_res).all()
    # inverse depths stay within the stereo triangulation uncertainty (sigma ~ 0.03 per pair) of the truth
    assert np.abs(batch.para_Feature[:, :40] - truth.lam).mean() < 0.03
    q = st["para_Pose"][:, :, 3:7]
    assert np.abs(np.linalg.norm(q, axis=-1) - 1).max() < 1e-12     # Plus() keeps unit quaternions


def test_constant_blocks_stay_constant():
    cfg = abi.default_config(); cfg.optimize_leg_bias = 0
    o2 = OracleBackend(cfg)
    batch = synth.generate_batch(1, 12, o2)
    batch.descs[0].extrinsic_open = 0
    st = batch.state_array()
    lb0, ex0, td0 = st["para_LegBias"].copy(), st["para_Ex_Pose"].copy(), st["para_Td"].copy()
    o2.solve_batch(batch)
    assert (st["para_LegBias"] == lb0).all() and (st["para_Ex_Pose"] == ex0).all() and (st["para_Td"] == td0).all()


def test_max_iterations_and_termination():
    cfg = abi.default_config(); cfg.max_num_iterations = 3
    batch = synth.generate_batch(1, 12, ob)
    rep = OracleBackend(cfg).solve_batch(batch)
    assert rep["iterations"][0] == 3 and rep["termination"][0] == abi.TERM_NO_CONVERGENCE
    cfg.max_num_iterations = 50; cfg.function_tolerance = 1e-2
    batch = synth.generate_batch(1, 12, ob)
    rep = OracleBackend(cfg).solve_batch(batch)
    assert rep["termination"][0] == abi.TERM_CONVERGENCE and rep["iterations"][0] < 50


def test_double2vector_keeps_the_gauge():
    """estimator.cpp:903-957: after the solve the window is rotated back so that frame 0 keeps its yaw and position."""
    batch = synth.generate_batch(1, 12, ob)
    before = abi.WindowState(); C.memmove(C.byref(before), C.byref(batch.states[0]), C.sizeof(abi.WindowState))
    ob.solve_batch(batch)
    Ps, Rs, Vs = ob.double2vector(before, batch.states[0])
    assert np.abs(Ps[0] - np.array(before.para_Pose[0][:3])).max() < 1e-12
    R0 = synth.R_from_quat(np.array(before.para_Pose[0][3:7]))
    yaw = lambda R: np.arctan2(R[1, 0], R[0, 0])
    assert abs(yaw(Rs[0]) - yaw(R0)) < 1e-9
    # the product's host helper is the same function
    sb = sim_backend(small_cfg())
    Ps2, Rs2, Vs2 = sb.double2vector(before, batch.states[0])
    assert np.abs(Ps - Ps2).max() < 1e-12 and np.abs(Rs - Rs2).max() < 1e-12 and np.abs(Vs - Vs2).max() < 1e-12
    # relative geometry is untouched by the gauge change
    d_solved = np.linalg.norm(np.array(batch.states[0].para_Pose[5][:3]) - np.array(batch.states[0].para_Pose[0][:3]))
    assert abs(np.linalg.norm(Ps[5] - Ps[0]) - d_solved) < 1e-12


def test_marginalization_prior_is_consistent():
    """The prior produced by MARGIN_OLD must be a valid linearised factor: at its linearisation point its gradient
    J^T r equals the marginal gradient, and it stays quadratic: cost(dx) = 0.5 |r0 + J dx|^2."""
    batch = synth.generate_batch(1, 16, ob, prior_features=8)
    pr = batch.descs[0].prior
    assert pr.valid == 1 and pr.n == 86 and pr.num_blocks == 15
    n = pr.n
    J = batch.prior_J[0][:n * n].reshape(n, n).T
    A = J.T @ J
    ev = np.linalg.eigvalsh(A)
    assert ev.min() > -1e-6 * ev.max()
    kinds = sorted((pr.block_kind[b], pr.block_index[b]) for b in range(pr.num_blocks))
    assert kinds == [(0, k) for k in range(10)] + [(1, 0), (2, 0), (3, 0), (3, 1), (4, 0)]
    # at x == x0 the residual is linearized_residuals
    r, _ = ob.eval_prior(pr, batch.states[0], 7 * 12 + 9 + 4 + 1)
    assert np.abs(r - batch.prior_r[0][:n]).max() < 1e-9
