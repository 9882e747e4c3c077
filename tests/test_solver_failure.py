"""Ceres' LINEAR_SOLVER_FAILURE path (DoglegStrategy::ComputeGaussNewtonStep / TrustRegionMinimizer::HandleInvalidStep, Ceres 1.14).

A factorisation of J^T J + 1e-8 D^2 (Jacobi-scaled) practically never fails in fp64, so the path is exercised by fault injection: with
CERB_TEST_FAIL_FACTORIZATIONS=k both the kernel and the oracle report the first k Gauss-Newton solves of every window as failed, with
CERB_TEST_INITIAL_MU they start the strategy at another mu.  What Ceres does then, and what both arms must agree on:
  * a failed solve multiplies mu by 10 and is retried inside the same ComputeStep -- no iteration, no invalid step is consumed;
  * once mu reaches max_mu = 1 the strategy reports LINEAR_SOLVER_FAILURE: the step is invalid (num_consecutive_invalid_steps + 1,
    StepIsInvalid -> mu *= 10 again), every later ComputeStep fails immediately, and after 5 consecutive invalid steps the minimizer
    terminates with FAILURE at iteration 5, states untouched."""
import os
import numpy as np
import pytest
from cerberus_b200 import abi, synth, lib
from oracle_lib import OracleBackend
from helpers import sim_backend, small_cfg, state_diffs


def _run(make_backend, env, nw=1, F=10, iters=6):
    old = {k: os.environ.get(k) for k in ("CERB_TEST_FAIL_FACTORIZATIONS", "CERB_TEST_INITIAL_MU")}
    os.environ.update(env)
    try:
        cfg = small_cfg(max_batch=4, max_features=16, iters=iters)
        o, be = OracleBackend(cfg), make_backend(cfg)           # the hooks are read when the handle is created / at every oracle solve
        batch = synth.generate_batch(nw, F, o, window0=610, prior_features=6)
        st = batch.state_array(); saved = batch.copy_states(); x0 = st.copy()
        rep_o = o.solve_batch(batch); ref = st.copy(); lam = batch.para_Feature.copy()
        batch.restore_states(saved)
        rep_b = be.solve_batch(batch)
        return rep_o, rep_b, ref, batch.state_array().copy(), lam, batch.para_Feature.copy(), x0
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v


def _check_retry_inside_the_step(make_backend):
    rep_o, rep_b, ref, got, lam_o, lam_b, _ = _run(make_backend, {"CERB_TEST_FAIL_FACTORIZATIONS": "3"})
    # three failed solves: mu 1e-8 -> 1e-5 inside the first ComputeStep; no iteration / invalid step consumed on either arm
    assert (rep_o["iterations"] == rep_b["iterations"]).all() and (rep_o["num_successful_steps"] == rep_b["num_successful_steps"]).all()
    assert (rep_o["termination"] == rep_b["termination"]).all() and (rep_b["termination"] != abi.TERM_FAILURE).all()
    assert (rep_b["num_successful_steps"] >= 1).all()
    assert np.abs(rep_o["final_cost"] - rep_b["final_cost"]).max() < 1e-7 * rep_o["final_cost"].max()
    d = state_diffs(got, ref)
    assert d["para_Pose"] < 1e-6 and d["para_SpeedBias"] < 1e-5 and np.abs(lam_o - lam_b).max() < 1e-6, d
    assert (rep_b["final_cost"] < rep_b["initial_cost"]).all()


def _check_failure_termination(make_backend, env):
    rep_o, rep_b, ref, got, lam_o, lam_b, x0 = _run(make_backend, env)
    for rep in (rep_o, rep_b):
        assert (rep["iterations"] == 5).all() and (rep["num_successful_steps"] == 0).all() and (rep["termination"] == abi.TERM_FAILURE).all()
        assert (rep["final_cost"] == rep["initial_cost"]).all()
    for k in ("para_Pose", "para_SpeedBias", "para_LegBias", "para_Ex_Pose"):
        assert (got[k] == x0[k]).all() and (ref[k] == x0[k]).all()


def test_mu_retry_inside_the_step_sim():
    _check_retry_inside_the_step(sim_backend)


def test_eight_failures_reach_max_mu_sim():
    """mu 1e-8 x 10^8 = 1 = max_mu after eight failed solves: LINEAR_SOLVER_FAILURE -> five invalid steps -> FAILURE."""
    _check_failure_termination(sim_backend, {"CERB_TEST_FAIL_FACTORIZATIONS": "8"})


def test_initial_mu_at_max_mu_sim():
    """mu >= max_mu: the solve is not even attempted (`while (mu_ < max_mu_)`)."""
    _check_failure_termination(sim_backend, {"CERB_TEST_INITIAL_MU": "1.0"})


def test_seven_failures_still_solve_sim():
    """mu = 0.1 after seven failures is still < max_mu: the eighth attempt succeeds and the solve goes on (heavily damped)."""
    rep_o, rep_b, ref, got, lam_o, lam_b, _ = _run(sim_backend, {"CERB_TEST_FAIL_FACTORIZATIONS": "7"}, iters=4)
    assert (rep_o["iterations"] == rep_b["iterations"]).all() and (rep_o["num_successful_steps"] == rep_b["num_successful_steps"]).all()
    assert (rep_b["termination"] != abi.TERM_FAILURE).all() and (rep_b["num_successful_steps"] >= 1).all()
    assert state_diffs(got, ref)["para_Pose"] < 1e-6


@pytest.mark.gpu
def test_mu_retry_and_failure_termination_gpu():
    mk = lambda cfg: lib.Backend(cfg)
    _check_retry_inside_the_step(mk)
    _check_failure_termination(mk, {"CERB_TEST_FAIL_FACTORIZATIONS": "8"})
    _check_failure_termination(mk, {"CERB_TEST_INITIAL_MU": "1.0"})
