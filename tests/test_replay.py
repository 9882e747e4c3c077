"""Sequence replay: B independent robots run the reference's steady-state frame loop
   processIMULeg -> processImage: addFeatureCheckParallax -> triangulate -> optimization (solve, double2vector, marginalization)
   -> outliersRejection -> slideWindow (estimator.cpp:590-846, 1054-1677)
through cerberus_b200.estimator (the host mirror of Estimator / FeatureManager); every numerical step is a call through the C ABI.
The same driver runs on the CPU oracle (tests/oracle_lib.OracleOps) -- the reference arm -- and the published pose of every frame is
compared.  Eigen-solver of the oracle arm = cyclic Jacobi like the device (identical arithmetic => tight tolerance); the distance to the
tridiagonal-QR arm (the reference's kind of solver) is reported and bounded by the measured sensitivity of the chain
(profiles/eig_study_r2.txt: 1e-4 .. 1e-3 m between ANY two rounding-different marginalizations)."""
import os
import time
import numpy as np
import pytest
from cerberus_b200 import abi, synth, estimator, lib
from oracle_lib import OracleOps
from helpers import sim_backend


def _replay(make_backend, n, n_frames, tracked, F, iters, eig_modes=(1,), csv_path=None):
    cfg = abi.default_config(); cfg.max_batch = n; cfg.max_features = 2 * F; cfg.max_obs = 2 * F * 11; cfg.max_num_iterations = iters
    pcfg = abi.default_preint_config()
    seq = synth.generate_sequence(n, n_frames, tracked=tracked, max_len=14, min_len=3)
    t0 = time.perf_counter()
    dev = estimator.ReplayDriver(estimator.DeviceOps(make_backend(cfg), cfg), cfg, pcfg, n, max_features=F).run(seq)
    t_dev = time.perf_counter() - t0
    arms = {m: estimator.ReplayDriver(OracleOps(cfg, eig_mode=m), cfg, pcfg, n, max_features=F).run(seq) for m in eig_modes}
    if csv_path:
        for e in dev.est: estimator.write_csv(csv_path, e, pcfg)
    return seq, dev, arms, t_dev


def _deltas(a, b):
    Pa, Ra = a.poses(); Pb, Rb = b.poses()
    return np.abs(Pa - Pb).max(axis=(0, 2)), np.abs(Ra - Rb).max(axis=(0, 2, 3))


def test_replay_matches_oracle_sim(tmp_path):
    """CPU tier: the kernels on the CPU simulator, 1 robot, 5 chained frames."""
    csv = str(tmp_path / "vilo.csv")
    seq, dev, arms, _ = _replay(sim_backend, 1, 15, 14, 24, 4, csv_path=csv)
    dP, dR = _deltas(dev, arms[1])
    assert dP.shape[0] == 5 and dP.max() < 1e-6 and dR.max() < 1e-6, (dP, dR)
    for a, b in zip(dev.reports, arms[1].reports):
        assert (a["iterations"] == b["iterations"]).all()
    # same bookkeeping on both arms: marginalization flags, feature lists, priors
    for e0, e1 in zip(dev.est, arms[1].est):
        assert [f.feature_id for f in e0.f_manager.feature] == [f.feature_id for f in e1.f_manager.feature]
        assert (e0.prior is None) == (e1.prior is None)
    # result file of the reference's main loop (main.cpp:153-197): one row of 20 comma-terminated columns per processed frame
    rows = open(csv).read().strip().split("\n")
    assert len(rows) == 5 and all(len(r.rstrip(",").split(",")) == 20 for r in rows)
    # the estimate stays near the truth (a replay that diverged would still be "equal" on both arms)
    P, _ = dev.poses()
    assert np.linalg.norm(P[0, -1] - seq.p[0, 10 + P.shape[1] - 1]) < 0.3


@pytest.mark.gpu
def test_replay_50_frames_gpu():
    """GPU tier: 4 robots x 62 frames = 52 chained optimization() calls each through the sm_100a library."""
    n, n_frames = 4, 62
    seq, dev, arms, t_dev = _replay(lambda cfg: lib.Backend(cfg), n, n_frames, 90, 160, 12, eig_modes=(1, 0))
    steps = n_frames - 10
    dP, dR = _deltas(dev, arms[1])
    dPq, dRq = _deltas(dev, arms[0])
    P, _ = dev.poses()
    err = np.linalg.norm(P - seq.p[:, 10:10 + P.shape[1]], axis=-1)
    T = dev.timing
    lines = [f"replay: {n} robots x {steps} frames through cerberus_b200 (sm_100a), wall {t_dev:.2f} s = {n * steps / t_dev:.1f} frames/s "
             f"(host mirror is Python; device + ABI time: solve {T['solve']:.2f} s, marginalize {T['marginalize']:.2f} s, preintegrate {T['preintegrate']:.2f} s, "
             f"triangulate {T['triangulate']:.2f} s, outliers {T['outliers']:.2f} s, shift {T['shift']:.2f} s, host bookkeeping {T['host']:.2f} s)",
             "per-frame max |published position delta| device vs oracle (same Jacobi eigen arithmetic) [m]: " + " ".join(f"{v:.1e}" for v in dP),
             "per-frame max rotation-matrix delta                                                          : " + " ".join(f"{v:.1e}" for v in dR),
             "per-frame max |published position delta| device vs oracle with the tridiagonal-QR solver [m]  : " + " ".join(f"{v:.1e}" for v in dPq),
             f"max over the replay: Jacobi arm {dP.max():.2e} m / {dR.max():.2e} rad; QR arm {dPq.max():.2e} m / {dRq.max():.2e} rad; "
             f"distance from ground truth at the end {err[:, -1].round(3)} m"]
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/replay_gpu.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    assert dP.shape[0] == steps >= 50
    # One optimization() on identical inputs agrees to ~1e-9 m (tests/test_gpu_parity.py); along a chain every arm re-linearises on its own
    # states, and the eps-clamped eigen factoring of the marginalization amplifies rounding-level differences: the first frames stay at
    # solve-level agreement, later ones wander inside the chain's measured sensitivity band (1e-4 .. 1.3e-3 m between ANY two
    # rounding-different arms over 40 frames, also two QR arms: profiles/eig_study_r2.txt; measured here on the B200: 3.7e-4 m max against
    # the Jacobi arm, 3.7e-4 m against the QR arm over 52 frames) -- three orders of magnitude below the 3 .. 9 cm distance of every arm from the truth.
    assert dP[:3].max() < 1e-5 and dR[:3].max() < 1e-5, (dP[:3], dR[:3])
    assert dP.max() < 2e-3 and dR.max() < 2e-3, (dP.max(), dR.max())
    assert dPq.max() < 2e-3
    Po, _ = arms[1].poses()
    err_o = np.linalg.norm(Po - seq.p[:, 10:10 + Po.shape[1]], axis=-1)
    assert np.abs(err - err_o).max() < 2e-3 and err.max() < 0.5
