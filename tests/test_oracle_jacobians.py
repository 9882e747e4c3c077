"""Self-consistency of the CPU oracle: analytic Jacobians vs central differences.

Automates the procedure the reference only prints (ProjectionTwoFrameOneCamFactor::check,
projectionTwoFrameOneCamFactor.cpp:152-272; IMULegFactor::checkJacobian, imu_leg_factor.cpp:7-171):
perturb every tangent direction (pose blocks through PoseLocalParameterization::Plus) and compare.
"""
import numpy as np
import pytest
from cerberus_b200 import abi, synth
from oracle_lib import OracleBackend

ob = OracleBackend()
EPS = 1e-6


def pose_plus(x, d):
    """PoseLocalParameterization::Plus (pose_local_parameterization.cpp:12-30), batched."""
    out = x.copy()
    out[:, :3] += d[:, :3]
    q = x[:, 3:7]
    dq = np.concatenate([d[:, 3:6] / 2, np.ones((x.shape[0], 1))], axis=1)
    x1, y1, z1, w1 = q.T
    x2, y2, z2, w2 = dq.T
    r = np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                  w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=1)
    out[:, 3:7] = r / np.linalg.norm(r, axis=1, keepdims=True)
    return out


def rand_pose(rng, n, scale=1.0):
    p = rng.normal(0, scale, (n, 3))
    q = rng.normal(0, 1, (n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([p, q], axis=1)


def proj_inputs(rng, n):
    """Geometrically sane random projection factors: a point in front of both frames."""
    batch = synth.generate_batch(1, 8, ob, with_prior=False)
    st = batch.state_array()
    pi = np.repeat(st["para_Pose"][0, 0][None], n, 0)
    pj = st["para_Pose"][0, rng.integers(1, 11, n)]
    ex0 = np.repeat(st["para_Ex_Pose"][0, 0][None], n, 0)
    ex1 = np.repeat(st["para_Ex_Pose"][0, 1][None], n, 0)
    lam = rng.uniform(0.07, 0.5, n)
    td = rng.normal(0, 0.002, n)
    pts_i = np.concatenate([rng.uniform(-0.5, 0.5, (n, 2)), np.ones((n, 1))], axis=1)
    pts_j = np.concatenate([rng.uniform(-0.5, 0.5, (n, 2)), np.ones((n, 1))], axis=1)
    vel_i, vel_j = rng.normal(0, 0.3, (n, 2)), rng.normal(0, 0.3, (n, 2))
    td_i, td_j = rng.normal(0, 0.002, n), rng.normal(0, 0.002, n)
    return [pi, pj, ex0, ex1, lam, td, pts_i, pts_j, vel_i, vel_j, td_i, td_j]


BLOCKS = {abi.PROJ_TWO_FRAME_ONE_CAM: [(0, 7), (1, 7), (2, 7), (4, 1), (5, 1)],
          abi.PROJ_TWO_FRAME_TWO_CAM: [(0, 7), (1, 7), (2, 7), (3, 7), (4, 1), (5, 1)],
          abi.PROJ_ONE_FRAME_TWO_CAM: [(2, 7), (3, 7), (4, 1), (5, 1)]}


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_projection_jacobians(kind):
    rng = np.random.default_rng(10 + kind)
    n = 64
    args = proj_inputs(rng, n)
    if kind == abi.PROJ_ONE_FRAME_TWO_CAM:
        args[10] = args[5].copy()   # td_i == td: the kept quirk (pts_i vs pts_i_td, :119) vanishes
    res, jac = ob.eval_projection(kind, *args)
    off = 0
    for (ai, size) in BLOCKS[kind]:
        local = 6 if size == 7 else 1
        for c in range(local):
            ap, am = [a.copy() for a in args], [a.copy() for a in args]
            if size == 7:
                d = np.zeros((n, 6)); d[:, c] = EPS
                ap[ai], am[ai] = pose_plus(args[ai], d), pose_plus(args[ai], -d)
            else:
                ap[ai] = args[ai] + EPS; am[ai] = args[ai] - EPS
            rp, _ = ob.eval_projection(kind, *ap, want_jac=False)
            rm, _ = ob.eval_projection(kind, *am, want_jac=False)
            num = (rp - rm) / (2 * EPS)
            ana = jac[:, off:off + 2 * size].reshape(n, 2, size)[:, :, c]
            scale = np.maximum(1.0, np.abs(ana).max())
            assert np.abs(num - ana).max() / scale < 2e-5, (kind, ai, c, np.abs(num - ana).max(), scale)
        if size == 7:
            assert np.all(jac[:, off:off + 14].reshape(n, 2, 7)[:, :, 6] == 0.0)
        off += 2 * size


def test_one_frame_two_cam_lambda_quirk():
    """d r / d lambda of ProjectionOneFrameTwoCamFactor uses pts_i, not pts_i_td (:119): with td != td_i
    the analytic column must differ from the numeric one by exactly that substitution."""
    rng = np.random.default_rng(5)
    n = 16
    args = proj_inputs(rng, n)
    args[5] = args[10] + 0.05            # td - td_i = 0.05 s
    _, jac = ob.eval_projection(2, *args)
    ap, am = [a.copy() for a in args], [a.copy() for a in args]
    ap[4] = args[4] + EPS; am[4] = args[4] - EPS
    num = (ob.eval_projection(2, *ap, want_jac=False)[0] - ob.eval_projection(2, *am, want_jac=False)[0]) / (2 * EPS)
    ana = jac[:, 28:30]
    assert np.abs(num - ana).max() > 1e-3        # they differ ...
    a2 = [a.copy() for a in args]
    a2[6] = args[6].copy(); a2[6][:, :2] = args[6][:, :2] - 0.05 * args[8]   # ... unless pts_i := pts_i_td, td := td_i
    a2[5] = args[10].copy()
    _, jac2 = ob.eval_projection(2, *a2)
    # same geometry now => analytic(pts_i_td) equals numeric of the original
    assert np.abs(jac2[:, 28:30] - num).max() / np.abs(num).max() < 1e-5


def imu_leg_setup(n=4):
    batch = synth.generate_batch(1, 4, ob, with_prior=False)
    st = batch.state_array()
    pre = batch.preint[0][:n].copy()
    params = np.zeros((n, 40))
    for k in range(n):
        params[k, 0:7] = st["para_Pose"][0, k]; params[k, 7:16] = st["para_SpeedBias"][0, k]; params[k, 16:20] = st["para_LegBias"][0, k]
        params[k, 20:27] = st["para_Pose"][0, k + 1]; params[k, 27:36] = st["para_SpeedBias"][0, k + 1]; params[k, 36:40] = st["para_LegBias"][0, k + 1]
    rng = np.random.default_rng(3)
    params[:, 10:16] += rng.normal(0, 1e-3, (n, 6))     # nonzero bias offsets so the correction terms are exercised
    params[:, 16:20] += rng.normal(0, 1e-3, (n, 4))
    return pre, params


def test_imu_leg_jacobians():
    pre, params = imu_leg_setup()
    n = params.shape[0]
    res, jac, si = ob.eval_imu_leg(pre, params)
    J = jac.reshape(n, -1)
    blocks = [(0, 7, 0), (7, 9, 31 * 7), (16, 4, 31 * 16), (20, 7, 31 * 20), (27, 9, 31 * 27), (36, 4, 31 * 36)]
    # unwhiten: compare S^-1 J (the physical Jacobian) against numeric differences of the unwhitened residual
    for (poff, size, joff) in blocks:
        local = 6 if size == 7 else size
        Jb = J[:, joff:joff + 31 * size].reshape(n, 31, size)
        for c in range(local):
            eps = 1e-7 if size != 4 else 1e-9
            pp, pm = params.copy(), params.copy()
            if size == 7:
                d = np.zeros((n, 6)); d[:, c] = eps
                pp[:, poff:poff + 7] = pose_plus(params[:, poff:poff + 7], d); pm[:, poff:poff + 7] = pose_plus(params[:, poff:poff + 7], -d)
            else:
                pp[:, poff + c] += eps; pm[:, poff + c] -= eps
            rp = ob.eval_imu_leg(pre, pp, want_jac=False)[0]
            rm = ob.eval_imu_leg(pre, pm, want_jac=False)[0]
            num = (rp - rm) / (2 * eps)
            ana = Jb[:, :, c]
            scale = np.maximum(1.0, np.abs(ana).max(axis=1, keepdims=True))
            # the reference's analytic d r_q / d theta blocks are first order (checkJacobian tolerates that): 2e-3 rel
            assert (np.abs(num - ana) / scale).max() < 2e-3, (poff, c, (np.abs(num - ana) / scale).max())
    # sqrt_info is upper triangular and S^T S == cov^-1
    S = si.reshape(n, 31, 31)
    cov = np.stack([np.array(p["covariance"]).reshape(31, 31) for p in pre])
    assert np.abs(np.tril(S, -1)).max() == 0.0
    prod = np.einsum("nki,nkj->nij", S, S) @ cov
    assert np.abs(prod - np.eye(31)).max() < 1e-6


def test_a1_kinematics_derivatives():
    rng = np.random.default_rng(0)
    n = 32
    q = rng.uniform(-1, 1, (n, 3)) + np.array([0, 0.8, -1.6])
    lc = rng.uniform(0.19, 0.23, n)
    fix = np.tile(np.array([0.1805, 0.047, 0.0838, 0.21]), (n, 1)) * rng.choice([-1, 1], (n, 4))
    fix[:, 3] = 0.21
    fk, jac, dfk, djq, djr = ob.a1_kinematics(q, lc, fix)
    assert np.abs(fk - synth.a1_fk(q, lc, fix)).max() < 1e-14          # independent numpy statement of the FK
    J = jac.reshape(n, 3, 3).transpose(0, 2, 1)                          # column-major storage -> J[r,c]
    assert np.abs(J - synth.a1_jac(q, lc, fix)).max() < 1e-14
    e = 1e-6
    for c in range(3):
        d = np.zeros(3); d[c] = e
        fp, jp = ob.a1_kinematics(q + d, lc, fix)[:2]
        fm, jm = ob.a1_kinematics(q - d, lc, fix)[:2]
        assert np.abs((fp - fm) / (2 * e) - J[:, :, c]).max() < 1e-8
        assert np.abs((jp - jm) / (2 * e) - djq[:, 9 * c:9 * c + 9]).max() < 1e-8
    fp, jp = ob.a1_kinematics(q, lc + e, fix)[:2]
    fm, jm = ob.a1_kinematics(q, lc - e, fix)[:2]
    assert np.abs((fp - fm) / (2 * e) - dfk).max() < 1e-8
    assert np.abs((jp - jm) / (2 * e) - djr).max() < 1e-8
