"""Synthetic sliding-window generator (SURVEY.md section 8(d) input spec).

Produces batches of 10-frame (11-state) windows with F stereo features, A1 trot leg data and a dense
marginalization prior, for the parity tests and for bench.py.  Pure numpy host code; the two steps
that belong to the backend under test are delegated to a `backend` object:

    backend.preintegrate(pcfg, jobs, n) -> np.ndarray[n] of abi.preint_dtype
        (IMULegIntegrationBase::push_back loop; product: cerb_preintegrate_batch on the GPU)
    backend.marginalize(cfg, batch, margin_old) -> fills prior of a *target* batch
        (optimization() marginalization half; see cerberus_b200.marginalization)

Window w uses numpy PCG64 seeded with 0xCE2BE205 + w, so any backend sees identical raw inputs.
Trajectory: forward 0.5 m/s, yaw rate U[-0.3,0.3] rad/s, roll/pitch 0.03 rad sinusoids, 1 cm z bounce
at 2 Hz; gravity (0,0,9.805); frames at 15 Hz, 33 IMU/leg samples per interval.
"""
import ctypes as C
import numpy as np
from . import abi

SEED0 = 0xCE2BE205
FRAME_DT = 1.0 / 15.0
SAMPLES_PER_FRAME = 33
G_NORM = 9.805
RIC = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])   # body_T_cam0/1 rotation (A1 yaml :55-71)
TIC = np.array([[0.10076, 0.025, 0.1114], [0.10076, -0.025, 0.1114]])
PHI_NOMINAL = np.array([0.0, 0.8, -1.6])
LC_NOMINAL = 0.21


# ----------------------------------------------------------------------------- small rotation helpers
def rot_zyx(y, p, r):
    """R = Rz(y) Ry(p) Rx(r); y,p,r broadcastable arrays -> [...,3,3]."""
    cy, sy, cp, sp, cr, sr = np.cos(y), np.sin(y), np.cos(p), np.sin(p), np.cos(r), np.sin(r)
    R = np.empty(np.broadcast(y, p, r).shape + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def quat_from_R(R):
    """[...,3,3] -> [...,4] (x,y,z,w), w >= 0."""
    R = np.asarray(R)
    out = np.empty(R.shape[:-2] + (4,))
    flatR = R.reshape(-1, 3, 3)
    flat = out.reshape(-1, 4)
    for k in range(flatR.shape[0]):
        m = flatR[k]
        t = m[0, 0] + m[1, 1] + m[2, 2]
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            w, x, y, z = 0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            w, x, y, z = (m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            w, x, y, z = (m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            w, x, y, z = (m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s
        if w < 0:
            w, x, y, z = -w, -x, -y, -z
        flat[k] = (x, y, z, w)
    return out


def R_from_quat(q):
    """[...,4] (x,y,z,w) -> [...,3,3]."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - w * z)
    R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y)
    R[..., 2, 1] = 2 * (y * z + w * x)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def so3_exp(v):
    """[...,3] rotation vector -> [...,3,3]."""
    th = np.linalg.norm(v, axis=-1)[..., None, None]
    K = np.zeros(v.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -v[..., 2], v[..., 1]
    K[..., 1, 0], K[..., 1, 2] = v[..., 2], -v[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -v[..., 1], v[..., 0]
    with np.errstate(invalid="ignore", divide="ignore"):
        a = np.where(th > 1e-9, np.sin(th) / th, 1.0)
        b = np.where(th > 1e-9, (1 - np.cos(th)) / (th * th), 0.5)
    return np.eye(3) + a * K + b * (K @ K)


# ----------------------------------------------------------------------------- A1 leg model (truth side)
def a1_fk(q, lc, fix):
    """q [...,3], lc [...], fix [...,4] = (ox, oy, d, lt) -> foot position in body frame [...,3]."""
    c0, s0, c1, s1 = np.cos(q[..., 0]), np.sin(q[..., 0]), np.cos(q[..., 1]), np.sin(q[..., 1])
    c12, s12 = np.cos(q[..., 1] + q[..., 2]), np.sin(q[..., 1] + q[..., 2])
    ox, oy, d, lt = fix[..., 0], fix[..., 1], fix[..., 2], fix[..., 3]
    return np.stack([ox - lt * s1 - lc * s12, oy + d * c0 + lt * c1 * s0 + lc * s0 * c12, d * s0 - lt * c0 * c1 - lc * c0 * c12], axis=-1)


def a1_jac(q, lc, fix):
    c0, s0, c1, s1 = np.cos(q[..., 0]), np.sin(q[..., 0]), np.cos(q[..., 1]), np.sin(q[..., 1])
    c12, s12 = np.cos(q[..., 1] + q[..., 2]), np.sin(q[..., 1] + q[..., 2])
    d, lt = fix[..., 2], fix[..., 3]
    A, B = lt * s1 + lc * s12, lt * c1 + lc * c12
    J = np.zeros(q.shape[:-1] + (3, 3))
    J[..., 1, 0] = -d * s0 + c0 * B
    J[..., 2, 0] = d * c0 + s0 * B
    J[..., 0, 1] = -B
    J[..., 1, 1] = -s0 * A
    J[..., 2, 1] = c0 * A
    J[..., 0, 2] = -lc * c12
    J[..., 1, 2] = -s0 * lc * s12
    J[..., 2, 2] = c0 * lc * s12
    return J


# ----------------------------------------------------------------------------- the generator
class SynthTruth:
    """Ground truth + raw measurements of a generated batch (frames -1..10 => index 0..11)."""
    pass


def _trajectory(par, t):
    """par: dict of [B,1] arrays; t [T] -> dict of [B,T,...] world pose / velocity / accel / body rate."""
    yaw = par["yaw0"] + par["wz"] * t
    roll = 0.03 * np.sin(2 * np.pi * 1.0 * t + par["ph_r"])
    pitch = 0.03 * np.sin(2 * np.pi * 1.3 * t + par["ph_p"])
    droll = 0.03 * 2 * np.pi * 1.0 * np.cos(2 * np.pi * 1.0 * t + par["ph_r"])
    dpitch = 0.03 * 2 * np.pi * 1.3 * np.cos(2 * np.pi * 1.3 * t + par["ph_p"])
    dyaw = par["wz"] + 0 * t
    R = rot_zyx(yaw, pitch, roll)
    wz = par["wz"]
    small = np.abs(wz) < 1e-6
    wz_s = np.where(small, 1.0, wz)
    sp = par.get("sp", 0.5)                      # forward speed [m/s] (0.5)
    x = np.where(small, sp * np.cos(par["yaw0"]) * t, sp * (np.sin(yaw) - np.sin(par["yaw0"])) / wz_s)
    y = np.where(small, sp * np.sin(par["yaw0"]) * t, -sp * (np.cos(yaw) - np.cos(par["yaw0"])) / wz_s)
    wb = 4 * np.pi
    z = 0.01 * (np.sin(wb * t + par["ph_z"]) - np.sin(par["ph_z"]))
    p = np.stack([x + par["x0"], y + par["y0"], z + par["z0"]], axis=-1)
    v = np.stack([sp * np.cos(yaw), sp * np.sin(yaw), 0.01 * wb * np.cos(wb * t + par["ph_z"])], axis=-1)
    a = np.stack([-sp * wz * np.sin(yaw), sp * wz * np.cos(yaw), -0.01 * wb * wb * np.sin(wb * t + par["ph_z"])], axis=-1)
    # body rate for ZYX euler angles
    w_b = np.stack([droll - dyaw * np.sin(pitch),
                    dpitch * np.cos(roll) + dyaw * np.sin(roll) * np.cos(pitch),
                    -dpitch * np.sin(roll) + dyaw * np.cos(roll) * np.cos(pitch)], axis=-1)
    return dict(R=R, p=p, v=v, a=a, w=w_b)


def generate_batch(n, n_features, backend, cfg=None, pcfg=None, window0=0, realistic=False, prior_features=None,
                   with_prior=True, outlier_fraction=0.0, return_truth=False, use_leg=True):
    """Generate windows window0 .. window0+n-1.  Returns an abi.WindowBatch (and SynthTruth if asked).
    use_leg=False produces USE_LEG == 0 windows (IMUFactor instead of IMULegFactor, no leg bias, no prior)."""
    cfg = cfg or abi.default_config()
    pcfg = pcfg or abi.default_preint_config()
    B, F = n, n_features
    NF = 12                                    # frames -1..10
    S = SAMPLES_PER_FRAME
    T = (NF - 1) * S + 1
    dt = FRAME_DT / S
    t = np.arange(T) * dt
    rngs = [np.random.Generator(np.random.PCG64(SEED0 + window0 + w)) for w in range(B)]

    def draw(fn):
        return np.stack([fn(r) for r in rngs])

    par = {
        "yaw0": draw(lambda r: r.uniform(-np.pi, np.pi, 1)), "wz": draw(lambda r: r.uniform(-0.3, 0.3, 1)),
        "ph_r": draw(lambda r: r.uniform(0, 2 * np.pi, 1)), "ph_p": draw(lambda r: r.uniform(0, 2 * np.pi, 1)),
        "ph_z": draw(lambda r: r.uniform(0, 2 * np.pi, 1)),
        "x0": draw(lambda r: r.uniform(-5, 5, 1)), "y0": draw(lambda r: r.uniform(-5, 5, 1)), "z0": draw(lambda r: r.uniform(0.25, 0.35, 1)),
    }
    tr = _trajectory(par, t)                                   # [B,T,...]
    ba = draw(lambda r: r.normal(0, 0.05, 3))
    bg = draw(lambda r: r.normal(0, 0.005, 3))
    gvec = np.array([0.0, 0.0, G_NORM])
    RT = np.swapaxes(tr["R"], -1, -2)
    acc = np.einsum("btij,btj->bti", RT, tr["a"] + gvec) + ba[:, None, :] + draw(lambda r: r.normal(0, 0.1, (T, 3)))
    gyr = tr["w"] + bg[:, None, :] + draw(lambda r: r.normal(0, 0.01, (T, 3)))

    # ---- legs: integrate joint angles so that stance feet are stationary in the world ---------------
    fix = np.array([[pcfg.rho_fix[l][k] for k in range(4)] for l in range(4)])      # [4,4]
    lc_true = LC_NOMINAL + draw(lambda r: r.normal(0, 0.005, 4))                     # [B,4]
    phase0 = draw(lambda r: r.uniform(0, 0.5, 1))                                    # [B,1]
    gait = (np.floor((t[None, :] + phase0) / 0.25).astype(int) % 2)                  # [B,T]
    contact = np.stack([gait == 0, gait == 1, gait == 1, gait == 0], axis=-1).astype(float)   # [B,T,4] legs (0,3)/(1,2)
    phi = np.zeros((B, T, 4, 3))
    dphi = np.zeros((B, T, 4, 3))
    phi[:, 0] = PHI_NOMINAL + draw(lambda r: r.normal(0, 0.05, (4, 3)))
    v_body = np.einsum("btij,btj->bti", RT, tr["v"])                                 # R^T v_world

    def joint_rate(ph, k):
        f = a1_fk(ph, lc_true, fix)                                                  # [B,4,3]
        J = a1_jac(ph, lc_true, fix)                                                 # [B,4,3,3]
        rhs = v_body[:, k, None, :] + np.cross(tr["w"][:, k, None, :], f)            # R_br = I, p_br = 0
        st = -np.linalg.solve(J, rhs[..., None])[..., 0]
        sw = 4.0 * (PHI_NOMINAL - ph)
        c = contact[:, k, :, None]
        return c * st + (1 - c) * sw

    for k in range(T - 1):
        k1 = joint_rate(phi[:, k], k)
        dphi[:, k] = k1
        mid = phi[:, k] + 0.5 * dt * k1
        k2 = 0.5 * (joint_rate(mid, k) + joint_rate(mid, k + 1))
        phi[:, k + 1] = phi[:, k] + dt * k2
    dphi[:, T - 1] = joint_rate(phi[:, T - 1], T - 1)
    phi_m = phi + draw(lambda r: r.normal(0, 1e-4, (T, 4, 3)))
    dphi_m = dphi + draw(lambda r: r.normal(0, 0.02, (T, 4, 3)))

    # ---- frame states: truth and initial guess -------------------------------------------------------
    fidx = np.arange(NF) * S
    R_f, p_f, v_f = tr["R"][:, fidx], tr["p"][:, fidx], tr["v"][:, fidx]              # [B,12,...]
    R_g = R_f @ so3_exp(draw(lambda r: r.normal(0, 0.01, (NF, 3))))
    p_g = p_f + draw(lambda r: r.normal(0, 0.02, (NF, 3)))
    v_g = v_f + draw(lambda r: r.normal(0, 0.05, (NF, 3)))
    q_g = quat_from_R(R_g)
    ric_g = RIC @ so3_exp(draw(lambda r: r.normal(0, 0.005, (2, 3))))
    tic_g = TIC + draw(lambda r: r.normal(0, 0.005, (2, 3)))
    qic_g = quat_from_R(ric_g)

    # ---- landmarks and observations ---------------------------------------------------------------------
    def make_features(nf, first_frame, n_frames, rs_key):
        """nf features; anchors/lengths relative to `first_frame` (index into the 12 frames)."""
        if realistic:
            start = draw(lambda r: r.integers(0, 8, nf))
            length = np.stack([np.array([rngs[w].integers(4, n_frames - s + 1) for s in start[w]]) for w in range(B)])
            stereo = draw(lambda r: (r.uniform(0, 1, (nf, n_frames)) < 0.9))
        else:
            start = np.zeros((B, nf), dtype=int)
            length = np.full((B, nf), n_frames, dtype=int)
            stereo = np.ones((B, nf, n_frames), dtype=bool)
        depth = draw(lambda r: r.uniform(2, 15, nf))
        nx = draw(lambda r: r.uniform(-0.6, 0.6, nf))
        ny = draw(lambda r: r.uniform(-0.45, 0.45, nf))
        pc = np.stack([nx * depth, ny * depth, depth], axis=-1)                     # in cam0 of the anchor frame
        bi = np.arange(B)[:, None]
        Ra, pa = R_f[bi, first_frame + start], p_f[bi, first_frame + start]         # [B,nf,3,3]
        pw = np.einsum("bfij,bfj->bfi", Ra, pc @ RIC.T + TIC[0]) + pa
        # project into every frame of the block, both cameras
        Rb, pb = R_f[:, first_frame:first_frame + n_frames], p_f[:, first_frame:first_frame + n_frames]
        pbody = np.einsum("bkji,bfkj->bfki", Rb, pw[:, :, None, :] - pb[:, None, :, :])     # R^T (pw - p)  [B,nf,K,3]
        uv = np.zeros((B, nf, n_frames, 2, 2))
        for cam in range(2):
            pcam = (pbody - TIC[cam]) @ RIC                                          # ric^T (.)
            uv[..., cam, :] = pcam[..., :2] / pcam[..., 2:3]
        uv += draw(lambda r: r.normal(0, 0.5 / 460.0, (nf, n_frames, 2, 2)))
        if outlier_fraction > 0:
            bad = draw(lambda r: r.uniform(0, 1, (nf, n_frames, 2, 1)) < outlier_fraction)
            uv += bad * draw(lambda r: r.normal(0, 20.0 / 460.0, (nf, n_frames, 2, 2)))
        vel = np.zeros_like(uv)
        vel[:, :, 1:] = (uv[:, :, 1:] - uv[:, :, :-1]) / FRAME_DT
        vel[:, :, 0] = vel[:, :, 1]
        lam = (1.0 / depth) * (1 + draw(lambda r: r.normal(0, 0.1, nf)))
        return dict(start=start, length=length, stereo=stereo, uv=uv, vel=vel, lam=lam, lam_true=1.0 / depth)

    def fill_batch(batch, feats, first_frame, pre, with_states_from=0):
        """Write descriptors + initial states for the 11 frames starting at index first_frame."""
        st = batch.state_array()
        sl = slice(first_frame, first_frame + 11)
        st["para_Pose"][:, :, 0:3] = p_g[:, sl]
        st["para_Pose"][:, :, 3:7] = q_g[:, sl]
        st["para_SpeedBias"][:, :, 0:3] = v_g[:, sl]
        st["para_SpeedBias"][:, :, 3:9] = 0.0
        st["para_LegBias"][:] = LC_NOMINAL
        st["para_Ex_Pose"][:, :, 0:3] = tic_g
        st["para_Ex_Pose"][:, :, 3:7] = qic_g
        st["para_Td"][:] = 0.0
        nf = feats["lam"].shape[1]
        batch.para_Feature[:, :nf] = feats["lam"]
        if use_leg:
            batch.preint[:] = pre
        else:
            batch.use_imu_only()
            batch.imu_preint[:] = pre
        for w in range(B):
            off = 0
            fw, ow = batch.features[w], batch.obs[w]
            for f in range(nf):
                s, L = int(feats["start"][w, f]), int(feats["length"][w, f])
                fw[f]["start_frame"], fw[f]["n_obs"], fw[f]["obs_offset"] = s, L, off
                o = ow[off:off + L]
                o["point"] = feats["uv"][w, f, s:s + L, 0]
                o["velocity"] = feats["vel"][w, f, s:s + L, 0]
                o["pointRight"] = feats["uv"][w, f, s:s + L, 1]
                o["velocityRight"] = feats["vel"][w, f, s:s + L, 1]
                o["cur_td"] = 0.0
                o["is_stereo"] = feats["stereo"][w, f, s:s + L]
                off += L
            d = batch.descs[w]
            d.n_features, d.n_obs = nf, off
            d.extrinsic_open, d.td_open = 1, 0

    # ---- preintegration of the 11 intervals (-1->0 ... 9->10) through the backend -----------------------
    samples = np.zeros((B, NF - 1, S), dtype=abi.sample_dtype)
    jobs = (abi.PreintJob * (B * (NF - 1)))()
    for i in range(NF - 1):
        ks = i * S + 1 + np.arange(S)
        samples["dt"][:, i] = dt
        samples["acc"][:, i] = acc[:, ks]
        samples["gyr"][:, i] = gyr[:, ks]
        samples["phi"][:, i] = phi_m[:, ks].reshape(B, S, 12)
        samples["dphi"][:, i] = dphi_m[:, ks].reshape(B, S, 12)
        samples["c"][:, i] = contact[:, ks]
    for w in range(B):
        for i in range(NF - 1):
            j = jobs[w * (NF - 1) + i]
            k0 = i * S
            j.acc_0[:] = acc[w, k0]
            j.gyr_0[:] = gyr[w, k0]
            j.phi_0[:] = phi_m[w, k0].reshape(12)
            j.dphi_0[:] = dphi_m[w, k0].reshape(12)
            j.c_0[:] = contact[w, k0]
            j.linearized_ba[:] = (0.0, 0.0, 0.0)
            j.linearized_bg[:] = (0.0, 0.0, 0.0)
            j.linearized_rho[:] = (LC_NOMINAL,) * 4
            j.n_samples = S
            j.samples = samples[w, i].ctypes.data_as(C.POINTER(abi.IMULegSample))
    if use_leg:
        pre_all = backend.preintegrate(pcfg, jobs, B * (NF - 1)).reshape(B, NF - 1)
    else:
        pre_all = backend.preintegrate_imu(pcfg, jobs, B * (NF - 1)).reshape(B, NF - 1)
        with_prior = False

    # ---- the window itself: frames 0..10 (index 1..11) ------------------------------------------------------
    batch = abi.WindowBatch(B, max(F, 1))
    feats = make_features(F, 1, 11, "main")
    fill_batch(batch, feats, 1, pre_all[:, 1:11])

    # ---- prior: marginalize frame -1 out of the window made of frames -1..9 -------------------------------------
    if with_prior:
        F0 = prior_features if prior_features is not None else min(F, 40)
        pb = abi.WindowBatch(B, max(F0, 1))
        pfeats = make_features(F0, 0, 11, "prior")
        fill_batch(pb, pfeats, 0, pre_all[:, 0:10])
        backend.marginalize(cfg, pb, batch, margin_old=True)
        batch.prior_window = pb        # the previous window (frames -1..9), for tests that chain solve -> marginalize -> solve
    if return_truth:
        tr_out = SynthTruth()
        tr_out.R, tr_out.p, tr_out.v = R_f[:, 1:], p_f[:, 1:], v_f[:, 1:]
        tr_out.ba, tr_out.bg, tr_out.lc = ba, bg, lc_true
        tr_out.lam = feats["lam_true"]
        tr_out.raw_jobs, tr_out.raw_samples = jobs, samples
        return batch, tr_out
    batch._keepalive = (samples, jobs)
    return batch


def tile_batch(batch, n):
    """Replicate the windows of `batch` cyclically into a new WindowBatch of n windows (distinct storage)."""
    big = abi.WindowBatch(n, batch.max_features, batch.max_obs)
    src_n = batch.n
    reps = n // src_n + 1
    for arr in ("features", "obs", "preint", "prior_J", "prior_r", "para_Feature"):
        a = getattr(batch, arr)
        getattr(big, arr)[:] = np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:n]
    for w in range(n):
        s = w % src_n
        C.memmove(C.byref(big.states[w]), C.byref(batch.states[s]), C.sizeof(abi.WindowState))
        big.states[w].para_Feature = big.para_Feature[w].ctypes.data_as(abi.c_dp)
        d, sd = big.descs[w], batch.descs[s]
        d.n_features, d.n_obs, d.extrinsic_open, d.td_open = sd.n_features, sd.n_obs, sd.extrinsic_open, sd.td_open
        C.memmove(C.byref(d.prior), C.byref(sd.prior), C.sizeof(abi.Prior))
        d.prior.linearized_jacobians = big.prior_J[w].ctypes.data_as(abi.c_dp)
        d.prior.linearized_residuals = big.prior_r[w].ctypes.data_as(abi.c_dp)
    return big


# ----------------------------------------------------------------------------- sequences (replay harness)
class SynthSequence:
    """Raw sensor streams of `n` independent synthetic robots over `n_frames` camera frames (what the reference's frontend hands to
    Estimator::inputIMU / inputLeg / inputFeature): per inter-frame interval 33 IMU + leg samples, per frame the tracked features
    (id -> normalised point / pixel velocity per camera).  See generate_sequence()."""
    pass


def generate_sequence(n, n_frames, tracked=60, seed0=7000, stereo_prob=0.9, outlier_fraction=0.0, min_len=2, max_len=14, speed=0.5, yaw_rate=0.3):
    """n robots, n_frames frames (>= 12).  Features: every frame spawns enough new landmarks to keep ~`tracked` alive; a landmark is seen
    for L in [min_len, max_len] consecutive frames (tracks shorter than 4 never enter the solve, like in the reference), in the left
    camera always and in the right one with probability stereo_prob per observation.
    Returns a SynthSequence with
      samples [n, n_frames - 1, 33] sample_dtype, first [n, n_frames] (acc, gyr, phi, dphi, c at the frame instants: the integrators' ctor args),
      images: list over frames of list over robots of dict(ids, pts0 [k,7], has1 [k], pts1 [k,7]) with the reference's 7-vector (x, y, z=1, u, v, vx, vy),
      truth R/p/v [n, n_frames, ...], initial guesses p_g, R_g, v_g of every frame, ric_g / tic_g, ba, bg, lc_true."""
    pcfg = abi.default_preint_config()
    B, NF, S = n, n_frames, SAMPLES_PER_FRAME
    T = (NF - 1) * S + 1
    dt = FRAME_DT / S
    t = np.arange(T) * dt
    rngs = [np.random.Generator(np.random.PCG64(seed0 + w)) for w in range(B)]

    def draw(fn):
        return np.stack([fn(r) for r in rngs])

    par = {
        "yaw0": draw(lambda r: r.uniform(-np.pi, np.pi, 1)), "wz": draw(lambda r: r.uniform(-yaw_rate, yaw_rate, 1)), "sp": speed,
        "ph_r": draw(lambda r: r.uniform(0, 2 * np.pi, 1)), "ph_p": draw(lambda r: r.uniform(0, 2 * np.pi, 1)),
        "ph_z": draw(lambda r: r.uniform(0, 2 * np.pi, 1)),
        "x0": draw(lambda r: r.uniform(-5, 5, 1)), "y0": draw(lambda r: r.uniform(-5, 5, 1)), "z0": draw(lambda r: r.uniform(0.25, 0.35, 1)),
    }
    tr = _trajectory(par, t)
    ba = draw(lambda r: r.normal(0, 0.05, 3))
    bg = draw(lambda r: r.normal(0, 0.005, 3))
    gvec = np.array([0.0, 0.0, G_NORM])
    RT = np.swapaxes(tr["R"], -1, -2)
    acc = np.einsum("btij,btj->bti", RT, tr["a"] + gvec) + ba[:, None, :] + draw(lambda r: r.normal(0, 0.1, (T, 3)))
    gyr = tr["w"] + bg[:, None, :] + draw(lambda r: r.normal(0, 0.01, (T, 3)))
    fix = np.array([[pcfg.rho_fix[l][k] for k in range(4)] for l in range(4)])
    lc_true = LC_NOMINAL + draw(lambda r: r.normal(0, 0.005, 4))
    phase0 = draw(lambda r: r.uniform(0, 0.5, 1))
    gait = (np.floor((t[None, :] + phase0) / 0.25).astype(int) % 2)
    contact = np.stack([gait == 0, gait == 1, gait == 1, gait == 0], axis=-1).astype(float)
    phi = np.zeros((B, T, 4, 3)); dphi = np.zeros((B, T, 4, 3))
    phi[:, 0] = PHI_NOMINAL + draw(lambda r: r.normal(0, 0.05, (4, 3)))
    v_body = np.einsum("btij,btj->bti", RT, tr["v"])

    def joint_rate(ph, k):
        f = a1_fk(ph, lc_true, fix); J = a1_jac(ph, lc_true, fix)
        rhs = v_body[:, k, None, :] + np.cross(tr["w"][:, k, None, :], f)
        st = -np.linalg.solve(J, rhs[..., None])[..., 0]
        sw = 4.0 * (PHI_NOMINAL - ph)
        c = contact[:, k, :, None]
        return c * st + (1 - c) * sw

    for k in range(T - 1):
        k1 = joint_rate(phi[:, k], k); dphi[:, k] = k1
        mid = phi[:, k] + 0.5 * dt * k1
        phi[:, k + 1] = phi[:, k] + dt * 0.5 * (joint_rate(mid, k) + joint_rate(mid, k + 1))
    dphi[:, T - 1] = joint_rate(phi[:, T - 1], T - 1)
    phi_m = phi + draw(lambda r: r.normal(0, 1e-4, (T, 4, 3)))
    dphi_m = dphi + draw(lambda r: r.normal(0, 0.02, (T, 4, 3)))

    seq = SynthSequence()
    seq.n, seq.n_frames, seq.dt = B, NF, dt
    fidx = np.arange(NF) * S
    seq.R, seq.p, seq.v = tr["R"][:, fidx], tr["p"][:, fidx], tr["v"][:, fidx]
    seq.ba, seq.bg, seq.lc_true = ba, bg, lc_true
    seq.R_g = seq.R @ so3_exp(draw(lambda r: r.normal(0, 0.01, (NF, 3))))
    seq.p_g = seq.p + draw(lambda r: r.normal(0, 0.02, (NF, 3)))
    seq.v_g = seq.v + draw(lambda r: r.normal(0, 0.05, (NF, 3)))
    seq.ric_g = RIC @ so3_exp(draw(lambda r: r.normal(0, 0.005, (2, 3))))
    seq.tic_g = TIC + draw(lambda r: r.normal(0, 0.005, (2, 3)))
    samples = np.zeros((B, NF - 1, S), dtype=abi.sample_dtype)
    first = np.zeros((B, NF), dtype=abi.sample_dtype)
    for i in range(NF - 1):
        ks = i * S + 1 + np.arange(S)
        samples["dt"][:, i] = dt
        samples["acc"][:, i] = acc[:, ks]; samples["gyr"][:, i] = gyr[:, ks]
        samples["phi"][:, i] = phi_m[:, ks].reshape(B, S, 12); samples["dphi"][:, i] = dphi_m[:, ks].reshape(B, S, 12)
        samples["c"][:, i] = contact[:, ks]
    first["acc"], first["gyr"] = acc[:, fidx], gyr[:, fidx]
    first["phi"], first["dphi"], first["c"] = phi_m[:, fidx].reshape(B, NF, 12), dphi_m[:, fidx].reshape(B, NF, 12), contact[:, fidx]
    seq.samples, seq.first = samples, first

    # ---- landmarks: per robot, per frame new tracks ---------------------------------------------------------------------------
    images = [[None] * B for _ in range(NF)]
    for w in range(B):
        r = rngs[w]
        tracks = []                                    # (id, start, L, p_world)
        alive_until = []
        next_id = 0
        for k in range(NF):
            alive = sum(1 for e in alive_until if e > k)
            need = tracked - alive
            for _ in range(max(need, 0)):
                L = int(r.integers(min_len, max_len + 1))
                depth = r.uniform(2, 15); nx = r.uniform(-0.6, 0.6); ny = r.uniform(-0.45, 0.45)
                pc = np.array([nx * depth, ny * depth, depth])
                pw = seq.R[w, k] @ (RIC @ pc + TIC[0]) + seq.p[w, k]
                tracks.append((next_id, k, L, pw)); alive_until.append(k + L); next_id += 1
        per_frame = [[] for _ in range(NF)]
        prev_uv = {}
        for (fid, s0, L, pw) in tracks:
            for k in range(s0, min(s0 + L, NF)):
                pb = seq.R[w, k].T @ (pw - seq.p[w, k])
                uv = np.zeros((2, 2)); ok = True
                for cam in range(2):
                    pcam = RIC.T @ (pb - TIC[cam])
                    if pcam[2] < 0.3: ok = False
                    uv[cam] = pcam[:2] / pcam[2]
                if not ok or abs(uv[0, 0]) > 1.2 or abs(uv[0, 1]) > 0.9:
                    break                                                   # the track ends when the point leaves the view
                uv += r.normal(0, 0.5 / 460.0, (2, 2))
                if outlier_fraction > 0 and r.uniform() < outlier_fraction: uv += r.normal(0, 20.0 / 460.0, (2, 2))
                has1 = r.uniform() < stereo_prob
                pu = prev_uv.get(fid)
                vel = (uv - pu) / FRAME_DT if pu is not None else np.zeros((2, 2))
                prev_uv[fid] = uv
                per_frame[k].append((fid, uv, vel, has1))
        for k in range(NF):
            lst = per_frame[k]
            m = len(lst)
            ids = np.array([e[0] for e in lst], dtype=np.int64)
            pts0, pts1 = np.zeros((m, 7)), np.zeros((m, 7))
            has1 = np.array([e[3] for e in lst], dtype=bool)
            for q, (fid, uv, vel, h1) in enumerate(lst):
                pts0[q] = (uv[0, 0], uv[0, 1], 1.0, 460.0 * uv[0, 0] + 320, 460.0 * uv[0, 1] + 240, vel[0, 0], vel[0, 1])
                pts1[q] = (uv[1, 0], uv[1, 1], 1.0, 460.0 * uv[1, 0] + 320, 460.0 * uv[1, 1] + 240, vel[1, 0], vel[1, 1])
            images[k][w] = dict(ids=ids, pts0=pts0, has1=has1, pts1=pts1)
    seq.images = images
    return seq
