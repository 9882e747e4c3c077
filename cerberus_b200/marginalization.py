"""Marginalization half of Estimator::optimization() (reference src/estimator/estimator.cpp:1247-1456 +
src/factor/marginalization_factor.cpp:12-333), host glue over the device factor evaluators.

The factor Jacobians at the solved state (ResidualBlockInfo::Evaluate) come from the sm_100a "kernel per
factor family" entry points (cerb_eval_projection / cerb_eval_imu_leg / cerb_eval_prior); the loss
corrector, the A = J^T J assembly, the eps = 1e-8 clamped eigen-Schur complement and the factoring of the
result into (linearized_jacobians, linearized_residuals) follow marginalization_factor.cpp:46-77,183-305
in numpy.  Moving this glue onto the device is the "next #1" row of SURVEY.md section 8(f).

Block order inside the new prior is fixed (the reference's order is that of an unordered_map keyed by
pointer value, i.e. arbitrary): dropped = [pose0, speedbias0, legbias0, lambdas...], kept = [pose k ...,
speedbias1, legbias1, ex0, ex1, td] restricted to the blocks that occur.
"""
import ctypes as C
import numpy as np
from . import abi

EPS = 1e-8
_LOCAL = {abi.BLOCK_POSE: 6, abi.BLOCK_SPEEDBIAS: 9, abi.BLOCK_LEGBIAS: 4, abi.BLOCK_EX_POSE: 6, abi.BLOCK_TD: 1}
_GLOBAL = {abi.BLOCK_POSE: 7, abi.BLOCK_SPEEDBIAS: 9, abi.BLOCK_LEGBIAS: 4, abi.BLOCK_EX_POSE: 7, abi.BLOCK_TD: 1}


def _huber_correct(res, jac, delta):
    """ResidualBlockInfo::Evaluate loss part (marginalization_factor.cpp:46-77) for HuberLoss: rho'' <= 0, so
    residuals and Jacobians are scaled by sqrt(rho')."""
    s = (res * res).sum(axis=1)
    w = np.where(s > delta * delta, np.sqrt(delta / np.sqrt(np.maximum(s, 1e-300))), 1.0)
    return res * w[:, None], jac * w[:, None]


def _state_block(st_row, kind, index):
    if kind == abi.BLOCK_POSE:
        return st_row["para_Pose"][index]
    if kind == abi.BLOCK_SPEEDBIAS:
        return st_row["para_SpeedBias"][index]
    if kind == abi.BLOCK_LEGBIAS:
        return st_row["para_LegBias"][index]
    if kind == abi.BLOCK_EX_POSE:
        return st_row["para_Ex_Pose"][index]
    return st_row["para_Td"]


def marginalize_batch(backend, cfg, src, dst, margin_old=True):
    """Marginalize every window of `src` (at its current states) and write the resulting prior, already address-
    shifted for the next window, into the descriptors of `dst` (dst may be src)."""
    B = src.n
    st = src.state_array()
    delta = cfg.huber_delta
    # ---- gather the projection factors anchored at frame 0 over the whole batch (MARGIN_OLD only) -------------
    per_kind = {0: [], 1: [], 2: []}
    if margin_old:
        for w in range(B):
            d = src.descs[w]
            nf = d.n_features
            ft = src.features[w][:nf]
            sel = np.nonzero(ft["start_frame"] == 0)[0]
            if sel.size == 0:
                continue
            ob = src.obs[w]
            L = ft["n_obs"][sel]
            off = ft["obs_offset"][sel]
            fidx = np.repeat(sel, L)                       # feature of every observation
            k = np.concatenate([np.arange(l) for l in L])   # frame of every observation (anchor is frame 0)
            oi = np.repeat(off, L)                          # anchor observation
            oj = oi + k
            stereo = ob["is_stereo"][oj] != 0
            for kind, mask in ((0, k > 0), (1, (k > 0) & stereo), (2, (k == 0) & stereo)):
                if mask.any():
                    per_kind[kind].append((w, fidx[mask], k[mask], oi[mask], oj[mask]))
    evals = {}
    for kind, items in per_kind.items():
        if not items:
            continue
        ws = np.concatenate([np.full(it[1].size, it[0]) for it in items])
        fidx = np.concatenate([it[1] for it in items]); kk = np.concatenate([it[2] for it in items])
        oi = np.concatenate([it[3] for it in items]); oj = np.concatenate([it[4] for it in items])
        obs_i, obs_j = src.obs[ws, oi], src.obs[ws, oj]
        ones = np.ones((ws.size, 1))
        pts_i = np.concatenate([obs_i["point"], ones], axis=1)
        pj = obs_j["point"] if kind == 0 else obs_j["pointRight"]
        vj = obs_j["velocity"] if kind == 0 else obs_j["velocityRight"]
        res, jac = backend.eval_projection(kind, st["para_Pose"][ws, 0], st["para_Pose"][ws, kk], st["para_Ex_Pose"][ws, 0], st["para_Ex_Pose"][ws, 1],
                                           src.para_Feature[ws, fidx], st["para_Td"][ws, 0], pts_i, np.concatenate([pj, ones], axis=1),
                                           obs_i["velocity"], vj, obs_i["cur_td"], obs_j["cur_td"])
        res, jac = _huber_correct(res, jac, delta)
        evals[kind] = (ws, fidx, kk, res, jac)
    # ---- IMU-leg factor between frames 0 and 1 ------------------------------------------------------------------
    imu = None
    if margin_old:
        params = np.concatenate([st["para_Pose"][:, 0], st["para_SpeedBias"][:, 0], st["para_LegBias"][:, 0],
                                 st["para_Pose"][:, 1], st["para_SpeedBias"][:, 1], st["para_LegBias"][:, 1]], axis=1)
        pre0 = np.ascontiguousarray(src.preint[:, 0])
        r_imu, j_imu, _ = backend.eval_imu_leg(pre0, params)
        imu = (r_imu, j_imu.reshape(B, -1))

    for w in range(B):
        d = src.descs[w]
        rows_J, rows_r = [], []       # list of (residual vector, [(block key, jac [nr, local])])
        blocks_seen = []              # insertion order of (kind, index) / ('f', feature)

        def touch(key):
            if key not in blocks_seen:
                blocks_seen.append(key)

        drop = set()
        # old prior
        if d.prior.valid:
            pr = d.prior
            ncols = sum(_GLOBAL[pr.block_kind[b]] for b in range(pr.num_blocks))
            r_p, j_p = backend.eval_prior(pr, src.states[w], ncols)
            blks, off = [], 0
            for b in range(pr.num_blocks):
                kind, index = pr.block_kind[b], pr.block_index[b]
                g = _GLOBAL[kind]
                Jb = j_p[off:off + pr.n * g].reshape(pr.n, g)[:, :_LOCAL[kind]]
                off += pr.n * g
                key = (kind, index)
                touch(key); blks.append((key, Jb))
                if margin_old and index == 0 and kind in (abi.BLOCK_POSE, abi.BLOCK_SPEEDBIAS, abi.BLOCK_LEGBIAS):
                    drop.add(key)
                if (not margin_old) and kind == abi.BLOCK_POSE and index == abi.WINDOW_SIZE - 1:
                    drop.add(key)
            rows_r.append(r_p); rows_J.append(blks)
        if not margin_old:
            if not any(k == (abi.BLOCK_POSE, abi.WINDOW_SIZE - 1) for k in blocks_seen):
                if dst is not src:      # prior carried over unchanged (estimator.cpp:1380-1381)
                    _copy_prior(src, w, dst, w)
                continue
        if margin_old and imu is not None and src.preint[w, 0]["sum_dt"] < 10.0:
            Jf = imu[1][w].reshape(-1)
            offs = [(0, 7, (abi.BLOCK_POSE, 0)), (7, 9, (abi.BLOCK_SPEEDBIAS, 0)), (16, 4, (abi.BLOCK_LEGBIAS, 0)),
                    (20, 7, (abi.BLOCK_POSE, 1)), (27, 9, (abi.BLOCK_SPEEDBIAS, 1)), (36, 4, (abi.BLOCK_LEGBIAS, 1))]
            blks = []
            for (o, g, key) in offs:
                Jb = Jf[31 * o:31 * (o + g)].reshape(31, g)[:, :_LOCAL[key[0]]]
                touch(key); blks.append((key, Jb))
            drop.update([(abi.BLOCK_POSE, 0), (abi.BLOCK_SPEEDBIAS, 0), (abi.BLOCK_LEGBIAS, 0)])
            rows_r.append(imu[0][w]); rows_J.append(blks)
        for kind in (0, 1, 2):
            if kind not in evals:
                continue
            ws, fidx, kk, res, jac = evals[kind]
            m = np.nonzero(ws == w)[0]
            for t in m:
                f, j = int(fidx[t]), int(kk[t])
                J = jac[t]
                fk = ("f", f)
                if kind == 0:
                    blks = [((abi.BLOCK_POSE, 0), J[0:14].reshape(2, 7)[:, :6]), ((abi.BLOCK_POSE, j), J[14:28].reshape(2, 7)[:, :6]),
                            ((abi.BLOCK_EX_POSE, 0), J[28:42].reshape(2, 7)[:, :6]), (fk, J[42:44].reshape(2, 1)), ((abi.BLOCK_TD, 0), J[44:46].reshape(2, 1))]
                    drop.add((abi.BLOCK_POSE, 0))
                elif kind == 1:
                    blks = [((abi.BLOCK_POSE, 0), J[0:14].reshape(2, 7)[:, :6]), ((abi.BLOCK_POSE, j), J[14:28].reshape(2, 7)[:, :6]),
                            ((abi.BLOCK_EX_POSE, 0), J[28:42].reshape(2, 7)[:, :6]), ((abi.BLOCK_EX_POSE, 1), J[42:56].reshape(2, 7)[:, :6]),
                            (fk, J[56:58].reshape(2, 1)), ((abi.BLOCK_TD, 0), J[58:60].reshape(2, 1))]
                    drop.add((abi.BLOCK_POSE, 0))
                else:
                    blks = [((abi.BLOCK_EX_POSE, 0), J[0:14].reshape(2, 7)[:, :6]), ((abi.BLOCK_EX_POSE, 1), J[14:28].reshape(2, 7)[:, :6]),
                            (fk, J[28:30].reshape(2, 1)), ((abi.BLOCK_TD, 0), J[30:32].reshape(2, 1))]
                drop.add(fk)
                for key, _ in blks:
                    touch(key)
                rows_r.append(res[t]); rows_J.append(blks)
        dropped = [k for k in blocks_seen if k in drop]
        kept = [k for k in blocks_seen if k not in drop]
        if not dropped:
            dst.descs[w].prior.valid = 0       # MarginalizationInfo::valid = false (marginalization_factor.cpp:205-210)
            continue
        # canonical order: dropped = pose0, sb0, lb0, features ascending ; kept = poses ascending, sb, lb, ex0, ex1, td
        def kept_rank(k):
            order = {abi.BLOCK_POSE: 0, abi.BLOCK_SPEEDBIAS: 1, abi.BLOCK_LEGBIAS: 2, abi.BLOCK_EX_POSE: 3, abi.BLOCK_TD: 4}
            return (order[k[0]], k[1])
        def drop_rank(k):
            return (1, k[1]) if k[0] == "f" else (0, {abi.BLOCK_POSE: 0, abi.BLOCK_SPEEDBIAS: 1, abi.BLOCK_LEGBIAS: 2}[k[0]] + 10 * k[1])
        dropped.sort(key=drop_rank); kept.sort(key=kept_rank)
        size = lambda k: 1 if k[0] == "f" else _LOCAL[k[0]]
        idx, pos = {}, 0
        for k in dropped + kept:
            idx[k] = pos; pos += size(k)
        m = sum(size(k) for k in dropped); n = pos - m
        R = sum(r.shape[0] for r in rows_r)
        Jbig, rbig = np.zeros((R, pos)), np.concatenate(rows_r)
        ro = 0
        for r, blks in zip(rows_r, rows_J):
            for key, Jb in blks:
                Jbig[ro:ro + r.shape[0], idx[key]:idx[key] + Jb.shape[1]] += Jb
            ro += r.shape[0]
        A = Jbig.T @ Jbig
        b = Jbig.T @ rbig
        # marginalization_factor.cpp:278-305
        Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
        ev, V = np.linalg.eigh(Amm)
        inv = np.where(ev > EPS, 1.0 / np.where(ev > EPS, ev, 1.0), 0.0)
        Amm_inv = (V * inv) @ V.T
        Arm = A[m:, :m]
        Ar = A[m:, m:] - Arm @ Amm_inv @ A[:m, m:]
        br = b[m:] - Arm @ Amm_inv @ b[:m]
        ev2, V2 = np.linalg.eigh(0.5 * (Ar + Ar.T))
        S = np.where(ev2 > EPS, ev2, 0.0)
        S_inv = np.where(ev2 > EPS, 1.0 / np.where(ev2 > EPS, ev2, 1.0), 0.0)
        lin_J = np.sqrt(S)[:, None] * V2.T
        lin_r = np.sqrt(S_inv) * (V2.T @ br)
        # getParameterBlocks + addr_shift (estimator.cpp:1357-1372 / :1413-1447)
        pr = dst.descs[w].prior
        x0s, metas = [], []
        for k in kept:
            kind, index = k
            x0 = np.array(_state_block(st[w], kind, index), dtype=np.float64).ravel().copy()
            if kind in (abi.BLOCK_POSE, abi.BLOCK_SPEEDBIAS, abi.BLOCK_LEGBIAS):
                index = index - 1 if margin_old else (index - 1 if index == abi.WINDOW_SIZE else index)
            metas.append((kind, index, idx[k] - m)); x0s.append(x0)
        pr.valid, pr.n, pr.num_blocks = 1, n, len(kept)
        for bi, ((kind, index, col), x0) in enumerate(zip(metas, x0s)):
            pr.block_kind[bi], pr.block_index[bi], pr.block_col[bi] = kind, index, col
            for t in range(7):
                pr.block_x0[bi][t] = x0[t] if t < x0.size else 0.0
        dst.prior_J[w, :n * n] = lin_J.T.ravel()         # column-major n x n
        dst.prior_r[w, :n] = lin_r
        pr.linearized_jacobians = dst.prior_J[w].ctypes.data_as(abi.c_dp)
        pr.linearized_residuals = dst.prior_r[w].ctypes.data_as(abi.c_dp)


def _copy_prior(src, ws, dst, wd):
    ps, pd = src.descs[ws].prior, dst.descs[wd].prior
    dst.prior_J[wd] = src.prior_J[ws]; dst.prior_r[wd] = src.prior_r[ws]
    pd.valid, pd.n, pd.num_blocks = ps.valid, ps.n, ps.num_blocks
    for b in range(abi.MAX_PRIOR_BLOCKS):
        pd.block_kind[b], pd.block_index[b], pd.block_col[b] = ps.block_kind[b], ps.block_index[b], ps.block_col[b]
        for t in range(7):
            pd.block_x0[b][t] = ps.block_x0[b][t]
    pd.linearized_jacobians = dst.prior_J[wd].ctypes.data_as(abi.c_dp)
    pd.linearized_residuals = dst.prior_r[wd].ctypes.data_as(abi.c_dp)
