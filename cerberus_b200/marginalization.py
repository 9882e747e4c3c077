"""Marginalization half of Estimator::optimization() (reference src/estimator/estimator.cpp:1247-1456 +
src/factor/marginalization_factor.cpp:12-333), host glue over the device factor evaluators.

The factor Jacobians at the solved state (ResidualBlockInfo::Evaluate) come from the sm_100a "kernel per
factor family" entry points (cerb_eval_projection / cerb_eval_imu_leg / cerb_eval_prior); the eps = 1e-8
clamped eigen-Schur complement and the factoring of the result into (linearized_jacobians,
linearized_residuals) (marginalization_factor.cpp:281-305) run on the device too (cerb_marginalize_schur,
csrc/marg_kernels.cuh).  What is left in numpy here is the bookkeeping in between: the loss corrector and the
A = J^T J, b = J^T r assembly (marginalization_factor.cpp:46-77,150-181) -- moving that onto the device (the solve
kernel's linearisation restricted to the factors that touch frame 0) is the rest of the "next #1" row of SURVEY.md 8(f).

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
    flags = np.broadcast_to(np.asarray(margin_old, dtype=bool), (B,))        # one flag for the batch or one per window (True: MARGIN_OLD)
    any_old = bool(flags.any())
    # ---- gather the projection factors anchored at frame 0 over the whole batch (MARGIN_OLD only) -------------
    per_kind = {0: [], 1: [], 2: []}
    if any_old:
        for w in range(B):
            if not flags[w]:
                continue
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
    if any_old:
        params = np.concatenate([st["para_Pose"][:, 0], st["para_SpeedBias"][:, 0], st["para_LegBias"][:, 0],
                                 st["para_Pose"][:, 1], st["para_SpeedBias"][:, 1], st["para_LegBias"][:, 1]], axis=1)
        pre0 = np.ascontiguousarray(src.preint[:, 0])
        r_imu, j_imu, _ = backend.eval_imu_leg(pre0, params)
        imu = (r_imu, j_imu.reshape(B, -1))

    if flags.all() and _uniform_structure(src) and not any(src.descs[w].prior.valid for w in range(B)) and imu is not None \
            and (src.preint[:, 0]["sum_dt"] < 10.0).all() and all(k in evals for k in (0, 1, 2)):
        _marginalize_uniform(backend, cfg, src, dst, st, evals, imu)
        return

    for w in range(B):
        d = src.descs[w]
        margin_old = bool(flags[w])
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
        # marginalization_factor.cpp:281-305 on the device (csrc/marg_kernels.cuh)
        lin_J, lin_r = backend.marginalize_schur(A[None], b[None], m, EPS)
        lin_J, lin_r = lin_J[0], lin_r[0]
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
            for t in range(9):
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
        for t in range(9):
            pd.block_x0[b][t] = ps.block_x0[b][t]
    pd.linearized_jacobians = dst.prior_J[wd].ctypes.data_as(abi.c_dp)
    pd.linearized_residuals = dst.prior_r[wd].ctypes.data_as(abi.c_dp)


def _uniform_structure(src):
    """True if every window has the same feature tracks / stereo flags (the dense synthetic configuration)."""
    nf = src.descs[0].n_features
    no = src.descs[0].n_obs
    if any(src.descs[w].n_features != nf or src.descs[w].n_obs != no for w in range(src.n)):
        return False
    f0 = src.features[0][:nf]
    for name in ("start_frame", "n_obs", "obs_offset"):
        if not (src.features[:, :nf][name] == f0[name]).all():
            return False
    return bool((src.obs[:, :no]["is_stereo"] == src.obs[0, :no]["is_stereo"]).all())


def _marginalize_uniform(backend, cfg, src, dst, st, evals, imu, chunk=64):
    """Vectorised MARGIN_OLD for batches whose windows all share one factor-graph structure and carry no prior."""
    B = src.n
    nf = src.descs[0].n_features
    sel = np.nonzero(src.features[0][:nf]["start_frame"] == 0)[0]
    fpos = {int(f): i for i, f in enumerate(sel)}
    m = 19 + sel.size
    # kept layout: pose k (k = 1..10) -> 6 (k - 1); speedbias1 -> 60; legbias1 -> 69; ex0 -> 73; ex1 -> 79; td -> 85
    KP = lambda k: m + 6 * (k - 1)
    K_SB, K_LB, K_E0, K_E1, K_TD = m + 60, m + 69, m + 73, m + 79, m + 85
    pos, n = m + 86, 86
    # column index templates of window 0 (identical for all windows)
    def per_window(kind):
        ws, fidx, kk, res, jac = evals[kind]
        cnt = int((ws == 0).sum())
        return fidx[:cnt], kk[:cnt], res.reshape(B, cnt, 2), jac.reshape(B, cnt, -1), cnt
    f1, k1, r1, j1, n1 = per_window(0)
    f2, k2, r2, j2, n2 = per_window(1)
    f3, k3, r3, j3, n3 = per_window(2)
    R = 2 * (n1 + n2 + n3) + 31
    lam1 = np.array([19 + fpos[int(f)] for f in f1]); lam2 = np.array([19 + fpos[int(f)] for f in f2]); lam3 = np.array([19 + fpos[int(f)] for f in f3])
    ar6 = np.arange(6)
    out_kinds = [(abi.BLOCK_POSE, k) for k in range(1, 11)] + [(abi.BLOCK_SPEEDBIAS, 1), (abi.BLOCK_LEGBIAS, 1), (abi.BLOCK_EX_POSE, 0), (abi.BLOCK_EX_POSE, 1), (abi.BLOCK_TD, 0)]
    out_cols = [6 * (k - 1) for k in range(1, 11)] + [60, 69, 73, 79, 85]
    for c0 in range(0, B, chunk):
        c1 = min(B, c0 + chunk); nb = c1 - c0
        J = np.zeros((nb, R, pos)); r = np.zeros((nb, R))
        ro = 0
        def put(rows, cols, block):      # rows [cnt,2], cols [cnt,w], block [nb,cnt,2,w]
            J[:, rows[:, :, None], cols[:, None, :]] += block
        # K1: blocks pose0 | pose_j | ex0 | lambda | td
        rows = ro + 2 * np.arange(n1)[:, None] + np.arange(2)[None, :]
        b = j1[c0:c1]
        put(rows, np.tile(ar6, (n1, 1)), b[:, :, 0:14].reshape(nb, n1, 2, 7)[..., :6])
        put(rows, KP(k1)[:, None] + ar6[None, :], b[:, :, 14:28].reshape(nb, n1, 2, 7)[..., :6])
        put(rows, np.tile(K_E0 + ar6, (n1, 1)), b[:, :, 28:42].reshape(nb, n1, 2, 7)[..., :6])
        put(rows, lam1[:, None], b[:, :, 42:44].reshape(nb, n1, 2, 1))
        put(rows, np.full((n1, 1), K_TD), b[:, :, 44:46].reshape(nb, n1, 2, 1))
        r[:, ro:ro + 2 * n1] = r1[c0:c1].reshape(nb, -1); ro += 2 * n1
        # K2: pose0 | pose_j | ex0 | ex1 | lambda | td
        rows = ro + 2 * np.arange(n2)[:, None] + np.arange(2)[None, :]
        b = j2[c0:c1]
        put(rows, np.tile(ar6, (n2, 1)), b[:, :, 0:14].reshape(nb, n2, 2, 7)[..., :6])
        put(rows, KP(k2)[:, None] + ar6[None, :], b[:, :, 14:28].reshape(nb, n2, 2, 7)[..., :6])
        put(rows, np.tile(K_E0 + ar6, (n2, 1)), b[:, :, 28:42].reshape(nb, n2, 2, 7)[..., :6])
        put(rows, np.tile(K_E1 + ar6, (n2, 1)), b[:, :, 42:56].reshape(nb, n2, 2, 7)[..., :6])
        put(rows, lam2[:, None], b[:, :, 56:58].reshape(nb, n2, 2, 1))
        put(rows, np.full((n2, 1), K_TD), b[:, :, 58:60].reshape(nb, n2, 2, 1))
        r[:, ro:ro + 2 * n2] = r2[c0:c1].reshape(nb, -1); ro += 2 * n2
        # K3: ex0 | ex1 | lambda | td
        rows = ro + 2 * np.arange(n3)[:, None] + np.arange(2)[None, :]
        b = j3[c0:c1]
        put(rows, np.tile(K_E0 + ar6, (n3, 1)), b[:, :, 0:14].reshape(nb, n3, 2, 7)[..., :6])
        put(rows, np.tile(K_E1 + ar6, (n3, 1)), b[:, :, 14:28].reshape(nb, n3, 2, 7)[..., :6])
        put(rows, lam3[:, None], b[:, :, 28:30].reshape(nb, n3, 2, 1))
        put(rows, np.full((n3, 1), K_TD), b[:, :, 30:32].reshape(nb, n3, 2, 1))
        r[:, ro:ro + 2 * n3] = r3[c0:c1].reshape(nb, -1); ro += 2 * n3
        # IMU-leg factor: pose0 | sb0 | lb0 | pose1 | sb1 | lb1
        ji = imu[1][c0:c1]
        for (o, g, col, loc) in ((0, 7, 0, 6), (7, 9, 6, 9), (16, 4, 15, 4), (20, 7, KP(1), 6), (27, 9, K_SB, 9), (36, 4, K_LB, 4)):
            J[:, ro:ro + 31, col:col + loc] = ji[:, 31 * o:31 * (o + g)].reshape(nb, 31, g)[:, :, :loc]
        r[:, ro:ro + 31] = imu[0][c0:c1]
        A = np.einsum("bri,brj->bij", J, J, optimize=True)
        bv = np.einsum("bri,br->bi", J, r)
        lin_J, lin_r = backend.marginalize_schur(A, bv, m, EPS)      # marginalization_factor.cpp:281-305 on the device
        for i in range(nb):
            w = c0 + i
            pr = dst.descs[w].prior
            pr.valid, pr.n, pr.num_blocks = 1, n, len(out_kinds)
            for bi, ((kind, index), col) in enumerate(zip(out_kinds, out_cols)):
                x0 = np.array(_state_block(st[w], kind, index), dtype=np.float64).ravel()
                if kind in (abi.BLOCK_POSE, abi.BLOCK_SPEEDBIAS, abi.BLOCK_LEGBIAS):
                    index -= 1
                pr.block_kind[bi], pr.block_index[bi], pr.block_col[bi] = kind, index, col
                for t in range(9):
                    pr.block_x0[bi][t] = x0[t] if t < x0.size else 0.0
            dst.prior_J[w, :n * n] = lin_J[i].T.ravel()
            dst.prior_r[w, :n] = lin_r[i]
            pr.linearized_jacobians = dst.prior_J[w].ctypes.data_as(abi.c_dp)
            pr.linearized_residuals = dst.prior_r[w].ctypes.data_as(abi.c_dp)
