"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (the CPU restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np
from cerberus_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        sizes = (C.c_int * 16)()
        k = _LIB.oracle_abi_sizes(sizes, 16)
        got = [sizes[i] for i in range(k)]
        want = [C.sizeof(s) for s in abi.ABI_STRUCTS]
        assert got == want, f"ABI struct size mismatch oracle={got} python={want}"
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(abi.c_dp)


class OracleBackend:
    """Implements the synth `backend` protocol and the solve/eval entry points with the CPU oracle."""

    def __init__(self, cfg=None):
        self.cfg = cfg or abi.default_config()
        self.lib = lib()

    def preintegrate(self, pcfg, jobs, n):
        out = np.zeros(n, dtype=abi.preint_dtype)
        rc = self.lib.oracle_preintegrate(C.byref(pcfg), n, jobs, out.ctypes.data_as(C.POINTER(abi.IMULegPreint)))
        assert rc == 0
        return out

    def marginalize(self, cfg, src, dst, margin_old=True):
        """margin_old: one flag for the batch, or one per window (True / 1: MARGIN_OLD, False / 0: MARGIN_SECOND_NEW)."""
        flags = np.broadcast_to(np.asarray(margin_old, dtype=bool), (src.n,))
        for w in range(src.n):
            pr = dst.descs[w].prior
            J, r = dst.prior_J[w], dst.prior_r[w]
            rc = self.lib.oracle_marginalize(C.byref(cfg), C.byref(src.descs[w]), C.byref(src.states[w]), 1 if flags[w] else 0,
                                             C.byref(pr), _p(J), _p(r))
            assert rc == 0

    def set_eig_mode(self, mode):
        """0: tridiagonal QR (the reference's Eigen::SelfAdjointEigenSolver; default), 1: cyclic Jacobi."""
        self.lib.oracle_set_eig_mode(int(mode))

    def solve_batch(self, batch, nthreads=1, cfg=None):
        cfg = cfg or self.cfg
        rc = self.lib.oracle_solve_batch(C.byref(cfg), batch.n, batch.descs, batch.states, batch.reports, nthreads)
        assert rc == 0
        return batch.report_array().copy()

    def solve_window(self, batch, w, want_probe=False, cfg=None):
        cfg = cfg or self.cfg
        nf = batch.descs[w].n_features
        g = np.zeros(abi.NUM_REDUCED + nf)
        d = np.zeros(abi.NUM_REDUCED + nf)
        rc = self.lib.oracle_solve_window(C.byref(cfg), C.byref(batch.descs[w]), C.byref(batch.states[w]), C.byref(batch.reports[w]),
                                          _p(g) if want_probe else None, _p(d) if want_probe else None, g.size)
        assert rc == 0
        return (g, d) if want_probe else None

    def eval_projection(self, kind, pose_i, pose_j, ex0, ex1, inv_dep, td, pts_i, pts_j, vel_i, vel_j, td_i, td_j, sqrt_info=460.0 / 1.5, want_jac=True):
        n = inv_dep.shape[0]
        res = np.zeros((n, 2))
        jac = np.zeros((n, abi.PROJ_JAC_SIZE[kind])) if want_jac else None
        self.lib.oracle_eval_projection.argtypes = [C.c_int, C.c_int, C.c_double] + [abi.c_dp] * 14
        rc = self.lib.oracle_eval_projection(kind, n, sqrt_info, _p(pose_i), _p(pose_j), _p(ex0), _p(ex1), _p(inv_dep), _p(td), _p(pts_i), _p(pts_j),
                                             _p(vel_i), _p(vel_j), _p(td_i), _p(td_j), _p(res), _p(jac))
        assert rc == 0
        return res, jac

    def eval_imu_leg(self, preint, params, g=(0.0, 0.0, 9.805), want_jac=True):
        n = params.shape[0]
        res = np.zeros((n, 31))
        jac = np.zeros((n, 31 * 40)) if want_jac else None
        si = np.zeros((n, 961))
        garr = np.array(g, dtype=np.float64)
        rc = self.lib.oracle_eval_imu_leg(n, _p(garr), preint.ctypes.data_as(C.POINTER(abi.IMULegPreint)), _p(params), _p(res), _p(jac), _p(si))
        assert rc == 0
        return res, jac, si

    def eval_imu(self, preint, params, g=(0.0, 0.0, 9.805), want_jac=True):
        n = params.shape[0]
        res, si = np.zeros((n, 15)), np.zeros((n, 225))
        jac = np.zeros((n, 15 * 32)) if want_jac else None
        garr = np.array(g, dtype=np.float64)
        rc = self.lib.oracle_eval_imu(n, _p(garr), preint.ctypes.data_as(C.POINTER(abi.IMUPreint)), _p(np.ascontiguousarray(params)), _p(res), _p(jac), _p(si))
        assert rc == 0
        return res, jac, si

    def preintegrate_imu(self, pcfg, jobs, n):
        out = np.zeros(n, dtype=abi.imu_preint_dtype)
        rc = self.lib.oracle_preintegrate_imu(C.byref(pcfg), n, jobs, out.ctypes.data_as(C.POINTER(abi.IMUPreint)))
        assert rc == 0
        return out

    def eval_prior(self, prior, state, n_cols):
        res = np.zeros(prior.n)
        jac = np.zeros(prior.n * n_cols)
        rc = self.lib.oracle_eval_prior(C.byref(prior), C.byref(state), _p(res), _p(jac))
        assert rc == 0
        return res, jac

    def a1_kinematics(self, q, rho_opt, rho_fix):
        n = q.shape[0]
        fk, jac, dfk, djq, djr = np.zeros((n, 3)), np.zeros((n, 9)), np.zeros((n, 3)), np.zeros((n, 27)), np.zeros((n, 9))
        rc = self.lib.oracle_a1_kinematics(n, _p(q), _p(rho_opt), _p(rho_fix), _p(fk), _p(jac), _p(dfk), _p(djq), _p(djr))
        assert rc == 0
        return fk, jac, dfk, djq, djr

    def double2vector(self, before_state, after_state):
        Ps, Rs, Vs = np.zeros((11, 3)), np.zeros((11, 3, 3)), np.zeros((11, 3))
        self.lib.oracle_double2vector(C.byref(before_state), C.byref(after_state), _p(Ps), _p(Rs), _p(Vs))
        return Ps, Rs, Vs

    def outlier_errors(self, batch):
        out = np.full((batch.n, batch.max_features), np.nan)
        for w in range(batch.n):
            assert self.lib.oracle_outlier_errors(C.byref(batch.descs[w]), C.byref(batch.states[w]), _p(out[w])) == 0
        return out

    def marginalize_schur(self, A, b, m, eps=1e-8):
        A = np.ascontiguousarray(A, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        B, pos = b.shape; n = pos - m
        J = np.zeros((B, n * n)); r = np.zeros((B, n))
        self.lib.oracle_marginalize_schur.argtypes = [C.c_int, C.c_int, abi.c_dp, abi.c_dp, C.c_double, abi.c_dp, abi.c_dp]
        for w in range(B):
            assert self.lib.oracle_marginalize_schur(m, n, _p(A[w]), _p(b[w]), eps, _p(J[w]), _p(r[w])) == 0
        return J.reshape(B, n, n).transpose(0, 2, 1), r

    def shift_depth(self, batch, init_depth=5.0):
        n, F = batch.n, batch.max_features
        start = np.full((n, F), -1, dtype=np.int32); depth = np.full((n, F), np.nan); keep = np.full((n, F), -1, dtype=np.int32)
        self.lib.oracle_shift_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_int32), abi.c_dp, C.POINTER(C.c_int32)]
        for w in range(n):
            assert self.lib.oracle_shift_depth(C.byref(batch.descs[w]), C.byref(batch.states[w]), init_depth, start[w].ctypes.data_as(C.POINTER(C.c_int32)),
                                               _p(depth[w]), keep[w].ctypes.data_as(C.POINTER(C.c_int32))) == 0
        return start, depth, keep

    def triangulate(self, batch, init_depth=5.0):
        out = np.full((batch.n, batch.max_features), np.nan)
        self.lib.oracle_triangulate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, abi.c_dp]
        for w in range(batch.n):
            assert self.lib.oracle_triangulate(C.byref(batch.descs[w]), C.byref(batch.states[w]), init_depth, _p(out[w])) == 0
        return out


# ---------------------------------------------------------------------------------------------------------------
# oracle/_ref: the reference's OWN factor sources compiled against the header shims (oracle/shim), `make -C oracle ref`
_REF = None


def ref_lib():
    global _REF
    if _REF is None:
        path = os.path.join(_ROOT, "oracle", "_ref", "libcerberus_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/src"):
                subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle"), "ref"])
            else:
                return None
        _REF = C.CDLL(path)
    return _REF


class RefBackend:
    """Factor evaluators / preintegration / kinematics of the compiled reference sources (no solver: Ceres is absent)."""

    def __init__(self, pcfg=None, g=(0.0, 0.0, 9.805), visual_sqrt_info=460.0 / 1.5):
        self.lib = ref_lib()
        if self.lib is None:
            raise RuntimeError("oracle/_ref/libcerberus_ref.so is not built and /root/reference is absent")
        self.set_globals(pcfg or abi.default_preint_config(), g, visual_sqrt_info)

    def set_globals(self, pcfg, g=(0.0, 0.0, 9.805), visual_sqrt_info=460.0 / 1.5):
        garr = (C.c_double * 3)(*g)
        self.lib.ref_set_globals.argtypes = [C.POINTER(abi.PreintConfig), C.POINTER(C.c_double), C.c_double]
        self.lib.ref_set_globals(C.byref(pcfg), garr, visual_sqrt_info)

    def eval_projection(self, kind, pose_i, pose_j, ex0, ex1, inv_dep, td, pts_i, pts_j, vel_i, vel_j, td_i, td_j, want_jac=True):
        n = inv_dep.shape[0]
        res = np.zeros((n, 2))
        jac = np.zeros((n, abi.PROJ_JAC_SIZE[kind])) if want_jac else None
        self.lib.ref_eval_projection(kind, n, _p(pose_i), _p(pose_j), _p(ex0), _p(ex1), _p(inv_dep), _p(td), _p(pts_i), _p(pts_j), _p(vel_i), _p(vel_j),
                                     _p(td_i), _p(td_j), _p(res), _p(jac))
        return res, jac

    def eval_imu_leg(self, preint, params, want_jac=True):
        n = params.shape[0]
        res, si = np.zeros((n, 31)), np.zeros((n, 961))
        jac = np.zeros((n, 31 * 40)) if want_jac else None
        self.lib.ref_eval_imu_leg(n, preint.ctypes.data_as(C.POINTER(abi.IMULegPreint)), _p(params), _p(res), _p(jac), _p(si))
        return res, jac, si

    def preintegrate(self, pcfg, jobs, n):
        self.set_globals(pcfg)
        out = np.zeros(n, dtype=abi.preint_dtype)
        self.lib.ref_preintegrate(n, jobs, out.ctypes.data_as(C.POINTER(abi.IMULegPreint)))
        return out

    def a1_kinematics(self, q, rho_opt, rho_fix):
        n = q.shape[0]
        fk, jac, dfk, djq, djr = np.zeros((n, 3)), np.zeros((n, 9)), np.zeros((n, 3)), np.zeros((n, 27)), np.zeros((n, 9))
        self.lib.ref_a1_kinematics(n, _p(q), _p(rho_opt), _p(rho_fix), _p(fk), _p(jac), _p(dfk), _p(djq), _p(djr))
        return fk, jac, dfk, djq, djr

    def pose_plus(self, x, delta):
        out = np.zeros(7)
        self.lib.ref_pose_plus(_p(np.ascontiguousarray(x)), _p(np.ascontiguousarray(delta)), _p(out))
        return out

    def eval_prior(self, prior, state, n_cols):
        res, jac = np.zeros(prior.n), np.zeros(prior.n * n_cols)
        self.lib.ref_eval_prior(C.byref(prior), C.byref(state), _p(res), _p(jac))
        return res, jac


def _ref_marginalize(self, cfg, src, dst, margin_old=True):
    """RefBackend.marginalize: the reference's own MarginalizationInfo / ResidualBlockInfo classes."""
    flags = np.broadcast_to(np.asarray(margin_old, dtype=bool), (src.n,))
    for w in range(src.n):
        pr = dst.descs[w].prior
        rc = self.lib.ref_marginalize(C.byref(src.descs[w]), C.byref(src.states[w]), 1 if flags[w] else 0, C.byref(pr), _p(dst.prior_J[w]), _p(dst.prior_r[w]))
        assert rc == 0


def _ref_double2vector(self, before_state, after_state):
    Ps, Rs, Vs = np.zeros((11, 3)), np.zeros((11, 3, 3)), np.zeros((11, 3))
    self.lib.ref_double2vector(C.byref(before_state), C.byref(after_state), _p(Ps), _p(Rs), _p(Vs))
    return Ps, Rs, Vs


def _ref_triangulate(self, batch, init_depth=5.0):
    out = np.full((batch.n, batch.max_features), np.nan)
    self.lib.ref_triangulate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, abi.c_dp]
    for w in range(batch.n):
        assert self.lib.ref_triangulate(C.byref(batch.descs[w]), C.byref(batch.states[w]), init_depth, _p(out[w])) == 0
    return out


def _ref_outlier_errors(self, batch):
    out = np.full((batch.n, batch.max_features), np.nan)
    for w in range(batch.n):
        assert self.lib.ref_outlier_errors(C.byref(batch.descs[w]), C.byref(batch.states[w]), _p(out[w])) == 0
    return out


def _ref_shift_depth(self, batch, init_depth=5.0):
    n, F = batch.n, batch.max_features
    start = np.full((n, F), -1, dtype=np.int32); depth = np.full((n, F), np.nan); keep = np.full((n, F), -1, dtype=np.int32)
    self.lib.ref_shift_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_int32), abi.c_dp, C.POINTER(C.c_int32)]
    for w in range(n):
        assert self.lib.ref_shift_depth(C.byref(batch.descs[w]), C.byref(batch.states[w]), init_depth, start[w].ctypes.data_as(C.POINTER(C.c_int32)),
                                        _p(depth[w]), keep[w].ctypes.data_as(C.POINTER(C.c_int32))) == 0
    return start, depth, keep


RefBackend.double2vector = _ref_double2vector
RefBackend.triangulate = _ref_triangulate
RefBackend.outlier_errors = _ref_outlier_errors
RefBackend.shift_depth = _ref_shift_depth


RefBackend.marginalize = _ref_marginalize


def _ref_eval_imu(self, preint, params, want_jac=True):
    n = params.shape[0]
    res, si = np.zeros((n, 15)), np.zeros((n, 225))
    jac = np.zeros((n, 15 * 32)) if want_jac else None
    self.lib.ref_eval_imu(n, preint.ctypes.data_as(C.POINTER(abi.IMUPreint)), _p(np.ascontiguousarray(params)), _p(res), _p(jac), _p(si))
    return res, jac, si


def _ref_preintegrate_imu(self, pcfg, jobs, n):
    self.set_globals(pcfg)
    out = np.zeros(n, dtype=abi.imu_preint_dtype)
    self.lib.ref_preintegrate_imu(n, jobs, out.ctypes.data_as(C.POINTER(abi.IMUPreint)))
    return out


RefBackend.eval_imu = _ref_eval_imu
RefBackend.preintegrate_imu = _ref_preintegrate_imu


class OracleOps:
    """Backend adapter of cerberus_b200.estimator.ReplayDriver on the CPU oracle (the reference arm of the replay comparison)."""

    def __init__(self, cfg, nthreads=8, eig_mode=0, marg=None, feat=None):
        self.o, self.cfg, self.nthreads, self.eig_mode = OracleBackend(cfg), cfg, nthreads, eig_mode
        self.marg = marg          # optional RefBackend: the reference's own MarginalizationInfo classes for the marginalization step
        self.feat = feat          # optional RefBackend: the reference's own FeatureManager for triangulation / depth shift / outlier errors

    def preintegrate(self, pcfg, jobs, n): return self.o.preintegrate(pcfg, jobs, n)
    def solve(self, batch): return self.o.solve_batch(batch, nthreads=self.nthreads)
    def double2vector(self, before, after): return self.o.double2vector(before, after)
    def triangulate(self, batch): return (self.feat or self.o).triangulate(batch)
    def outlier_errors(self, batch): return (self.feat or self.o).outlier_errors(batch)
    def shift_depth(self, batch): return (self.feat or self.o).shift_depth(batch)

    def marginalize(self, src, dst, flags):
        if self.marg is not None:
            self.marg.lib.ref_set_eigen_mode(int(self.eig_mode))
            self.marg.marginalize(self.cfg, src, dst, margin_old=(np.asarray(flags) == 0))
        else:
            self.o.set_eig_mode(self.eig_mode)
            self.o.marginalize(self.cfg, src, dst, margin_old=(np.asarray(flags) == 0))
