// factors.cuh -- per-factor device math of the Cerberus factor families (fp64), shared by the batched
// "one kernel per factor family" evaluators (eval_kernels.cuh) and the fused window solver
// (solve_kernel.cuh).  Written against rotation matrices staged in shared memory rather than the
// quaternion-per-call style of the reference; the formulas are those of
//   src/factor/projectionTwoFrameOneCamFactor.cpp:43-150   (K1)
//   src/factor/projectionTwoFrameTwoCamFactor.cpp:43-166   (K2)
//   src/factor/projectionOneFrameTwoCamFactor.cpp:42-134   (K3, incl. the pts_i quirk at :119)
//   src/factor/imu_leg_factor.cpp:173-386 + imu_leg_integration_base.cpp:845-898   (K5)
//   src/legKinematics/A1Kinematics.cpp:43-220
#pragma once
#include "vmath.cuh"

namespace cerb {

enum { PROJ_K1 = 0, PROJ_K2 = 1, PROJ_K3 = 2 };

// ---- compact device layout of one IMULegIntegrationBase result (see DESIGN.md "HBM layout") --------
enum {
    PRE_SUM_DT = 0, PRE_DP = 1, PRE_DQ = 4, PRE_DV = 8, PRE_DEPS = 11, PRE_BA = 23, PRE_BG = 26, PRE_RHO = 29,
    PRE_DP_DBA = 33, PRE_DP_DBG = 42, PRE_DQ_DBG = 51, PRE_DV_DBA = 60, PRE_DV_DBG = 69, PRE_DEP_DBG = 78, PRE_DEP_DRHO = 114,
    PRE_IMU_ONLY = 126,          // != 0: plain IMUFactor (imu_factor.h) embedded in the 31-row layout: EPS / RHO rows are absent
    PRE_INFO = 128,              // 31x31 row-major covariance (sqrt_info goes to a separate [961] array per factor)
    PRE_STRIDE = 128 + 961 + 7   // 1096 doubles
};
// ILStateOrder, src/utils/parameters.h:135-150
enum { ILO_P = 0, ILO_R = 3, ILO_V = 6, ILO_EPS1 = 9, ILO_BA = 21, ILO_BG = 24, ILO_RHO1 = 27, IL_RES = 31 };

// Tangent Jacobian of one projection factor: 2x6 blocks row-major (the reference's 2x7 blocks have a zero
// 7th column, projectionTwoFrameOneCamFactor.cpp:110), plus d/d(inverse depth) and d/d(td).
struct ProjJac {
    double Ji[12], Jj[12], Je0[12], Je1[12], Jl[2], Jtd[2];
};

CERB_HD void reduce_mul(const double red[6], const m33 &M, double *, int col0, double *dst) {
    // dst[r*6 + col0 + c] = sum_k red[r*3+k] * M[k][c]
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++)
            dst[r * 6 + col0 + c] = red[r * 3] * M.m[c] + red[r * 3 + 1] * M.m[3 + c] + red[r * 3 + 2] * M.m[6 + c];
}
CERB_HD void reduce_vec(const double red[6], d3 v, double out[2]) {
    out[0] = red[0] * v.x + red[1] * v.y + red[2] * v.z;
    out[1] = red[3] * v.x + red[4] * v.y + red[5] * v.z;
}

// One projection factor.  Ri/Pi, Rj/Pj: body poses of anchor / observing frame (unused for K3);
// ric/tic, ric2/tic2: extrinsics of camera 0 / 1.  pts have z = 1, velocities z = 0.
// want_jac = false skips the Jacobian (cost-only pass).
CERB_HD void proj_eval(int kind, const m33 &Ri, d3 Pi, const m33 &Rj, d3 Pj, const m33 &ric, d3 tic, const m33 &ric2, d3 tic2,
                       double inv_dep, double td, double pix, double piy, double pjx, double pjy, double vix, double viy,
                       double vjx, double vjy, double td_i, double td_j, double sqrt_info, double r[2], ProjJac *J) {
    const d3 pts_i = mk3(pix, piy, 1.0);
    const d3 vel_i = mk3(vix, viy, 0.0);
    const d3 pts_i_td = pts_i - (td - td_i) * vel_i;
    const double pjx_td = pjx - (td - td_j) * vjx, pjy_td = pjy - (td - td_j) * vjy;
    const double inv_l = 1.0 / inv_dep;
    const d3 p_ci = inv_l * pts_i_td;
    const d3 p_bi = mv33(ric, p_ci) + tic;
    d3 p_bj, p_cj;
    const bool second_cam = (kind != PROJ_K1);
    const m33 &Rc = second_cam ? ric2 : ric;
    const d3 tc = second_cam ? tic2 : tic;
    if (kind == PROJ_K3) {
        p_bj = p_bi;
    } else {
        const d3 p_w = mv33(Ri, p_bi) + Pi;
        p_bj = mTv33(Rj, p_w - Pj);
    }
    p_cj = mTv33(Rc, p_bj - tc);
    const double inv_z = 1.0 / p_cj.z;
    r[0] = sqrt_info * (p_cj.x * inv_z - pjx_td);
    r[1] = sqrt_info * (p_cj.y * inv_z - pjy_td);
    if (!J) return;

    double red[6];
    red[0] = sqrt_info * inv_z; red[1] = 0.0; red[2] = -sqrt_info * p_cj.x * inv_z * inv_z;
    red[3] = 0.0; red[4] = sqrt_info * inv_z; red[5] = -sqrt_info * p_cj.y * inv_z * inv_z;
    const m33 RcT = tr33(Rc);
    double dummy[6];
    if (kind == PROJ_K3) {
        for (int k = 0; k < 12; k++) { J->Ji[k] = 0.0; J->Jj[k] = 0.0; }
        const m33 T = mul33(RcT, ric);                                     // ric2^T ric
        reduce_mul(red, RcT, dummy, 0, J->Je0);
        reduce_mul(red, scale33(mul33(T, skew33(p_ci)), -1.0), dummy, 3, J->Je0);
        reduce_mul(red, scale33(RcT, -1.0), dummy, 0, J->Je1);
        reduce_mul(red, skew33(p_cj), dummy, 3, J->Je1);
        reduce_vec(red, (-inv_l * inv_l) * mv33(T, pts_i), J->Jl);        // pts_i, not pts_i_td (reference quirk)
        double t2[2]; reduce_vec(red, (-inv_l) * mv33(T, vel_i), t2);
        J->Jtd[0] = t2[0] + sqrt_info * vjx; J->Jtd[1] = t2[1] + sqrt_info * vjy;
        return;
    }
    const m33 A = mulT33(Rc, tr33(Rj));          // Rc^T Rj^T
    const m33 ARi = mul33(A, Ri);
    const m33 T = mul33(ARi, ric);               // Rc^T Rj^T Ri ric
    // pose_i
    reduce_mul(red, A, dummy, 0, J->Ji);
    reduce_mul(red, scale33(mul33(ARi, skew33(p_bi)), -1.0), dummy, 3, J->Ji);
    // pose_j
    reduce_mul(red, scale33(A, -1.0), dummy, 0, J->Jj);
    reduce_mul(red, mul33(RcT, skew33(p_bj)), dummy, 3, J->Jj);
    if (kind == PROJ_K1) {
        reduce_mul(red, sub33(ARi, RcT), dummy, 0, J->Je0);                // ric^T (Rj^T Ri - I)
        const d3 Tp = mv33(T, p_ci);
        const d3 rest = mTv33(Rc, mTv33(Rj, mv33(Ri, tic) + Pi - Pj) - tic);
        m33 right = add33(scale33(mul33(T, skew33(p_ci)), -1.0), add33(skew33(Tp), skew33(rest)));
        reduce_mul(red, right, dummy, 3, J->Je0);
        for (int k = 0; k < 12; k++) J->Je1[k] = 0.0;
    } else {
        reduce_mul(red, ARi, dummy, 0, J->Je0);
        reduce_mul(red, scale33(mul33(T, skew33(p_ci)), -1.0), dummy, 3, J->Je0);
        reduce_mul(red, scale33(RcT, -1.0), dummy, 0, J->Je1);
        reduce_mul(red, skew33(p_cj), dummy, 3, J->Je1);
    }
    reduce_vec(red, (-inv_l * inv_l) * mv33(T, pts_i_td), J->Jl);
    double t2[2]; reduce_vec(red, (-inv_l) * mv33(T, vel_i), t2);
    J->Jtd[0] = t2[0] + sqrt_info * vjx; J->Jtd[1] = t2[1] + sqrt_info * vjy;
}

// ceres::HuberLoss(a) + Corrector for rho'' <= 0 (corrector.cc): returns sqrt(rho') to scale r and J by,
// *cost receives 0.5*rho(s).
CERB_HD double huber_weight(double a, double s, double *cost) {
    const double b = a * a;
    if (s > b) {
        const double rr = sqrt(s);
        *cost = 0.5 * (2.0 * a * rr - b);
        return sqrt(a / rr);
    }
    *cost = 0.5 * s;
    return 1.0;
}

// ---- IMU-leg factor -----------------------------------------------------------------------------
// The distinct pieces of the unwhitened 31-row residual / Jacobian (imu_leg_factor.cpp:221-383).
struct IMULegLin {
    double ru[31];
    m33 RiT;        // Qi^-1 as a matrix
    m33 skP;        // [Ri^T (0.5 G dt^2 + Pj - Pi - Vi dt)]x
    m33 skV;        // [Ri^T (G dt + Vj - Vi)]x
    m33 skE;        // [Ri^T (Pj - Pi)]x
    m33 M1;         // -(Qleft(Qj^-1 Qi) Qright(corrected gamma)).br        d r_q / d theta_i
    m33 M2;         // -Qleft(Qj^-1 Qi gamma).br * dq_dbg                    d r_q / d bg_i
    m33 M3;         //  Qleft(corrected gamma^-1 Qi^-1 Qj).br                d r_q / d theta_j
};

// pre: compact layout above.  pose = [p(3) q(4)], sb = [v ba bg], lb = rho[4].
CERB_HD void imu_leg_linearize(const double *pre, const double *pose_i, const double *sb_i, const double *lb_i,
                               const double *pose_j, const double *sb_j, const double *lb_j, const double *G, bool want_jac,
                               IMULegLin *out) {
    const d3 Pi = ld3(pose_i), Pj = ld3(pose_j);
    const quat Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3);
    const d3 Vi = ld3(sb_i), Bai = ld3(sb_i + 3), Bgi = ld3(sb_i + 6);
    const d3 Vj = ld3(sb_j), Baj = ld3(sb_j + 3), Bgj = ld3(sb_j + 6);
    const d3 g = ld3(G);
    const double dt = ldro(pre + PRE_SUM_DT);
    const d3 dba = Bai - ld3ro(pre + PRE_BA), dbg = Bgi - ld3ro(pre + PRE_BG);
    const m33 dq_dbg = ldm33ro(pre + PRE_DQ_DBG);
    const quat gamma = mkq(ldro(pre + PRE_DQ), ldro(pre + PRE_DQ + 1), ldro(pre + PRE_DQ + 2), ldro(pre + PRE_DQ + 3));
    const quat cgamma = qmul(gamma, qdelta(mv33(dq_dbg, dbg)));
    const d3 cdv = ld3ro(pre + PRE_DV) + mv33(ldm33ro(pre + PRE_DV_DBA), dba) + mv33(ldm33ro(pre + PRE_DV_DBG), dbg);
    const d3 cdp = ld3ro(pre + PRE_DP) + mv33(ldm33ro(pre + PRE_DP_DBA), dba) + mv33(ldm33ro(pre + PRE_DP_DBG), dbg);
    const quat Qi_inv = qinv(Qi);
    const d3 aP = qrot(Qi_inv, (0.5 * dt * dt) * g + Pj - Pi - dt * Vi);
    const d3 aV = qrot(Qi_inv, dt * g + Vj - Vi);
    const d3 aE = qrot(Qi_inv, Pj - Pi);
    const d3 rq = 2.0 * qvec(qmul(qinv(cgamma), qmul(Qi_inv, Qj)));
    double *r = out->ru;
    st3(r + ILO_P, aP - cdp);
    st3(r + ILO_R, rq);
    st3(r + ILO_V, aV - cdv);
    for (int k = 0; k < 4; k++) {
        const double drho = lb_i[k] - ldro(pre + PRE_RHO + k);
        const d3 ce = ld3ro(pre + PRE_DEPS + 3 * k) + mv33(ldm33ro(pre + PRE_DEP_DBG + 9 * k), dbg) + drho * ld3ro(pre + PRE_DEP_DRHO + 3 * k);
        st3(r + ILO_EPS1 + 3 * k, aE - ce);
        r[ILO_RHO1 + k] = lb_j[k] - lb_i[k];
    }
    st3(r + ILO_BA, Baj - Bai);
    st3(r + ILO_BG, Bgj - Bgi);
    if (ldro(pre + PRE_IMU_ONLY) != 0.0) {      // IMUFactor: 15 rows P, R, V, BA, BG only
        for (int k = ILO_EPS1; k < ILO_BA; k++) r[k] = 0.0;
        for (int k = ILO_RHO1; k < IL_RES; k++) r[k] = 0.0;
    }
    if (!want_jac) return;
    out->RiT = qtoR(Qi_inv);
    out->skP = skew33(aP);
    out->skV = skew33(aV);
    out->skE = skew33(aE);
    out->M1 = scale33(qleft_qright_br(qmul(qinv(Qj), Qi), cgamma), -1.0);
    out->M2 = scale33(mul33(qleft_br(qmul(qmul(qinv(Qj), Qi), gamma)), dq_dbg), -1.0);
    out->M3 = qleft_br(qmul(qmul(qinv(cgamma), Qi_inv), Qj));
}

// Expand the unwhitened Jacobian into a dense [31][ld] tangent matrix with columns
// [pose_i 6 | sb_i 9 | lb_i 4 | pose_j 6 | sb_j 9 | lb_j 4] (38).  Ju must be zero on entry.  The work is cut
// into 11 independent parts (part 0..8: one (a, b) entry of every 3x3 block, 9: the +-identity entries,
// 10: the leg-length columns) so that 11 threads can fill one factor concurrently.
CERB_HD void imu_leg_fill_ju_part(const IMULegLin &L, const double *pre, double *Ju, int ld, int part) {
#define JU(rr, cc) Ju[(rr) * ld + (cc)]
    // every value of `pre` (HBM / L2 in the fused solver) is loaded before the first store, so the loads are issued back to back
    const bool imu_only = pre[PRE_IMU_ONLY] != 0.0;
    if (part < 9) {
        const int a = part / 3, b = part % 3;
        const double dt = pre[PRE_SUM_DT];
        const double dp_dba = pre[PRE_DP_DBA + 3 * a + b], dp_dbg = pre[PRE_DP_DBG + 3 * a + b], dv_dba = pre[PRE_DV_DBA + 3 * a + b], dv_dbg = pre[PRE_DV_DBG + 3 * a + b];
        double dep_dbg[4];
        for (int k = 0; k < 4; k++) dep_dbg[k] = pre[PRE_DEP_DBG + 9 * k + 3 * a + b];
        const double rit = L.RiT.m[3 * a + b];
        // pose_i (cols 0..5)
        JU(ILO_P + a, b) = -rit;               JU(ILO_P + a, 3 + b) = L.skP.m[3 * a + b];
        JU(ILO_R + a, 3 + b) = L.M1.m[3 * a + b];
        JU(ILO_V + a, 3 + b) = L.skV.m[3 * a + b];
        // sb_i (cols 6..14): v, ba, bg
        JU(ILO_P + a, 6 + b) = -rit * dt;
        JU(ILO_P + a, 9 + b) = -dp_dba;
        JU(ILO_P + a, 12 + b) = -dp_dbg;
        JU(ILO_R + a, 12 + b) = L.M2.m[3 * a + b];
        JU(ILO_V + a, 6 + b) = -rit;
        JU(ILO_V + a, 9 + b) = -dv_dba;
        JU(ILO_V + a, 12 + b) = -dv_dbg;
        // pose_j (cols 19..24)
        JU(ILO_P + a, 19 + b) = rit;
        JU(ILO_R + a, 22 + b) = L.M3.m[3 * a + b];
        // sb_j (cols 25..33)
        JU(ILO_V + a, 25 + b) = rit;
        for (int k = 0; k < 4 && !imu_only; k++) {
            JU(ILO_EPS1 + 3 * k + a, b) = -rit;
            JU(ILO_EPS1 + 3 * k + a, 3 + b) = L.skE.m[3 * a + b];
            JU(ILO_EPS1 + 3 * k + a, 12 + b) = -dep_dbg[k];
            JU(ILO_EPS1 + 3 * k + a, 19 + b) = rit;
        }
    } else if (part == 9) {
        for (int a = 0; a < 3; a++) {
            JU(ILO_BA + a, 9 + a) = -1.0;  JU(ILO_BG + a, 12 + a) = -1.0;
            JU(ILO_BA + a, 28 + a) = 1.0;  JU(ILO_BG + a, 31 + a) = 1.0;
        }
    } else if (!imu_only) {
        double dr[12];
        for (int k = 0; k < 12; k++) dr[k] = pre[PRE_DEP_DRHO + k];
        for (int k = 0; k < 4; k++) {
            for (int a = 0; a < 3; a++) JU(ILO_EPS1 + 3 * k + a, 15 + k) = -dr[3 * k + a];
            JU(ILO_RHO1 + k, 15 + k) = -1.0;
            JU(ILO_RHO1 + k, 34 + k) = 1.0;
        }
    }
#undef JU
}
CERB_HD void imu_leg_fill_ju(const IMULegLin &L, const double *pre, double *Ju, int ld) {
    for (int part = 0; part < 11; part++) imu_leg_fill_ju_part(L, pre, Ju, ld, part);
}

// ---- A1 leg kinematics (closed forms, A1Kinematics.cpp:43-220) -------------------------------------
// q = joint angles (3), lc = rho_opt (lower-leg length), fix = [ox, oy, d, lt].  Matrices column-major
// like the Eigen objects the reference fills through .data().
struct A1Trig { double c0, s0, c1, s1, c12, s12; };
CERB_HD A1Trig a1_trig(const double *q) {
    A1Trig t; t.c0 = cos(q[0]); t.s0 = sin(q[0]); t.c1 = cos(q[1]); t.s1 = sin(q[1]);
    t.c12 = cos(q[1] + q[2]); t.s12 = sin(q[1] + q[2]); return t;
}
CERB_HD void a1_fk(const A1Trig &t, double lc, const double *fix, double *p) {
    p[0] = fix[0] - fix[3] * t.s1 - lc * t.s12;
    p[1] = fix[1] + fix[2] * t.c0 + fix[3] * t.c1 * t.s0 + lc * t.s0 * t.c12;
    p[2] = fix[2] * t.s0 - fix[3] * t.c0 * t.c1 - lc * t.c0 * t.c12;
}
CERB_HD void a1_jac(const A1Trig &t, double lc, const double *fix, double *J) {   // J[c*3+r]
    const double A = fix[3] * t.s1 + lc * t.s12, B = fix[3] * t.c1 + lc * t.c12;
    J[0] = 0.0;          J[1] = -fix[2] * t.s0 + t.c0 * B; J[2] = fix[2] * t.c0 + t.s0 * B;
    J[3] = -B;           J[4] = -t.s0 * A;                 J[5] = t.c0 * A;
    J[6] = -lc * t.c12;  J[7] = -t.s0 * lc * t.s12;        J[8] = t.c0 * lc * t.s12;
}
CERB_HD void a1_dfk_drho(const A1Trig &t, double *o) { o[0] = -t.s12; o[1] = t.c12 * t.s0; o[2] = -t.c12 * t.c0; }
CERB_HD void a1_dJ_drho(const A1Trig &t, double *o) {
    o[0] = 0.0;     o[1] = t.c0 * t.c12;  o[2] = t.s0 * t.c12;
    o[3] = -t.c12;  o[4] = -t.s0 * t.s12; o[5] = t.c0 * t.s12;
    o[6] = -t.c12;  o[7] = -t.s0 * t.s12; o[8] = t.c0 * t.s12;
}
CERB_HD void a1_dJ_dq(const A1Trig &t, double lc, const double *fix, double *o) {   // o[m*9 + c*3 + r]
    const double d = fix[2], lt = fix[3];
    const double A = lt * t.s1 + lc * t.s12, B = lt * t.c1 + lc * t.c12, Cc = lc * t.c12, Ss = lc * t.s12;
    o[0] = 0.0;  o[1] = -d * t.c0 - t.s0 * B; o[2] = -d * t.s0 + t.c0 * B;
    o[3] = 0.0;  o[4] = -t.c0 * A;            o[5] = -t.s0 * A;
    o[6] = 0.0;  o[7] = -t.c0 * Ss;           o[8] = -t.s0 * Ss;
    o[9] = 0.0;  o[10] = -t.c0 * A;           o[11] = -t.s0 * A;
    o[12] = A;   o[13] = -t.s0 * B;           o[14] = t.c0 * B;
    o[15] = Ss;  o[16] = -t.s0 * Cc;          o[17] = t.c0 * Cc;
    o[18] = 0.0; o[19] = -t.c0 * Ss;          o[20] = -t.s0 * Ss;
    o[21] = Ss;  o[22] = -t.s0 * Cc;          o[23] = t.c0 * Cc;
    o[24] = Ss;  o[25] = -t.s0 * Cc;          o[26] = t.c0 * Cc;
}

}  // namespace cerb
