// eval_kernels.cuh -- "one kernel per factor family": batched residual + Jacobian evaluation with the
// output conventions of ceres::CostFunction::Evaluate (row-major rows x global_size, zero 7th pose
// column), plus the per-solve preparation kernels (IMU-leg sqrt_info, prior J0^T J0).
//   projection_eval_kernel   K1/K2/K3   thread per factor, SoA-friendly strides
//   imu_leg_prepare_kernel   K5 prep    warp per factor: covariance -> sqrt_info (imu_leg_factor.cpp:197-198)
//   imu_leg_eval_kernel      K5         CTA per factor
//   prior_prepare_kernel     K6 prep    CTA per window: Hp = J0^T J0 (constant per solve: the prior is linear)
//   prior_eval_kernel        K6         single prior, MarginalizationFactor::Evaluate
//   a1_kinematics_kernel     a10        thread per leg
#pragma once
#include "factors.cuh"

namespace cerb {

// ------------------------------------------------------------------------------------------- K1..K3
CERB_GLOBAL void projection_eval_kernel(int kind, int n, const double *pose_i, const double *pose_j, const double *ex0, const double *ex1,
                                        const double *inv_dep, const double *td, const double *pts_i, const double *pts_j,
                                        const double *vel_i, const double *vel_j, const double *td_i, const double *td_j,
                                        double sqrt_info, double *residuals, double *jacobians) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        m33 Ri = ident33(), Rj = ident33(), ric2 = ident33();
        d3 Pi = mk3(0, 0, 0), Pj = mk3(0, 0, 0), tic2 = mk3(0, 0, 0);
        if (kind != PROJ_K3) {
            Ri = qtoR(ldq(pose_i + 7 * k + 3)); Pi = ld3(pose_i + 7 * k);
            Rj = qtoR(ldq(pose_j + 7 * k + 3)); Pj = ld3(pose_j + 7 * k);
        }
        const m33 ric = qtoR(ldq(ex0 + 7 * k + 3)); const d3 tic = ld3(ex0 + 7 * k);
        if (kind != PROJ_K1) { ric2 = qtoR(ldq(ex1 + 7 * k + 3)); tic2 = ld3(ex1 + 7 * k); }
        double r[2]; ProjJac J;
        proj_eval(kind, Ri, Pi, Rj, Pj, ric, tic, ric2, tic2, inv_dep[k], td[k], pts_i[3 * k], pts_i[3 * k + 1], pts_j[3 * k], pts_j[3 * k + 1],
                  vel_i[2 * k], vel_i[2 * k + 1], vel_j[2 * k], vel_j[2 * k + 1], td_i[k], td_j[k], sqrt_info, r, jacobians ? &J : nullptr);
        if (residuals) { residuals[2 * k] = r[0]; residuals[2 * k + 1] = r[1]; }
        if (jacobians) {
            const int JS = kind == PROJ_K1 ? 46 : (kind == PROJ_K2 ? 60 : 32);
            double *o = jacobians + (size_t)k * JS;
            int off = 0;
            const double *blocks[4]; int nb = 0;
            if (kind != PROJ_K3) { blocks[nb++] = J.Ji; blocks[nb++] = J.Jj; }
            blocks[nb++] = J.Je0;
            if (kind != PROJ_K1) blocks[nb++] = J.Je1;
            for (int b = 0; b < nb; b++) {
                for (int rr = 0; rr < 2; rr++) { for (int c = 0; c < 6; c++) o[off + rr * 7 + c] = blocks[b][rr * 6 + c]; o[off + rr * 7 + 6] = 0.0; }
                off += 14;
            }
            o[off] = J.Jl[0]; o[off + 1] = J.Jl[1]; o[off + 2] = J.Jtd[0]; o[off + 3] = J.Jtd[1];
        }
    }
}

// ------------------------------------------------------------------------------------------- K5 prep
// One warp: sm points at >= 3*31*33 doubles of shared memory private to this warp.
// cov (global, 31x31 row-major) in; info out = sqrt_info = LLT(cov^-1).matrixL().transpose()
// (upper triangular, zeros below).  Same sequence as the reference: inverse by LU with partial pivoting,
// then Cholesky of the (lower triangle of the) inverse.  Returns false (in *ok) on a non-positive pivot.
CERB_D void warp_sqrt_info(const double *cov, double *info, double *sm, int lane, int *ok_flag) {
    const int N = 31, LD = 33;
    double *A = sm, *X = sm + N * LD, *piv = sm + 2 * N * LD;   // piv: permutation as doubles + scratch
    for (int i = lane; i < N * N; i += 32) A[(i / N) * LD + (i % N)] = cov[i];
    if (lane < N) piv[lane] = (double)lane;
    __syncwarp();
    // LU with partial pivoting (right-looking), lanes = rows for the update
    for (int k = 0; k < N; k++) {
        // partial pivoting: first row of maximal |A[i][k]|, i >= k (same choice as a serial scan), by a warp arg-max
        double bv = (lane >= k && lane < N) ? fabs(A[lane * LD + k]) : -1.0, bi = (double)lane;
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_sync(0xffffffffu, bv, (lane + o) & 31), oi = __shfl_sync(0xffffffffu, bi, (lane + o) & 31);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0 && bv == 0.0) *ok_flag = 0;
        const int p = (int)bi;
        if (p != k) {
            if (lane < N) { double t = A[k * LD + lane]; A[k * LD + lane] = A[p * LD + lane]; A[p * LD + lane] = t; }
            if (lane == 0) { double t = piv[k]; piv[k] = piv[p]; piv[p] = t; }
        }
        __syncwarp();
        if (lane > k && lane < N) {
            const double f = A[lane * LD + k] / A[k * LD + k];
            A[lane * LD + k] = f;
            for (int j = k + 1; j < N; j++) A[lane * LD + j] -= f * A[k * LD + j];
        }
        __syncwarp();
    }
    // inverse: lane c solves L U x = P e_c
    if (lane < N) {
        const int c = lane;
        for (int i = 0; i < N; i++) {
            double s = ((int)piv[i] == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= A[i * LD + k] * X[k * LD + c];
            X[i * LD + c] = s;
        }
        for (int i = N - 1; i >= 0; i--) {
            double s = X[i * LD + c];
            for (int k = i + 1; k < N; k++) s -= A[i * LD + k] * X[k * LD + c];
            X[i * LD + c] = s / A[i * LD + i];
        }
    }
    __syncwarp();
    // Cholesky of the lower triangle of X (left-looking), L overwrites the lower triangle
    for (int j = 0; j < N; j++) {
        {   // diagonal: X[j][j] - sum_k L[j][k]^2 with the products spread over the lanes (fixed-order shuffle reduction)
            double part = (lane < j) ? X[j * LD + lane] * X[j * LD + lane] : 0.0;
            for (int o = 16; o > 0; o >>= 1) part += __shfl_sync(0xffffffffu, part, (lane + o) & 31);
            if (lane == 0) {
                double s = X[j * LD + j] - part;
                if (!(s > 0.0)) { *ok_flag = 0; s = 1.0; }
                X[j * LD + j] = sqrt(s);
            }
        }
        __syncwarp();
        if (lane > j && lane < N) {
            double t = X[lane * LD + j];
            for (int k = 0; k < j; k++) t -= X[lane * LD + k] * X[j * LD + k];
            X[lane * LD + j] = t / X[j * LD + j];
        }
        __syncwarp();
    }
    for (int i = lane; i < N * N; i += 32) { const int r = i / N, c = i % N; info[i] = (c >= r) ? X[c * LD + r] : 0.0; }
    __syncwarp();
}

// grid = ceil(n_factors / 2), block = 64 (2 warps, one factor each); sinfo [n_factors][961]
CERB_GLOBAL void imu_leg_prepare_kernel(int n_factors, const double *pre, double *sinfo) {
    __shared__ double sm[2][2 * 31 * 33 + 64];
    __shared__ int okf[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * 2 + warp;
    if (f < n_factors) {   // whole warps take the branch together
        if (lane == 0) okf[warp] = 1;
        __syncwarp();
        warp_sqrt_info(pre + (size_t)f * PRE_STRIDE + PRE_INFO, sinfo + (size_t)f * 961, sm[warp], lane, &okf[warp]);
        // a failed factorisation (covariance not positive definite) poisons the factor: the solve reports non-finite
        if (!okf[warp]) for (int i = lane; i < 961; i += 32) sinfo[(size_t)f * 961 + i] = nan("");
    }
}

// ------------------------------------------------------------------------------------------- K5
// Host-facing evaluate of n factors.  pre: compact device layout with sqrt_info already prepared.
// params [n][40], residuals [n][31], jacobians [n][31*40] (blocks 31x7,31x9,31x4,31x7,31x9,31x4 row-major).
// grid = n, block = 128.
CERB_GLOBAL void imu_leg_eval_kernel(int n, const double *pre_all, const double *sinfo, const double *params, const double *G, double *residuals, double *jacobians) {
    __shared__ double Ju[31 * 39];
    __shared__ IMULegLin lin;
    const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double *pre = pre_all + (size_t)k * PRE_STRIDE;
    const double *p = params + (size_t)k * 40;
    for (int i = tid; i < 31 * 39; i += nt) Ju[i] = 0.0;
    if (tid == 0) imu_leg_linearize(pre, p, p + 7, p + 16, p + 20, p + 27, p + 36, G, true, &lin);
    __syncthreads();
    if (tid == 0) { imu_leg_fill_ju(lin, pre, Ju, 39); for (int r = 0; r < 31; r++) Ju[r * 39 + 38] = lin.ru[r]; }
    __syncthreads();
    const double *S = sinfo + (size_t)k * 961;
    const int goff[6] = {0, 7, 16, 20, 27, 36}, gsz[6] = {7, 9, 4, 7, 9, 4}, toff[6] = {0, 6, 15, 19, 25, 34}, tsz[6] = {6, 9, 4, 6, 9, 4};
    for (int idx = tid; idx < 31 * 39; idx += nt) {
        const int r = idx / 39, c = idx % 39;
        double s = 0.0;
        for (int q = r; q < 31; q++) s += S[r * 31 + q] * Ju[q * 39 + c];
        if (c == 38) { if (residuals) residuals[(size_t)k * 31 + r] = s; continue; }
        if (!jacobians) continue;
        int b = 0; while (b < 5 && c >= toff[b + 1]) b++;
        double *o = jacobians + (size_t)k * 31 * 40 + 31 * goff[b];
        o[r * gsz[b] + (c - toff[b])] = s;
    }
    if (jacobians) {   // zero 7th column of the pose blocks
        for (int r = tid; r < 31; r += nt) { jacobians[(size_t)k * 31 * 40 + r * 7 + 6] = 0.0; jacobians[(size_t)k * 31 * 40 + 31 * 20 + r * 7 + 6] = 0.0; }
    }
    (void)tsz;
}

// ------------------------------------------------------------------------------------------- K6
// Prior device layout per window: J0 [96*96] column-major n x n (leading dim n), r0 [96],
// meta[0] = valid, meta[1] = n, meta[2] = num_blocks, meta[4+3b..] = (kind, index, col), x0 [16][7].
enum { PRIOR_META_STRIDE = 64, PRIOR_LD = 96 };

// Hp (row-major [n][n], leading dim PRIOR_LD) = J0^T J0.
// J0 (n x n, column-major) is staged once into shared memory as tile[k][a] (row stride PRIOR_TLD, zero padded to multiples of 8),
// then the upper 8 x 8 blocks of the Gram matrix are contracted on the fp64 tensor cores and mirrored on the way out.
// grid = n_windows, block = 256, dynamic shared memory = PRIOR_TROWS * PRIOR_TLD doubles.
enum { PRIOR_TLD = 108, PRIOR_TROWS = 96 };                  // 108 = 12 (mod 16): conflict-free fragment loads
CERB_GLOBAL void prior_prepare_kernel(const double *J_all, const int *meta_all, double *Hp_all) {
    CERB_DYN_SMEM(double, tile);
    const int w = blockIdx.x, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int *meta = meta_all + (size_t)w * PRIOR_META_STRIDE;
    const int n = meta[0] ? meta[1] : 0;
    if (n == 0) return;
    const double *J = J_all + (size_t)w * PRIOR_LD * PRIOR_LD;
    double *Hp = Hp_all + (size_t)w * PRIOR_LD * PRIOR_LD;
    const int nb = (n + 7) >> 3, np = 8 * nb;                 // blocks per side, padded size
    for (int e = tid; e < np * PRIOR_TLD; e += blockDim.x) tile[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < n * n; e += blockDim.x) { const int a = e / n, k = e % n; tile[k * PRIOR_TLD + a] = J[e]; }      // J[a * n + k]: column a, row k
    __syncthreads();
    const int nblk = nb * (nb + 1) / 2;
    for (int b = wid; b < nblk; b += (int)(blockDim.x >> 5)) {
        int mi = 0, idx = b;
        while (idx >= nb - mi) { idx -= nb - mi; mi++; }
        const int ni = mi + idx;
        double a0 = 0.0, a1 = 0.0;
        for (int ks = 0; ks < 2 * nb; ks++) {
            const double av = tile[(4 * ks + (lane & 3)) * PRIOR_TLD + 8 * mi + (lane >> 2)];
            const double bv = tile[(4 * ks + (lane & 3)) * PRIOR_TLD + 8 * ni + (lane >> 2)];
            CERB_DMMA(a0, a1, av, bv, a0, a1);
        }
        const int ra = 8 * mi + (lane >> 2);
        for (int e = 0; e < 2; e++) {
            const int rb = 8 * ni + 2 * (lane & 3) + e;
            const double v = e ? a1 : a0;
            if (ra < n && rb < n && ra <= rb) { Hp[ra * PRIOR_LD + rb] = v; Hp[rb * PRIOR_LD + ra] = v; }
        }
    }
}

CERB_HD int prior_block_size(int kind) { return (kind == 0 || kind == 3) ? 7 : (kind == 1 ? 9 : (kind == 2 ? 4 : 1)); }
// state layout: pose[11][7] @0, speedbias[11][9] @77, legbias[11][4] @176, ex[2][7] @220, td @234
enum { ST_POSE = 0, ST_SB = 77, ST_LB = 176, ST_EX = 220, ST_TD = 234, ST_SIZE = 235, ST_STRIDE = 240 };
CERB_HD int prior_block_state_offset(int kind, int index) {
    return kind == 0 ? ST_POSE + 7 * index : (kind == 1 ? ST_SB + 9 * index : (kind == 2 ? ST_LB + 4 * index : (kind == 3 ? ST_EX + 7 * index : ST_TD)));
}
// dx of one kept block (marginalization_factor.cpp:357-377); writes local_size entries
CERB_HD void prior_block_dx(int kind, const double *x, const double *x0, double *dx) {
    const int size = prior_block_size(kind);
    if (size != 7) { for (int k = 0; k < size; k++) dx[k] = x[k] - x0[k]; return; }
    for (int k = 0; k < 3; k++) dx[k] = x[k] - x0[k];
    const quat dq = qmul(qinv(ldq(x0 + 3)), ldq(x + 3));
    double sgn = (dq.w >= 0) ? 2.0 : -2.0;
    dx[3] = sgn * dq.x; dx[4] = sgn * dq.y; dx[5] = sgn * dq.z;
}

// MarginalizationFactor::Evaluate of one prior at one state.  block = 128, grid = 1.
CERB_GLOBAL void prior_eval_kernel(const double *J, const double *r0, const int *meta, const double *x0, const double *state,
                                   double *residuals, double *jacobians) {
    __shared__ double dx[PRIOR_LD];
    const int n = meta[1], nb = meta[2], tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < PRIOR_LD; i += nt) dx[i] = 0.0;
    __syncthreads();
    if (tid < nb) {
        const int kind = meta[4 + 3 * tid], index = meta[5 + 3 * tid], col = meta[6 + 3 * tid];
        prior_block_dx(kind, state + prior_block_state_offset(kind, index), x0 + 9 * tid, dx + col);
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) { double s = r0[i]; for (int k = 0; k < n; k++) s += J[(size_t)k * n + i] * dx[k]; residuals[i] = s; }
    if (jacobians) {
        size_t off = 0;
        for (int b = 0; b < nb; b++) {
            const int kind = meta[4 + 3 * b], col = meta[6 + 3 * b];
            const int size = prior_block_size(kind), local = size == 7 ? 6 : size;
            for (int idx = tid; idx < n * size; idx += nt) {
                const int r = idx / size, c = idx % size;
                jacobians[off + idx] = (c < local) ? J[(size_t)(col + c) * n + r] : 0.0;
            }
            off += (size_t)n * size;
        }
    }
}

// ------------------------------------------------------------------------------------------- a10
CERB_GLOBAL void a1_kinematics_kernel(int n, const double *q, const double *rho_opt, const double *rho_fix, double *fk, double *jac,
                                      double *dfk, double *djq, double *djr) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const A1Trig t = a1_trig(q + 3 * k);
        if (fk) a1_fk(t, rho_opt[k], rho_fix + 4 * k, fk + 3 * k);
        if (jac) a1_jac(t, rho_opt[k], rho_fix + 4 * k, jac + 9 * k);
        if (dfk) a1_dfk_drho(t, dfk + 3 * k);
        if (djq) a1_dJ_dq(t, rho_opt[k], rho_fix + 4 * k, djq + 27 * k);
        if (djr) a1_dJ_drho(t, djr + 9 * k);
    }
}

}  // namespace cerb
