// marg_kernels.cuh -- the dense tail of MarginalizationInfo::marginalize() on the device, one CTA per window:
//   reference src/factor/marginalization_factor.cpp:281-305
//     Amm = 0.5 (A_mm + A_mm^T); SelfAdjointEigenSolver(Amm); Amm_inv = V diag(lambda > eps ? 1 / lambda : 0) V^T
//     A   = Arr - Arm Amm_inv Amr;  b = brr - Arm Amm_inv bmm
//     SelfAdjointEigenSolver(A) (lower triangle);  S = lambda > eps ? lambda : 0;  S_inv = lambda > eps ? 1 / lambda : 0
//     linearized_jacobians = diag(sqrt S) V^T;  linearized_residuals = diag(sqrt S_inv) V^T b
// The eigen-solver is a parallel two-sided Jacobi iteration (round-robin pair ordering: kp / 2 disjoint rotations per round,
// column phase, row phase, re-symmetrisation), in fp64, on matrices that live in SHARED memory (m = 19 + tracks anchored at frame 0 -> 169 x 169
// for the 150-feature configuration = 223 KB, 86 x 86 for the kept block; larger m falls back to the L2-resident workspace).  Rotation formulas as in the CPU restatement
// (oracle/ref_math.h sym_eig_jacobi), different pair order; the rows of linearized_jacobians come out in the order the
// eigenvalues sit on the diagonal (J^T J and J^T r, the only things MarginalizationFactor uses, do not depend on it).
#pragma once
#include "compat.h"

namespace cerb {

#if defined(CERB_CUSIM)
constexpr int MARG_THREADS = 64;      // the CPU simulator pays ~ one futex wake-up per thread and barrier; two warps still exercise every path
#else
constexpr int MARG_THREADS = 512;      // 64 registers per thread, one CTA per SM (shared memory): more loads in flight for the L2-resident T pass
#endif
constexpr int MARG_MAX_SWEEPS = 60;

// Memory plan of one CTA (ld = k | 1: odd leading dimensions keep the row pass free of shared-memory bank conflicts):
//   phase 1  M1 (m x m) in SHARED memory when it fits (m <= 169: the 150-feature configuration), else in the global workspace;
//            T = V1^T [Amr | bm] (m x (n + 1)) in the global workspace (L2): the rotations are applied to its rows as they are applied to the
//            rows of M1, so V1 itself is never formed:  Arm Amm_inv [Amr | bm] = T^T diag(lambda > eps ? 1 / lambda : 0) T
//   phase 2  T staged in shared memory for the contraction, M2 and V2 (n x n, n <= 96) in shared memory
constexpr size_t MARG_SMEM_MAX = 232448;               // 227 KB: the opt-in limit of dynamic shared memory per block on sm_100 (the kernel has no static shared memory)
CERB_HD int marg_ld(int k) { return k | 1; }
CERB_HD size_t marg_fixed_doubles(int m, int n) { const int k = m > n ? m : n; return 2 * (size_t)(k + 2) + 4 + (size_t)(k + 2) / 4 + 1; }   // (c, s) pairs | 1 / lambda | flags | pair table (ints)
CERB_HD bool marg_m1_in_smem(int m, int n) { return (marg_fixed_doubles(m, n) + (size_t)marg_ld(m) * marg_ld(m)) * sizeof(double) <= MARG_SMEM_MAX; }
// phase 2 with T staged in shared memory: [M2 | X], X = T during the contraction, V2 afterwards (whatever the m of the window)
CERB_HD size_t marg_t_body(int m, int n) { const size_t t = (size_t)m * (n + 1), q = (size_t)marg_ld(n) * marg_ld(n); return q + (t > q ? t : q); }
CERB_HD bool marg_t_in_smem(int m, int n) { return (marg_fixed_doubles(m, n) + marg_t_body(m, n)) * sizeof(double) <= MARG_SMEM_MAX; }
CERB_HD size_t marg_smem_bytes(int m, int n) {
    size_t body = 2 * (size_t)marg_ld(n) * marg_ld(n);
    if (marg_m1_in_smem(m, n)) body = body > (size_t)marg_ld(m) * marg_ld(m) ? body : (size_t)marg_ld(m) * marg_ld(m);
    if (marg_t_in_smem(m, n)) { const size_t t = marg_t_body(m, n); body = body > t ? body : t; }
    return (marg_fixed_doubles(m, n) + body) * sizeof(double);
}
// per-CTA global workspace in doubles: T | M1 (only when it does not fit in shared memory) | br
CERB_HD size_t marg_ws_doubles(int m, int n) {
    return (size_t)m * (n + 1) + 1 + (marg_m1_in_smem(m, n) ? 0 : (size_t)marg_ld(m) * marg_ld(m)) + (size_t)n + 8;
}

// pair t (0 .. kp / 2 - 1) of round r (0 .. kp - 2) of a round-robin tournament over kp (even) players, p < q
CERB_D void jacobi_pair(int t, int r, int kp, int &p, int &q) {
    const int md = kp - 1;
    int a, b;
    if (t == 0) { a = md; b = r; }
    else { a = (r + t) % md; b = (r - t + md) % md; }
    p = a < b ? a : b; q = a < b ? b : a;
}

// rotation angles of round r from the upper triangle of M as it stands; (c, s) -> cs, the pair -> pq (p | q << 16; the passes of the
// round read it instead of redoing the modulo arithmetic), a non-trivial rotation raises *flag
CERB_D void jacobi_angles(const double *M, int ld, int k, int kp, int r, double *cs, int *pq, int *flag) {
    for (int t = threadIdx.x; t < kp / 2; t += blockDim.x) {
        int p, q; jacobi_pair(t, r, kp, p, q);
        double c = 1.0, s = 0.0;
        if (q < k) {
            const double apq = M[p + q * ld], app = M[p + p * ld], aqq = M[q + q * ld];
            if (apq != 0.0 && fabs(apq) > 1e-15 * sqrt(fabs(app * aqq))) {
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                c = 1.0 / sqrt(tt * tt + 1.0); s = tt * c;
                if (s != 0.0) *flag = 1;
            }
        }
        cs[2 * t] = c; cs[2 * t + 1] = s; pq[t] = p | (q << 16);
    }
}

// Eigen-decomposition of the symmetric k x k matrix M (column-major, leading dimension ld): on return the eigenvalues are on the
// diagonal of M.  V (optional, k x k, leading dimension ldv, set to the identity here): the eigenvectors as columns.  T (optional,
// k x nct ROW-major, leading dimension ldt, nct <= 128): replaced by V^T T (its rows are rotated like the rows of M).  Called by all
// threads of the CTA; returns the number of sweeps.  Three CTA barriers per round: column pass | row pass | re-symmetrisation together
// with the angles of the next round (both only read the upper triangle the row pass left).  flag[sweep & 1] collects "some rotation was
// non-trivial".  T lives in global memory (L2): a warp loads the two rows of its pair (coalesced) BEFORE it does its share of the row
// pass on the shared-memory matrix and rotates / stores them afterwards, so the L2 round trip hides behind the row pass.
CERB_D int jacobi_eig(double *M, int ld, int k, double *V, int ldv, double *T, int ldt, int nct, double *cs, int *pq, int *flag) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
    const int kp = k + (k & 1), half = kp / 2, rounds = kp - 1;
    if (V) for (int j = wid; j < k; j += nwarp) for (int i = lane; i < k; i += 32) V[i + j * ldv] = (i == j) ? 1.0 : 0.0;
    if (tid == 0) { flag[0] = 0; flag[1] = 0; }
    __syncthreads();
    jacobi_angles(M, ld, k, kp, 0, cs, pq, flag);
    __syncthreads();
    int sweeps = 0;
    for (; sweeps < MARG_MAX_SWEEPS; sweeps++) {
        for (int r = 0; r < rounds; r++) {
            for (int t = wid; t < half; t += nwarp) {         // columns p, q of M (and V)
                const double c = cs[2 * t], s = cs[2 * t + 1];
                if (s == 0.0) continue;
                const int p = pq[t] & 0xffff, q = pq[t] >> 16;
                double *mp = M + p * ld, *mq = M + q * ld;
                for (int i = lane; i < k; i += 32) { const double a = mp[i], b = mq[i]; mp[i] = c * a - s * b; mq[i] = s * a + c * b; }
                if (V) {
                    double *vp = V + p * ldv, *vq = V + q * ldv;
                    for (int i = lane; i < k; i += 32) { const double e = vp[i], f = vq[i]; vp[i] = c * e - s * f; vq[i] = s * e + c * f; }
                }
            }
            __syncthreads();
            for (int t0 = wid; t0 < half; t0 += 2 * nwarp) {      // rows p, q of M (and T): two pairs per warp iteration
                double ta[2][4], tb[2][4];
                _Pragma("unroll")
                for (int u = 0; u < 2; u++) {                     // T rows first: the loads stay in flight during the row pass below
                    const int t = t0 + u * nwarp;
                    if (T && t < half && cs[2 * t + 1] != 0.0) {
                        const double *tp = T + (pq[t] & 0xffff) * ldt, *tq = T + (pq[t] >> 16) * ldt;
                        _Pragma("unroll")
                        for (int v = 0; v < 4; v++) { const int j = lane + 32 * v; ta[u][v] = j < nct ? tp[j] : 0.0; tb[u][v] = j < nct ? tq[j] : 0.0; }
                    }
                }
                _Pragma("unroll")
                for (int u = 0; u < 2; u++) {
                    const int t = t0 + u * nwarp;
                    if (t >= half) continue;
                    const double c = cs[2 * t], s = cs[2 * t + 1];
                    if (s == 0.0) continue;
                    const int p = pq[t] & 0xffff, q = pq[t] >> 16;
                    for (int j = lane; j < k; j += 32) {
                        const double a = M[p + j * ld], b = M[q + j * ld];
                        M[p + j * ld] = c * a - s * b; M[q + j * ld] = s * a + c * b;
                    }
                    if (T) {
                        double *tp = T + p * ldt, *tq = T + q * ldt;
                        _Pragma("unroll")
                        for (int v = 0; v < 4; v++) { const int j = lane + 32 * v; if (j < nct) { tp[j] = c * ta[u][v] - s * tb[u][v]; tq[j] = s * ta[u][v] + c * tb[u][v]; } }
                    }
                }
            }
            // every thread has read last sweep's verdict by now (it did so before this sweep's first column pass)
            if (r == 0 && tid == 0) flag[(sweeps + 1) & 1] = 0;
            __syncthreads();
            // keep M exactly symmetric: the column and the row pass round differently, and an asymmetric residue of eps |M| is enough to
            // keep the null space of a rank-deficient Schur complement rotating for ever (the angles are taken from the upper triangle)
            for (int j = wid; j < k; j += nwarp) for (int i = j + 1 + lane; i < k; i += 32) M[i + j * ld] = M[j + i * ld];
            if (r + 1 < rounds) jacobi_angles(M, ld, k, kp, r + 1, cs, pq, flag + (sweeps & 1));
            else jacobi_angles(M, ld, k, kp, 0, cs, pq, flag + ((sweeps + 1) & 1));
            __syncthreads();
        }
        if (!flag[sweeps & 1]) break;
    }
    return sweeps;
}

// A [n_windows][(m + n)^2] row-major (dropped block first), b [n_windows][m + n]; lin_J [n_windows][n * n] column-major,
// lin_r [n_windows][n]; ws [gridDim.x][marg_ws_doubles(m, n)]; sweeps [n_windows][2] (diagnostics, may be null)
// dims (optional): per-window sizes [n_windows][4] = m, n, status (only status == 1 windows are processed), -; then mmax / nmax bound them
// and A / b / lin_J / lin_r are strided by A_stride / b_stride / J_stride / r_stride doubles per window (0: tight, from m and n)
CERB_GLOBAL void marg_schur_kernel(int n_windows, int mmax, int nmax, const int *dims, const double *A_all, long A_stride, const double *b_all, long b_stride, double eps,
                                   double *ws_all, double *lin_J, long J_stride, double *lin_r, long r_stride, int *sweeps) {
    CERB_DYN_SMEM(double, sm);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int kmax = mmax > nmax ? mmax : nmax;
    double *cs = sm, *inv = sm + (kmax + 2); int *flag = reinterpret_cast<int *>(sm + 2 * (kmax + 2)), *pq = flag + 8;
    double *body = sm + marg_fixed_doubles(mmax, nmax);
    const bool m1_smem = marg_m1_in_smem(mmax, nmax), t_smem = marg_t_in_smem(mmax, nmax);
    double *wsp = ws_all + (size_t)blockIdx.x * marg_ws_doubles(mmax, nmax);
    double *Tg = wsp, *M1g = Tg + (size_t)mmax * (nmax + 1) + 1, *br = M1g + (m1_smem ? 0 : (size_t)marg_ld(mmax) * marg_ld(mmax));
    for (int w = blockIdx.x; w < n_windows; w += gridDim.x) {
        const int m = dims ? dims[4 * w] : mmax, n = dims ? dims[4 * w + 1] : nmax;
        if (dims && dims[4 * w + 2] != 1) continue;
        const int pos = m + n, nc = n + 1, ld1 = marg_ld(m), ld2 = marg_ld(n), ldt = nc;       // T: m rows of nc doubles
        const double *A = A_all + (size_t)w * (A_stride ? A_stride : (long)pos * pos), *b = b_all + (size_t)w * (b_stride ? b_stride : (long)pos);
        // ---- phase 1: Amm = 0.5 (Amm + Amm^T) = V1 diag(lambda) V1^T;  T = V1^T [Amr | bm] --------------------------------------
        double *M1 = m1_smem ? body : M1g;
        for (int j = tid / 32; j < m; j += nt / 32) for (int i = tid & 31; i < m; i += 32) M1[i + j * ld1] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]);
        for (int e = tid; e < m * nc; e += nt) { const int c = e % nc, i = e / nc; Tg[(size_t)i * ldt + c] = c < n ? A[(size_t)i * pos + m + c] : b[i]; }
        __syncthreads();
        // two instantiations: with the matrix in shared memory the compiler sees the address space (LDS / STS with 32-bit addresses instead of
        // generic 64-bit loads: the passes are instruction-bound)
        const int sw1 = m1_smem ? jacobi_eig(body, ld1, m, nullptr, 0, Tg, ldt, nc, cs, pq, flag) : jacobi_eig(M1g, ld1, m, nullptr, 0, Tg, ldt, nc, cs, pq, flag);
        for (int i = tid; i < m; i += nt) { const double lam = M1[i + (size_t)i * ld1]; inv[i] = lam > eps ? 1.0 / lam : 0.0; }
        __syncthreads();
        // ---- [Ar | br] = [Arr | brr] - T^T diag(inv) T (lower triangle; SelfAdjointEigenSolver reads the lower triangle) ---------------
        const size_t tsz = (size_t)m * nc;
        double *M2 = body, *Ts = t_smem ? body + (size_t)ld2 * ld2 : Tg;
        if (t_smem) { for (int e = tid; e < (int)tsz; e += nt) Ts[e] = Tg[e]; __syncthreads(); }
        for (int e = tid; e < n * nc; e += nt) {
            const int r = e % n, c = e / n;
            if (c < n && c > r) continue;
            double acc = 0.0;
            for (int i = 0; i < m; i++) acc += Ts[(size_t)i * ldt + r] * inv[i] * Ts[(size_t)i * ldt + c];
            if (c < n) { const double v = A[(size_t)(m + r) * pos + m + c] - acc; M2[r + (size_t)c * ld2] = v; M2[c + (size_t)r * ld2] = v; }
            else br[r] = b[m + r] - acc;
        }
        __syncthreads();
        // ---- phase 2: A = V2 diag(lambda) V2^T;  linearized_jacobians = sqrt(S) V2^T, linearized_residuals = sqrt(S_inv) V2^T b ----------
        double *V2 = M2 + (size_t)ld2 * ld2;                                               // over the dead staged T
        const int sw2 = jacobi_eig(M2, ld2, n, V2, ld2, nullptr, 0, 0, cs, pq, flag);
        double *Jo = lin_J + (size_t)w * (J_stride ? J_stride : (long)n * n), *ro = lin_r + (size_t)w * (r_stride ? r_stride : (long)n);
        for (int e = tid; e < n * n; e += nt) {
            const int kk = e % n, j = e / n;
            const double lam = M2[kk + (size_t)kk * ld2];
            Jo[e] = lam > eps ? sqrt(lam) * V2[j + (size_t)kk * ld2] : 0.0;
        }
        for (int kk = tid; kk < n; kk += nt) {
            const double lam = M2[kk + (size_t)kk * ld2];
            double acc = 0.0;
            if (lam > eps) { const double *v = V2 + (size_t)kk * ld2; for (int i = 0; i < n; i++) acc += v[i] * br[i]; acc *= sqrt(1.0 / lam); }
            ro[kk] = acc;
        }
        if (sweeps && tid == 0) { sweeps[2 * w] = sw1; sweeps[2 * w + 1] = sw2; }
        __syncthreads();
    }
}

}  // namespace cerb
