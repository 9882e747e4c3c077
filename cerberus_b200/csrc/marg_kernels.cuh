// marg_kernels.cuh -- the dense tail of MarginalizationInfo::marginalize() on the device, one CTA per window:
//   reference src/factor/marginalization_factor.cpp:281-305
//     Amm = 0.5 (A_mm + A_mm^T); SelfAdjointEigenSolver(Amm); Amm_inv = V diag(lambda > eps ? 1 / lambda : 0) V^T
//     A   = Arr - Arm Amm_inv Amr;  b = brr - Arm Amm_inv bmm
//     SelfAdjointEigenSolver(A) (lower triangle);  S = lambda > eps ? lambda : 0;  S_inv = lambda > eps ? 1 / lambda : 0
//     linearized_jacobians = diag(sqrt S) V^T;  linearized_residuals = diag(sqrt S_inv) V^T b
// The eigen-solver is a parallel two-sided Jacobi iteration (round-robin pair ordering: kp / 2 disjoint rotations per round,
// column phase, row phase, re-symmetrisation), in fp64, on matrices that live in L2 (m = 19 + tracks anchored at frame 0 -> 169 x 169 for
// the 150-feature configuration, 86 x 86 for the kept block).  Rotation formulas as in the CPU restatement
// (oracle/ref_math.h sym_eig_jacobi), different pair order; the rows of linearized_jacobians come out in the order the
// eigenvalues sit on the diagonal (J^T J and J^T r, the only things MarginalizationFactor uses, do not depend on it).
#pragma once
#include "compat.h"

namespace cerb {

#if defined(CERB_CUSIM)
constexpr int MARG_THREADS = 64;      // the CPU simulator pays ~ one futex wake-up per thread and barrier; two warps still exercise every path
#else
constexpr int MARG_THREADS = 256;
#endif
constexpr int MARG_MAX_SWEEPS = 60;

// per-CTA workspace in doubles: M1, V1 (m x m) | Y, X (m x (n + 1)) | T2, M2, V2 (n x n) | br (n)
CERB_HD size_t marg_ws_doubles(int m, int n) {
    return 2 * (size_t)m * m + 2 * (size_t)m * (n + 1) + 3 * (size_t)n * n + (size_t)n;
}
CERB_HD size_t marg_smem_bytes(int m, int n) {           // (c, s) per concurrent rotation + the sweep flag
    const int k = m > n ? m : n;
    return (size_t)(k + 1) * sizeof(double) + 16;      // flag[2] behind the (c, s) pairs
}

// pair t (0 .. kp / 2 - 1) of round r (0 .. kp - 2) of a round-robin tournament over kp (even) players, p < q
CERB_D void jacobi_pair(int t, int r, int kp, int &p, int &q) {
    const int md = kp - 1;
    int a, b;
    if (t == 0) { a = md; b = r; }
    else { a = (r + t) % md; b = (r - t + md) % md; }
    p = a < b ? a : b; q = a < b ? b : a;
}

// rotation angles of round r from the upper triangle of M as it stands; (c, s) -> cs, a non-trivial rotation raises *flag
CERB_D void jacobi_angles(const double *M, int k, int kp, int r, double *cs, int *flag) {
    for (int t = threadIdx.x; t < kp / 2; t += blockDim.x) {
        int p, q; jacobi_pair(t, r, kp, p, q);
        double c = 1.0, s = 0.0;
        if (q < k) {
            const double apq = M[p + (size_t)q * k], app = M[p + (size_t)p * k], aqq = M[q + (size_t)q * k];
            if (apq != 0.0 && fabs(apq) > 1e-15 * sqrt(fabs(app * aqq))) {
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                c = 1.0 / sqrt(tt * tt + 1.0); s = tt * c;
                if (s != 0.0) *flag = 1;
            }
        }
        cs[2 * t] = c; cs[2 * t + 1] = s;
    }
}

// Eigen-decomposition of the symmetric k x k matrix M (column-major, leading dimension k): on return the eigenvalues are on the
// diagonal of M and the eigenvectors are the columns of V.  Called by all threads of the CTA; returns the number of sweeps.
// Three CTA barriers per round: column pass | row pass | re-symmetrisation together with the angles of the next round (both only
// read the upper triangle the row pass left).  flag[sweep & 1] collects "some rotation was non-trivial" for that sweep.
CERB_D int jacobi_eig(double *M, double *V, int k, double *cs, int *flag) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
    const int kp = k + (k & 1), half = kp / 2, rounds = kp - 1;
    for (int i = tid; i < k * k; i += nt) V[i] = (i / k == i % k) ? 1.0 : 0.0;
    if (tid == 0) { flag[0] = 0; flag[1] = 0; }
    __syncthreads();
    jacobi_angles(M, k, kp, 0, cs, flag);
    __syncthreads();
    int sweeps = 0;
    for (; sweeps < MARG_MAX_SWEEPS; sweeps++) {
        for (int r = 0; r < rounds; r++) {
            for (int t = wid; t < half; t += nwarp) {         // columns p, q of M and V
                const double c = cs[2 * t], s = cs[2 * t + 1];
                if (s == 0.0) continue;
                int p, q; jacobi_pair(t, r, kp, p, q);
                double *mp = M + (size_t)p * k, *mq = M + (size_t)q * k, *vp = V + (size_t)p * k, *vq = V + (size_t)q * k;
                for (int i = lane; i < k; i += 32) {
                    const double a = mp[i], b = mq[i], e = vp[i], f = vq[i];
                    mp[i] = c * a - s * b; mq[i] = s * a + c * b;
                    vp[i] = c * e - s * f; vq[i] = s * e + c * f;
                }
            }
            __syncthreads();
            for (int t = wid; t < half; t += nwarp) {         // rows p, q of M
                const double c = cs[2 * t], s = cs[2 * t + 1];
                if (s == 0.0) continue;
                int p, q; jacobi_pair(t, r, kp, p, q);
                for (int j = lane; j < k; j += 32) {
                    const double a = M[p + (size_t)j * k], b = M[q + (size_t)j * k];
                    M[p + (size_t)j * k] = c * a - s * b; M[q + (size_t)j * k] = s * a + c * b;
                }
            }
            // every thread has read last sweep's verdict by now (it did so before this sweep's first column pass)
            if (r == 0 && tid == 0) flag[(sweeps + 1) & 1] = 0;
            __syncthreads();
            // keep M exactly symmetric: the column and the row pass round differently, and an asymmetric residue of eps |M| is enough to
            // keep the null space of a rank-deficient Schur complement rotating for ever (the angles are taken from the upper triangle)
            for (int e = tid; e < k * k; e += nt) { const int i = e % k, j = e / k; if (i > j) M[e] = M[j + (size_t)i * k]; }
            if (r + 1 < rounds) jacobi_angles(M, k, kp, r + 1, cs, flag + (sweeps & 1));
            else jacobi_angles(M, k, kp, 0, cs, flag + ((sweeps + 1) & 1));
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
    double *cs = sm; int *flag = reinterpret_cast<int *>(sm + kmax + 1);
    for (int w = blockIdx.x; w < n_windows; w += gridDim.x) {
        const int m = dims ? dims[4 * w] : mmax, n = dims ? dims[4 * w + 1] : nmax;
        if (dims && dims[4 * w + 2] != 1) continue;
        const int pos = m + n, nc = n + 1;
        double *M1 = ws_all + (size_t)blockIdx.x * marg_ws_doubles(mmax, nmax), *V1 = M1 + (size_t)m * m, *Y = V1 + (size_t)m * m, *X = Y + (size_t)m * nc;
        double *T2 = X + (size_t)m * nc, *M2 = T2 + (size_t)n * n, *V2 = M2 + (size_t)n * n, *br = V2 + (size_t)n * n;
        const double *A = A_all + (size_t)w * (A_stride ? A_stride : (long)pos * pos), *b = b_all + (size_t)w * (b_stride ? b_stride : (long)pos);
        for (int e = tid; e < m * m; e += nt) { const int i = e % m, j = e / m; M1[e] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]); }
        __syncthreads();
        const int sw1 = jacobi_eig(M1, V1, m, cs, flag);
        // Y = diag(inv) V1^T [Amr | bm]
        for (int e = tid; e < m * nc; e += nt) {
            const int c = e % nc, kk = e / nc;
            const double lam = M1[kk + (size_t)kk * m];
            double acc = 0.0;
            if (lam > eps) {
                const double *v = V1 + (size_t)kk * m;
                if (c < n) for (int i = 0; i < m; i++) acc += v[i] * A[(size_t)i * pos + m + c];
                else for (int i = 0; i < m; i++) acc += v[i] * b[i];
                acc *= 1.0 / lam;
            }
            Y[kk + (size_t)c * m] = acc;
        }
        __syncthreads();
        // X = V1 Y = Amm_inv [Amr | bm]
        for (int e = tid; e < m * nc; e += nt) {
            const int i = e % m, c = e / m;
            const double *y = Y + (size_t)c * m;
            double acc = 0.0;
            for (int kk = 0; kk < m; kk++) acc += V1[i + (size_t)kk * m] * y[kk];
            X[i + (size_t)c * m] = acc;
        }
        __syncthreads();
        // [Ar | br] = [Arr | brr] - Arm X
        for (int e = tid; e < n * nc; e += nt) {
            const int c = e % nc, r = e / nc;
            const double *arow = A + (size_t)(m + r) * pos, *x = X + (size_t)c * m;
            double acc = 0.0;
            for (int i = 0; i < m; i++) acc += arow[i] * x[i];
            if (c < n) T2[r + (size_t)c * n] = arow[m + c] - acc; else br[r] = b[m + r] - acc;
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += nt) { const int i = e % n, j = e / n; M2[e] = i >= j ? T2[i + (size_t)j * n] : T2[j + (size_t)i * n]; }
        __syncthreads();
        const int sw2 = jacobi_eig(M2, V2, n, cs, flag);
        double *Jo = lin_J + (size_t)w * (J_stride ? J_stride : (long)n * n), *ro = lin_r + (size_t)w * (r_stride ? r_stride : (long)n);
        for (int e = tid; e < n * n; e += nt) {
            const int kk = e % n, j = e / n;
            const double lam = M2[kk + (size_t)kk * n];
            Jo[e] = lam > eps ? sqrt(lam) * V2[j + (size_t)kk * n] : 0.0;
        }
        for (int kk = tid; kk < n; kk += nt) {
            const double lam = M2[kk + (size_t)kk * n];
            double acc = 0.0;
            if (lam > eps) { const double *v = V2 + (size_t)kk * n; for (int i = 0; i < n; i++) acc += v[i] * br[i]; acc *= sqrt(1.0 / lam); }
            ro[kk] = acc;
        }
        if (sweeps && tid == 0) { sweeps[2 * w] = sw1; sweeps[2 * w + 1] = sw2; }
        __syncthreads();
    }
}

}  // namespace cerb
