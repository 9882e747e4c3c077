// oracle/sym_eig_qr.h -- TEST INFRASTRUCTURE.
//
// Symmetric eigen-decomposition the way Eigen 3.3.x's SelfAdjointEigenSolver::compute() does it, restated from the
// published algorithm (Eigen is a third-party dependency of the reference, pinned to libeigen3-dev 3.3.4 by
// .devcontainer/Dockerfile, and is absent from this image):
//   1. copy the LOWER triangle, scale it by 1 / max|a_ij|;
//   2. Householder tridiagonalisation in place (internal::tridiagonalization_inplace, Tridiagonalization.h): for every column
//      the reflector of its sub-diagonal part (makeHouseholderInPlace: beta = -sign(c0) ||x||, essential = tail / (c0 - beta),
//      tau = (beta - c0) / beta), p = tau A v, p -= (tau / 2)(p . v) v, A -= v p^T + p v^T on the lower triangle;
//   3. Q = H_0 H_1 ... H_{n-2} accumulated from the last reflector backwards (HouseholderSequence::evalTo);
//   4. implicit symmetric QR steps with the Wilkinson shift on the largest unreduced block (internal::tridiagonal_qr_step,
//      deflation test |e_i| <= 2 eps (|d_i| + |d_i+1|)), Givens rotations applied to Q from the right, <= 30 n steps;
//   5. eigenvalues sorted ascending by selection (eigenvector columns swapped along), scaled back.
// Used by the Eigen shim (oracle/shim/Eigen/Dense: what the reference's MarginalizationInfo::marginalize() runs on in
// oracle/_ref) and by the oracle restatement (oracle/ref_window.cpp).  Summation order inside the matrix-vector products is
// plain left-to-right (Eigen's kernels are vectorised), so results agree with an Eigen build to rounding, not bit for bit.
#pragma once
#include <vector>
#include <cmath>
#include <limits>
#include <algorithm>

namespace symeig {

// A: n x n, element (i, j) at A[i * lda_r + j * lda_c] (any layout); only i >= j is read.  evals[n] ascending; Q column-major
// n x n (column k = eigenvector k).  Returns false on NoConvergence.
inline bool tridiag_qr(int n, const double *A, long lda_r, long lda_c, double *evals, double *Q) {
    if (n <= 0) return true;
    if (n == 1) { evals[0] = A[0]; Q[0] = 1.0; return true; }
    std::vector<double> M((size_t)n * n, 0.0);                       // column-major working copy, lower triangle
    auto m = [&](int i, int j) -> double & { return M[(size_t)j * n + i]; };
    double scale = 0.0;
    for (int j = 0; j < n; j++) for (int i = j; i < n; i++) { const double v = A[i * lda_r + j * lda_c]; m(i, j) = v; scale = std::max(scale, std::fabs(v)); }
    if (scale == 0.0) scale = 1.0;
    for (int j = 0; j < n; j++) for (int i = j; i < n; i++) m(i, j) /= scale;
    std::vector<double> h(n - 1, 0.0), p(n, 0.0), diag(n), sub(n - 1);
    const double tol = std::numeric_limits<double>::min();
    // ---- tridiagonalisation
    for (int i = 0; i < n - 1; i++) {
        const int rem = n - i - 1;
        double *x = &m(i + 1, i);                                    // contiguous column tail
        double tail2 = 0.0; for (int k = 1; k < rem; k++) tail2 += x[k] * x[k];
        const double c0 = x[0];
        double tau, beta;
        if (rem == 1 || tail2 <= tol) { tau = 0.0; beta = c0; for (int k = 1; k < rem; k++) x[k] = 0.0; }
        else {
            beta = std::sqrt(c0 * c0 + tail2); if (c0 >= 0.0) beta = -beta;
            for (int k = 1; k < rem; k++) x[k] /= (c0 - beta);
            tau = (beta - c0) / beta;
        }
        x[0] = 1.0;
        // p = A22 (symmetric, lower stored) * (tau v)
        for (int r = 0; r < rem; r++) p[r] = 0.0;
        for (int c = 0; c < rem; c++) {
            const double tv = tau * x[c];
            p[c] += m(i + 1 + c, i + 1 + c) * tv;
            for (int r = c + 1; r < rem; r++) { const double a = m(i + 1 + r, i + 1 + c); p[r] += a * tv; p[c] += a * (tau * x[r]); }
        }
        double pv = 0.0; for (int r = 0; r < rem; r++) pv += p[r] * x[r];
        const double f = tau * -0.5 * pv;
        for (int r = 0; r < rem; r++) p[r] += f * x[r];
        for (int c = 0; c < rem; c++) for (int r = c; r < rem; r++) m(i + 1 + r, i + 1 + c) -= x[r] * p[c] + p[r] * x[c];
        x[0] = beta; h[i] = tau;
    }
    for (int i = 0; i < n; i++) diag[i] = m(i, i);
    for (int i = 0; i < n - 1; i++) sub[i] = m(i + 1, i);
    // ---- Q = H_0 ... H_{n-2}
    auto q = [&](int i, int j) -> double & { return Q[(size_t)j * n + i]; };
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) q(i, j) = (i == j) ? 1.0 : 0.0;
    std::vector<double> tmp(n);
    for (int k = n - 2; k >= 0; k--) {
        const int cs = n - k - 1, r0 = k + 1;                         // acts on Q(r0.., r0..)
        const double tau = h[k];
        if (cs == 1) { q(r0, r0) *= 1.0 - tau; continue; }
        if (tau == 0.0) continue;
        const double *ess = &m(k + 2, k);                              // essential part (cs - 1 entries)
        for (int c = 0; c < cs; c++) { double s = 0.0; for (int r = 1; r < cs; r++) s += ess[r - 1] * q(r0 + r, r0 + c); tmp[c] = s + q(r0, r0 + c); }
        for (int c = 0; c < cs; c++) { q(r0, r0 + c) -= tau * tmp[c]; for (int r = 1; r < cs; r++) q(r0 + r, r0 + c) -= tau * ess[r - 1] * tmp[c]; }
    }
    // ---- implicit QR on (diag, sub)
    int end = n - 1, start = 0, iter = 0;
    const double precision = 2.0 * std::numeric_limits<double>::epsilon();
    const int max_iter = 30 * n;
    while (end > 0) {
        for (int i = start; i < end; i++)
            if (std::fabs(sub[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision || std::fabs(sub[i]) <= tol) sub[i] = 0.0;
        while (end > 0 && sub[end - 1] == 0.0) end--;
        if (end <= 0) break;
        iter++;
        if (iter > max_iter) break;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.0) start--;
        // tridiagonal_qr_step
        const double td = (diag[end - 1] - diag[end]) * 0.5, e = sub[end - 1];
        double mu = diag[end];
        if (td == 0.0) mu -= std::fabs(e);
        else if (e != 0.0) {
            const double e2 = e * e, hh = std::hypot(td, e);
            if (e2 == 0.0) mu -= e / ((td + (td > 0.0 ? hh : -hh)) / e);
            else mu -= e2 / (td + (td > 0.0 ? hh : -hh));
        }
        double x = diag[start] - mu, z = sub[start];
        for (int k = start; k < end && z != 0.0; k++) {
            double c, s;                                              // JacobiRotation::makeGivens(x, z)
            if (z == 0.0) { c = x < 0.0 ? -1.0 : 1.0; s = 0.0; }
            else if (x == 0.0) { c = 0.0; s = z < 0.0 ? 1.0 : -1.0; }
            else if (std::fabs(x) > std::fabs(z)) { const double t = z / x; double u = std::sqrt(1.0 + t * t); if (x < 0.0) u = -u; c = 1.0 / u; s = -t * c; }
            else { const double t = x / z; double u = std::sqrt(1.0 + t * t); if (z < 0.0) u = -u; s = -1.0 / u; c = -t * s; }
            const double sdk = s * diag[k] + c * sub[k], dkp1 = s * sub[k] + c * diag[k + 1];
            diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
            diag[k + 1] = s * sdk + c * dkp1;
            sub[k] = c * sdk - s * dkp1;
            if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
            x = sub[k];
            if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
            for (int r = 0; r < n; r++) { const double a = q(r, k), b = q(r, k + 1); q(r, k) = c * a - s * b; q(r, k + 1) = s * a + c * b; }   // Q = Q G
        }
    }
    const bool ok = iter <= max_iter;
    if (ok)
        for (int i = 0; i < n - 1; i++) {
            int k = 0; for (int j = 1; j < n - i; j++) if (diag[i + j] < diag[i + k]) k = j;
            if (k > 0) { std::swap(diag[i], diag[k + i]); for (int r = 0; r < n; r++) std::swap(q(r, i), q(r, k + i)); }
        }
    for (int i = 0; i < n; i++) evals[i] = diag[i] * scale;
    return ok;
}

}  // namespace symeig
