// feature_kernels.cuh -- the per-feature steps either side of the solve, batched over the resident
// windows: one thread per (window, feature), streaming the observation planes (HBM / L2 bound, 0.5 KB per feature).
//   outlier_error_kernel   Estimator::outliersRejection + reprojectionError   src/estimator/estimator.cpp:1729-1798
//   triangulate_kernel     FeatureManager::triangulate + triangulatePoint      src/featureTracker/feature_manager.cpp:198-212,302-385
//   shift_depth_kernel     FeatureManager::removeBackShiftDepth (MARGIN_OLD slide)  src/featureTracker/feature_manager.cpp:450-488, estimator.cpp:1660-1677
#pragma once
#include "eval_kernels.cuh"

namespace cerb {

// Estimator::reprojectionError (estimator.cpp:1729-1739): anchor camera 0 of frame i -> camera (ricj, ticj) of frame j
CERB_HD double reprojection_error(const m33 &Ri, d3 Pi, const m33 &rici, d3 tici, const m33 &Rj, d3 Pj, const m33 &ricj, d3 ticj,
                                  double depth, d3 uvi, double ujx, double ujy) {
    const d3 pts_w = mv33(Ri, mv33(rici, depth * uvi) + tici) + Pi;
    const d3 pts_cj = mTv33(ricj, mTv33(Rj, pts_w - Pj) - ticj);
    const double rx = pts_cj.x / pts_cj.z - ujx, ry = pts_cj.y / pts_cj.z - ujy;
    return sqrt(rx * rx + ry * ry);
}

// ave_err[w * maxF + f] = mean reprojection error of feature f (device order) over its K1 / K2 / K3 observations
CERB_GLOBAL void outlier_error_kernel(int n_windows, int maxF, int maxObs, const int *n_features, const int *feat_start, const int *feat_nobs,
                                      const int *feat_off, const double *obs, const int *obs_stereo, const double *state, const double *lam, double *ave_err) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = idx / maxF, f = idx % maxF;
    if (w >= n_windows || f >= n_features[w]) return;
    const double *x = state + (size_t)w * ST_STRIDE;
    const double *ob = obs + (size_t)w * 9 * maxObs;
    const int *st = obs_stereo + (size_t)w * maxObs;
    const int i = feat_start[(size_t)w * maxF + f], nobs = feat_nobs[(size_t)w * maxF + f], off = feat_off[(size_t)w * maxF + f];
    const m33 Ri = qtoR(ldq(x + ST_POSE + 7 * i + 3)); const d3 Pi = ld3(x + ST_POSE + 7 * i);
    const m33 ric0 = qtoR(ldq(x + ST_EX + 3)), ric1 = qtoR(ldq(x + ST_EX + 7 + 3));
    const d3 tic0 = ld3(x + ST_EX), tic1 = ld3(x + ST_EX + 7);
    const d3 uvi = mk3(ob[0 * maxObs + off], ob[1 * maxObs + off], 1.0);
    const double depth = 1.0 / lam[(size_t)w * maxF + f];                     // estimated_depth (feature_manager.cpp:189)
    double err = 0.0; int cnt = 0;
    for (int k = 0; k < nobs; k++) {
        const int j = i + k, o = off + k;
        const m33 Rj = qtoR(ldq(x + ST_POSE + 7 * j + 3)); const d3 Pj = ld3(x + ST_POSE + 7 * j);
        if (k != 0) { err += reprojection_error(Ri, Pi, ric0, tic0, Rj, Pj, ric0, tic0, depth, uvi, ob[0 * maxObs + o], ob[1 * maxObs + o]); cnt++; }
        if (st[o]) { err += reprojection_error(Ri, Pi, ric0, tic0, Rj, Pj, ric1, tic1, depth, uvi, ob[4 * maxObs + o], ob[5 * maxObs + o]); cnt++; }   // both branches of :1771-1788 are the same call
    }
    ave_err[(size_t)w * maxF + f] = err / cnt;
}

// Right singular vector of the smallest singular value of a 4 x 4 matrix by one-sided (Hestenes) Jacobi rotations of its
// columns: A V = U S.  FeatureManager::triangulatePoint takes design_matrix.jacobiSvd(ComputeFullV).matrixV().rightCols<1>();
// only ratios of the components are used, so the sign convention is irrelevant.
CERB_HD void svd4_null_vector(double A[16], double v[4]) {
    double V[16];
    for (int k = 0; k < 16; k++) V[k] = (k % 5 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int r = 0; r < 4; r++) { al += A[4 * r + p] * A[4 * r + p]; be += A[4 * r + q] * A[4 * r + q]; ga += A[4 * r + p] * A[4 * r + q]; }
                if (ga == 0.0) continue;
                off = fmax(off, fabs(ga) / sqrt(al * be + 1e-300));
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < 4; r++) {
                    const double ap = A[4 * r + p], aq = A[4 * r + q]; A[4 * r + p] = c * ap - s * aq; A[4 * r + q] = s * ap + c * aq;
                    const double vp = V[4 * r + p], vq = V[4 * r + q]; V[4 * r + p] = c * vp - s * vq; V[4 * r + q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    int best = 0; double bn = 1e300;
    for (int c = 0; c < 4; c++) { double n2 = 0.0; for (int r = 0; r < 4; r++) n2 += A[4 * r + c] * A[4 * r + c]; if (n2 < bn) { bn = n2; best = c; } }
    for (int r = 0; r < 4; r++) v[r] = V[4 * r + best];
}

// depth of the anchor observation from two views (feature_manager.cpp:302-385): camera poses as [R^T | -R^T t] rows
CERB_HD double triangulate_two_view(const m33 &R0, d3 t0, const m33 &R1, d3 t1, double u0x, double u0y, double u1x, double u1y, double init_depth) {
    double P0[12], P1[12];                                                       // 3 x 4 row-major: leftCols = R^T, rightCols = -R^T t
    const d3 m0 = -mTv33(R0, t0), m1 = -mTv33(R1, t1);
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) { P0[4 * r + c] = R0.m[3 * c + r]; P1[4 * r + c] = R1.m[3 * c + r]; }
        P0[4 * r + 3] = get3(m0, r); P1[4 * r + 3] = get3(m1, r);
    }
    double A[16], v[4];
    for (int c = 0; c < 4; c++) {                                               // triangulatePoint, feature_manager.cpp:201-205
        A[c] = u0x * P0[8 + c] - P0[c]; A[4 + c] = u0y * P0[8 + c] - P0[4 + c];
        A[8 + c] = u1x * P1[8 + c] - P1[c]; A[12 + c] = u1y * P1[8 + c] - P1[4 + c];
    }
    svd4_null_vector(A, v);
    const d3 p = mk3(v[0] / v[3], v[1] / v[3], v[2] / v[3]);
    const double depth = P0[8] * p.x + P0[9] * p.y + P0[10] * p.z + P0[11];     // (leftPose * point).z()
    return depth > 0.0 ? depth : init_depth;
}

// depth[w * maxF + f]: for features whose para_Feature <= 0 (estimated_depth <= 0: not yet triangulated) the two-view depth --
// left/right cameras of the anchor frame if its observation is stereo, else camera 0 of the anchor frame and the next frame --
// otherwise the current 1 / para_Feature.  (The multi-frame SVD branch :387-428 is unreachable: size() > 1 always takes :351.)
CERB_GLOBAL void triangulate_kernel(int n_windows, int maxF, int maxObs, const int *n_features, const int *feat_start, const int *feat_nobs,
                                    const int *feat_off, const double *obs, const int *obs_stereo, const double *state, const double *lam,
                                    double init_depth, double *depth) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = idx / maxF, f = idx % maxF;
    if (w >= n_windows || f >= n_features[w]) return;
    const double l = lam[(size_t)w * maxF + f];
    double out = 1.0 / l;
    if (!(l > 0.0)) {
        const double *x = state + (size_t)w * ST_STRIDE;
        const double *ob = obs + (size_t)w * 9 * maxObs;
        const int i = feat_start[(size_t)w * maxF + f], nobs = feat_nobs[(size_t)w * maxF + f], off = feat_off[(size_t)w * maxF + f];
        const m33 Ri = qtoR(ldq(x + ST_POSE + 7 * i + 3)); const d3 Pi = ld3(x + ST_POSE + 7 * i);
        const m33 ric0 = qtoR(ldq(x + ST_EX + 3)); const d3 tic0 = ld3(x + ST_EX);
        const m33 R0 = mul33(Ri, ric0); const d3 t0 = Pi + mv33(Ri, tic0);
        if (obs_stereo[(size_t)w * maxObs + off]) {
            const m33 ric1 = qtoR(ldq(x + ST_EX + 7 + 3)); const d3 tic1 = ld3(x + ST_EX + 7);
            out = triangulate_two_view(R0, t0, mul33(Ri, ric1), Pi + mv33(Ri, tic1), ob[0 * maxObs + off], ob[1 * maxObs + off], ob[4 * maxObs + off], ob[5 * maxObs + off], init_depth);
        } else if (nobs > 1) {
            const m33 Rj = qtoR(ldq(x + ST_POSE + 7 * (i + 1) + 3)); const d3 Pj = ld3(x + ST_POSE + 7 * (i + 1));
            out = triangulate_two_view(R0, t0, mul33(Rj, ric0), Pj + mv33(Rj, tic0), ob[0 * maxObs + off], ob[1 * maxObs + off], ob[0 * maxObs + off + 1], ob[1 * maxObs + off + 1], init_depth);
        } else out = l;                                                           // left untouched by the reference
    }
    depth[(size_t)w * maxF + f] = out;
}

// Depth bookkeeping of slideWindowOld(): the oldest frame leaves the window.  Tracks anchored later just move one frame down
// (start_frame - 1, depth unchanged); tracks anchored at frame 0 lose their first observation, are erased if fewer than 2 remain
// (keep = 0), else their depth is re-expressed in camera 0 of the new anchor (old frame 1): dep_j = (R1^T (R0 (uv_i depth) + P0 - P1)).z
// with R0 = Rs[0] ric0, P0 = Ps[0] + Rs[0] tic0, R1 = Rs[1] ric0, P1 = Ps[1] + Rs[1] tic0; dep_j <= 0 becomes init_depth.
// out [3][n_windows * maxF]: new start_frame, new estimated_depth, keep flag.
CERB_GLOBAL void shift_depth_kernel(int n_windows, int maxF, int maxObs, const int *n_features, const int *feat_start, const int *feat_nobs,
                                    const int *feat_off, const double *obs, const double *state, const double *lam, double init_depth, double *out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = idx / maxF, f = idx % maxF;
    if (w >= n_windows || f >= n_features[w]) return;
    const size_t N = (size_t)n_windows * maxF, k = (size_t)w * maxF + f;
    const int start = feat_start[k], nobs = feat_nobs[k], off = feat_off[k];
    const double depth = 1.0 / lam[k];
    if (start != 0) { out[k] = (double)(start - 1); out[N + k] = depth; out[2 * N + k] = 1.0; return; }
    out[k] = 0.0;
    if (nobs - 1 < 2) { out[N + k] = depth; out[2 * N + k] = 0.0; return; }
    const double *x = state + (size_t)w * ST_STRIDE;
    const double *ob = obs + (size_t)w * 9 * maxObs;
    const m33 Rs0 = qtoR(ldq(x + ST_POSE + 3)), Rs1 = qtoR(ldq(x + ST_POSE + 7 + 3)), ric0 = qtoR(ldq(x + ST_EX + 3));
    const d3 Ps0 = ld3(x + ST_POSE), Ps1 = ld3(x + ST_POSE + 7), tic0 = ld3(x + ST_EX);
    const m33 R0 = mul33(Rs0, ric0), R1 = mul33(Rs1, ric0);
    const d3 P0 = Ps0 + mv33(Rs0, tic0), P1 = Ps1 + mv33(Rs1, tic0);
    const d3 pts_i = depth * mk3(ob[0 * maxObs + off], ob[1 * maxObs + off], 1.0);
    const d3 pts_j = mTv33(R1, (mv33(R0, pts_i) + P0) - P1);
    out[N + k] = pts_j.z > 0.0 ? pts_j.z : init_depth;
    out[2 * N + k] = 1.0;
}

}  // namespace cerb
