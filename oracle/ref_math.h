// oracle/ref_math.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Minimal fp64 linear algebra for the CPU restatement of the Cerberus hot path.  The reference
// uses Eigen 3.3.4 (not available in this container); the handful of Eigen operations its factor
// code relies on are restated here with the same formulas so that round-off behaves alike:
//   Quaterniond * Vector3d       -> Eigen QuaternionBase::_transformVector
//   Quaterniond::toRotationMatrix, inverse() (= conjugate / squaredNorm), normalized()
//   Utility::{deltaQ, skewSymmetric, Qleft, Qright, R2ypr, ypr2R}   src/utils/utility.h:28-125
// PARITY UNPINNED: the reference ships no golden vectors for this path.
#pragma once
#include "sym_eig_qr.h"
#include <cmath>
#include <cstring>
#include <vector>
#include <cassert>

namespace oracle {

struct V3 {
    double x = 0, y = 0, z = 0;
    V3() {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit V3(const double *p) : x(p[0]), y(p[1]), z(p[2]) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 {
    double m[3][3];
    M3() { std::memset(m, 0, sizeof(m)); }
    static M3 identity() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
    double operator()(int i, int j) const { return m[i][j]; }
    double &operator()(int i, int j) { return m[i][j]; }
};
inline M3 operator*(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
    return r;
}
inline V3 operator*(const M3 &a, V3 v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline M3 operator*(double s, const M3 &a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j]; return r; }
inline M3 operator*(const M3 &a, double s) { return s * a; }
inline M3 operator+(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
inline M3 operator-(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
inline M3 operator-(const M3 &a) { return -1.0 * a; }
inline M3 transpose(const M3 &a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
// Utility::skewSymmetric, utility.h:43-51
inline M3 skew(V3 q) {
    M3 r;
    r.m[0][1] = -q.z; r.m[0][2] = q.y;
    r.m[1][0] = q.z;  r.m[1][2] = -q.x;
    r.m[2][0] = -q.y; r.m[2][1] = q.x;
    return r;
}

struct Quat {
    double w = 1, x = 0, y = 0, z = 0;
    Quat() {}
    Quat(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}  // Eigen ctor order
    V3 vec() const { return {x, y, z}; }
};
inline Quat operator*(const Quat &a, const Quat &b) {  // Eigen quat product
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
            a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline V3 operator*(const Quat &q, V3 v) {  // Eigen _transformVector
    V3 uv = cross(q.vec(), v);
    uv = uv + uv;
    return v + q.w * uv + cross(q.vec(), uv);
}
inline double sqnorm(const Quat &q) { return q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; }
inline Quat inverse(const Quat &q) {  // Eigen: conjugate / squaredNorm
    double n2 = sqnorm(q);
    return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}
inline Quat normalized(const Quat &q) {
    double n = std::sqrt(sqnorm(q));
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}
inline M3 toR(const Quat &q) {  // Eigen toRotationMatrix
    M3 r;
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz;       r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;       r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;       r.m[2][1] = tyz + twx;       r.m[2][2] = 1 - (txx + tyy);
    return r;
}
// Eigen Quaternion(Matrix3) constructor (used by vector2double, estimator.cpp:855)
inline Quat fromR(const M3 &mat) {
    Quat q;
    double t = mat(0, 0) + mat(1, 1) + mat(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t; t = 0.5 / t;
        q.x = (mat(2, 1) - mat(1, 2)) * t; q.y = (mat(0, 2) - mat(2, 0)) * t; q.z = (mat(1, 0) - mat(0, 1)) * t;
    } else {
        int i = 0;
        if (mat(1, 1) > mat(0, 0)) i = 1;
        if (mat(2, 2) > mat(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t; t = 0.5 / t;
        q.w = (mat(k, j) - mat(j, k)) * t;
        v[j] = (mat(j, i) + mat(i, j)) * t;
        v[k] = (mat(k, i) + mat(i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
// Utility::deltaQ, utility.h:28-41 (NOT normalised)
inline Quat deltaQ(V3 theta) { return {1.0, theta.x / 2.0, theta.y / 2.0, theta.z / 2.0}; }

// Utility::Qleft / Qright bottom-right 3x3 corners (utility.h:63-83): the only part the factors use.
inline M3 QleftBR(const Quat &q) { return q.w * M3::identity() + skew(q.vec()); }
inline M3 QrightBR(const Quat &q) { return q.w * M3::identity() - skew(q.vec()); }
// (Qleft(a) * Qright(b)).bottomRightCorner<3,3>() : rows 1..3 of Qleft times cols 1..3 of Qright
inline M3 QleftQrightBR(const Quat &a, const Quat &b) {
    // Qleft(a) = [a.w, -a.v^T; a.v, a.w I + [a.v]x], Qright(b) = [b.w, -b.v^T; b.v, b.w I - [b.v]x]
    M3 r = QleftBR(a) * QrightBR(b);
    V3 av = a.vec(), bv = b.vec();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] += av[i] * (-bv[j]);
    return r;
}

// Utility::R2ypr / ypr2R, utility.h:85-125 (degrees)
inline V3 R2ypr(const M3 &R) {
    V3 n(R(0, 0), R(1, 0), R(2, 0)), o(R(0, 1), R(1, 1), R(2, 1)), a(R(0, 2), R(1, 2), R(2, 2));
    double y = std::atan2(n.y, n.x);
    double p = std::atan2(-n.z, n.x * std::cos(y) + n.y * std::sin(y));
    double r = std::atan2(a.x * std::sin(y) - a.y * std::cos(y), -o.x * std::sin(y) + o.y * std::cos(y));
    return V3(y, p, r) / M_PI * 180.0;
}
inline M3 ypr2R(V3 ypr) {
    double y = ypr.x / 180.0 * M_PI, p = ypr.y / 180.0 * M_PI, r = ypr.z / 180.0 * M_PI;
    M3 Rz, Ry, Rx;
    Rz(0, 0) = std::cos(y); Rz(0, 1) = -std::sin(y); Rz(1, 0) = std::sin(y); Rz(1, 1) = std::cos(y); Rz(2, 2) = 1;
    Ry(0, 0) = std::cos(p); Ry(0, 2) = std::sin(p); Ry(1, 1) = 1; Ry(2, 0) = -std::sin(p); Ry(2, 2) = std::cos(p);
    Rx(0, 0) = 1; Rx(1, 1) = std::cos(r); Rx(1, 2) = -std::sin(r); Rx(2, 1) = std::sin(r); Rx(2, 2) = std::cos(r);
    return Rz * Ry * Rx;
}

// ---- small dynamic row-major matrix -------------------------------------------------------------
struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return d[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
    void setZero() { std::fill(d.begin(), d.end(), 0.0); }
    static Mat identity(int n) { Mat m(n, n); for (int i = 0; i < n; i++) m(i, i) = 1; return m; }
    void setBlock(int i0, int j0, const M3 &b) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) (*this)(i0 + i, j0 + j) = b(i, j); }
    M3 block3(int i0, int j0) const { M3 b; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) b(i, j) = (*this)(i0 + i, j0 + j); return b; }
};
inline Mat matmul(const Mat &a, const Mat &b) {
    assert(a.c == b.r);
    Mat o(a.r, b.c);
    for (int i = 0; i < a.r; i++)
        for (int k = 0; k < a.c; k++) {
            double aik = a(i, k);
            if (aik == 0.0) continue;
            const double *bp = &b.d[(size_t)k * b.c];
            double *op = &o.d[(size_t)i * o.c];
            for (int j = 0; j < b.c; j++) op[j] += aik * bp[j];
        }
    return o;
}
inline Mat transpose(const Mat &a) { Mat o(a.c, a.r); for (int i = 0; i < a.r; i++) for (int j = 0; j < a.c; j++) o(j, i) = a(i, j); return o; }

// In-place Cholesky A = L L^T on the lower triangle (Eigen LLT semantics: fails on a non-positive
// pivot).  Returns false on failure.
inline bool cholesky_lower(Mat &A) {
    int n = A.r;
    for (int j = 0; j < n; j++) {
        double s = A(j, j);
        for (int k = 0; k < j; k++) s -= A(j, k) * A(j, k);
        if (!(s > 0.0)) return false;
        double ljj = std::sqrt(s);
        A(j, j) = ljj;
        for (int i = j + 1; i < n; i++) {
            double t = A(i, j);
            const double *ai = &A.d[(size_t)i * n], *aj = &A.d[(size_t)j * n];
            for (int k = 0; k < j; k++) t -= ai[k] * aj[k];
            A(i, j) = t / ljj;
        }
    }
    return true;
}
inline void chol_solve_inplace(const Mat &L, double *b) {  // solves L L^T x = b
    int n = L.r;
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L(i, k) * b[k]; b[i] = s / L(i, i); }
    for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= L(k, i) * b[k]; b[i] = s / L(i, i); }
}

// Inverse by LU with partial pivoting (what Eigen's MatrixBase::inverse() does for sizes > 4,
// PartialPivLU); returns false if a zero pivot is hit.
inline bool inverse_partial_piv_lu(const Mat &Ain, Mat &inv) {
    int n = Ain.r;
    Mat A = Ain;
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int k = 0; k < n; k++) {
        int p = k; double best = std::fabs(A(k, k));
        for (int i = k + 1; i < n; i++) if (std::fabs(A(i, k)) > best) { best = std::fabs(A(i, k)); p = i; }
        if (best == 0.0) return false;
        if (p != k) { for (int j = 0; j < n; j++) std::swap(A(k, j), A(p, j)); std::swap(perm[k], perm[p]); }
        for (int i = k + 1; i < n; i++) {
            A(i, k) /= A(k, k);
            double f = A(i, k);
            for (int j = k + 1; j < n; j++) A(i, j) -= f * A(k, j);
        }
    }
    inv = Mat(n, n);
    std::vector<double> col(n);
    for (int c = 0; c < n; c++) {
        for (int i = 0; i < n; i++) col[i] = (perm[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i < n; i++) { double s = col[i]; for (int k = 0; k < i; k++) s -= A(i, k) * col[k]; col[i] = s; }
        for (int i = n - 1; i >= 0; i--) { double s = col[i]; for (int k = i + 1; k < n; k++) s -= A(i, k) * col[k]; col[i] = s / A(i, i); }
        for (int i = 0; i < n; i++) inv(i, c) = col[i];
    }
    return true;
}

// Symmetric eigendecomposition by cyclic Jacobi (stands in for Eigen::SelfAdjointEigenSolver in
// marginalization_factor.cpp:281,297).  A is symmetric n x n; on return evals[i], evecs column i.
inline void sym_eig_jacobi(const Mat &Ain, std::vector<double> &evals, Mat &V) {
    int n = Ain.r;
    Mat A = Ain;
    V = Mat::identity(n);
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0, diag = 0;
        for (int i = 0; i < n; i++) { diag += A(i, i) * A(i, i); for (int j = i + 1; j < n; j++) off += A(i, j) * A(i, j); }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A(p, q);
                if (apq == 0.0) continue;
                double theta = (A(q, q) - A(p, p)) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) { double akp = A(k, p), akq = A(k, q); A(k, p) = c * akp - s * akq; A(k, q) = s * akp + c * akq; }
                for (int k = 0; k < n; k++) { double apk = A(p, k), aqk = A(q, k); A(p, k) = c * apk - s * aqk; A(q, k) = s * apk + c * aqk; }
                for (int k = 0; k < n; k++) { double vkp = V(k, p), vkq = V(k, q); V(k, p) = c * vkp - s * vkq; V(k, q) = s * vkp + c * vkq; }
            }
    }
    evals.resize(n);
    for (int i = 0; i < n; i++) evals[i] = A(i, i);
}

// Eigen::SelfAdjointEigenSolver as the reference uses it (marginalization_factor.cpp:281,297): Householder tridiagonalisation + implicit QR,
// eigenvalues ascending (oracle/sym_eig_qr.h restates Eigen 3.3.x's algorithm).  eig_mode() == 1 switches to the cyclic Jacobi solver above
// (what the device kernels use), kept so that the sensitivity of the marginalization prior to the eigen-solver can be measured.
inline int &eig_mode() { static int mode = 0; return mode; }
inline void sym_eig(const Mat &A, std::vector<double> &evals, Mat &V) {
    if (eig_mode() == 1) { sym_eig_jacobi(A, evals, V); return; }
    const int n = A.r;
    evals.resize(n); std::vector<double> Q((size_t)n * n);
    symeig::tridiag_qr(n, A.d.data(), n, 1, evals.data(), Q.data());          // row-major A: (i, j) at i * n + j
    V = Mat(n, n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V(i, j) = Q[(size_t)j * n + i];
}

}  // namespace oracle
