// oracle/shim: parameters.h includes OpenCV for its YAML loader (not compiled here); feature_manager.cpp names a handful of
// cv:: types in its PnP initialisation (FeatureManager::solvePoseByPnP / initFramePoseByPnP, only used while the estimator is
// initialising, outside the hot path).  TEST INFRASTRUCTURE: just enough declarations for that file to compile; solvePnP aborts.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <set>
#include <map>
#include <list>
#include <string>
#include <vector>
namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point3f { float x, y, z; Point3f() : x(0), y(0), z(0) {} Point3f(float a, float b, float c) : x(a), y(b), z(c) {} };
class Mat { public: std::vector<double> d; int rows = 0, cols = 0; };
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) { rows = r; cols = c; d.assign((size_t)r * c, 0.0); }
    struct Init { Mat_ &m; int k; Init &operator,(T v) { m.d[k++] = v; return *this; } operator Mat() const { return m; } };
    Init operator<<(T v) { d[0] = v; return Init{*this, 1}; }
};
inline void Rodrigues(const Mat &, Mat &) { std::fprintf(stderr, "cv::Rodrigues: not available in the oracle shim\n"); std::abort(); }
template <typename A, typename B> inline bool solvePnP(const A &, const B &, const Mat &, const Mat &, Mat &, Mat &, int) {
    std::fprintf(stderr, "cv::solvePnP: not available in the oracle shim\n"); std::abort(); return false; }
}  // namespace cv
