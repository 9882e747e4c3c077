// oracle/shim: cv <-> Eigen converters named by feature_manager.cpp's PnP initialisation (never executed by the oracle)
#pragma once
#include "../opencv.hpp"
namespace cv {
template <typename E> inline void eigen2cv(const E &, Mat &) { std::abort(); }
template <typename E> inline void cv2eigen(const Mat &, E &) { std::abort(); }
}  // namespace cv
