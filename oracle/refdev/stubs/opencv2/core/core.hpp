// oracle/refdev/stubs/opencv2/core/core.hpp -- stand-in: the reference headers Pyramid.h / parameters.h only NAME cv::Mat
// in declarations (std::vector<cv::Mat>&); Morph::cpu_optimize_level uses the handful of operations of mat_stub.h.
// OpenCV-C++ is not installed here.
#pragma once
#include "../mat_stub.h"
