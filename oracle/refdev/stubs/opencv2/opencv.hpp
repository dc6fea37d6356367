// oracle/refdev/stubs/opencv2/opencv.hpp -- stand-in: the reference headers Pyramid.h / parameters.h only NAME cv::Mat
// in declarations (std::vector<cv::Mat>&); OpenCV-C++ is not installed here and none of it is called by the device code.
#pragma once
namespace cv { class Mat; }
