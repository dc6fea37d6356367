// oracle/refdev/stubs/opencv2/opencv.hpp -- stand-in (see core/core.hpp, mat_stub.h): OpenCV-C++ is not installed here.
#pragma once
#include "mat_stub.h"
