// oracle/refdev/stubs/opencv2/mat_stub.h -- stand-in for the few cv::Mat operations Morph::cpu_optimize_level uses
// (morph.cu:433-437 zeros, 440-561 at<float>, 565-570 inv / countNonZero / operator*, 581-582 at<float>) and for the
// at<Vec2f>(y, x) element access of CQuadraticPath::optimize (QuadraticPath.cpp:38-64, 213) and CMatchingThread::Resize /
// BiLinear (MatchingThread.cpp:86-136) and of the temporal flow composition of Pyramid::build (pyramid.cu:406-441, 488-523).
// OpenCV-C++ is not installed here.  The ASSEMBLY of the dense system is the reference's text running on this class; the inverse itself
// (cv::Mat::inv of OpenCV, a third-party operation) is NOT reproduced: inv() only marks the matrix, and the product
// "A^-1 * B" RECORDS (A, B) for the test and returns B, so that the reference's load of X / Y into lvl.v (morph.cu:573-584)
// can be checked for its layout.  The oracle's own solve is deviation D4 (oracle/vmo.h).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <vector>
#ifndef CV_32FC1
#define CV_32FC1 5
#endif
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))          // OpenCV's definitions (core/cvdef.h)
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
namespace cv {
enum { DECOMP_LU = 0, DECOMP_SVD = 1 };
struct Vec2f {
    float val[2];
    Vec2f() { val[0] = val[1] = 0.0f; }
    Vec2f(float a, float b) { val[0] = a; val[1] = b; }
    float &operator[](int i) { return val[i]; }
    const float &operator[](int i) const { return val[i]; }
    Vec2f &operator+=(const Vec2f &b) { val[0] += b.val[0]; val[1] += b.val[1]; return *this; }
};
// cv::Vec arithmetic used by CMatchingThread::BiLinear (MatchingThread.cpp:133-134): element-wise in float
inline Vec2f operator*(const Vec2f &a, float s) { return Vec2f(a.val[0] * s, a.val[1] * s); }
inline Vec2f operator+(const Vec2f &a, const Vec2f &b) { return Vec2f(a.val[0] + b.val[0], a.val[1] + b.val[1]); }
class Mat {
public:
    int rows = 0, cols = 0;
    bool inverse = false;
    std::vector<float> d;                            // rows * cols elements of sizeof(T) / 4 floats each
    static Mat zeros(int r, int c, int) { Mat m; m.rows = r; m.cols = c; m.d.assign((size_t)r * c, 0.0f); return m; }
    static Mat zeros2(int r, int c) { Mat m; m.rows = r; m.cols = c; m.d.assign((size_t)r * c * 2, 0.0f); return m; }   // CV_32FC2
    template <class T> T &at(int i, int j) { return reinterpret_cast<T *>(d.data())[(size_t)i * cols + j]; }
    Mat inv(int = DECOMP_LU) const { Mat m = *this; m.inverse = true; return m; }
};
struct SolveCapture { std::vector<Mat> A, B; };
inline SolveCapture &solve_capture() { static SolveCapture c; return c; }
inline int countNonZero(const Mat &m) {
    if (m.inverse) return 1;
    int n = 0;
    for (float v : m.d) n += (v != 0.0f);
    return n;
}
inline Mat operator*(const Mat &a, const Mat &b) {
    if (a.inverse) { Mat A = a; A.inverse = false; solve_capture().A.push_back(A); solve_capture().B.push_back(b); return b; }
    Mat r = Mat::zeros(a.rows, b.cols, CV_32FC1);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) { float s = 0; for (int k = 0; k < a.cols; k++) s += a.d[(size_t)i * a.cols + k] * b.d[(size_t)k * b.cols + j]; r.d[(size_t)i * r.cols + j] = s; }
    return r;
}
}  // namespace cv
