// TEST INFRASTRUCTURE ONLY — not product code.
//
// Minimal stand-in for <opencv2/opencv.hpp> so that the reference's hot-path
// translation units (which use cv::Mat purely as an image container, plus one
// algorithm, cv::FAST) compile on this host without OpenCV binaries.  The
// reference ships OpenCV headers only (its libs are in .MISSING_LARGE_BLOBS).
//
// What is provided:
//   * cv::Mat with 64-byte aligned storage (so the reference's is_aligned16
//     dispatch in vision.cpp:78 behaves as with a real cv::Mat),
//     data/rows/cols/step.p[0]/type()/empty()/size()/clone()/at<T>()
//   * cv::FAST — declared here, defined in oracle/shim/cv_fast.cpp on top of
//     oracle/svo_oracle.c:svo_oracle_fast(), which is pinned against python
//     cv2 4.13 golden vectors (tests/golden/fast_*.npz).
//   * no-op / aborting stubs for calib3d + highgui calls that the hot path
//     never reaches with a distortion-free pinhole camera.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cassert>
#include <vector>
#include <memory>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <functional>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6
#define CV_16SC2 11
#define CV_32FC2 13
typedef unsigned char uchar;

namespace cv {

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
typedef Size Size2i;

struct Scalar {
  double v[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
};

template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<float> Point2f;

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0) : pt(x, y), size(s), angle(a), response(r), octave(0), class_id(-1) {}
};

struct MatStep {
  size_t p[2];
  MatStep() { p[0] = p[1] = 0; }
};

static inline size_t shim_elem_size(int t) {
  switch (t) {
    case CV_8U: return 1;
    case CV_32F: return 4;
    case CV_16SC2: return 4;
    case CV_64F: return 8;
    case CV_32FC2: return 8;
    default: return 8;
  }
}

struct Mat {
  int rows, cols, type_;
  uchar* data;
  MatStep step;
  std::shared_ptr<uchar> hold;

  Mat() : rows(0), cols(0), type_(0), data(nullptr) {}
  Mat(int r, int c, int t) : rows(r), cols(c), type_(t), data(nullptr) {
    size_t es = shim_elem_size(t);
    step.p[0] = (size_t)c * es;
    step.p[1] = es;
    void* m = nullptr;
    if (posix_memalign(&m, 64, step.p[0] * (size_t)r + 64)) abort();
    hold.reset((uchar*)m, free);
    data = (uchar*)m;
  }
  // header over external memory (no ownership)
  Mat(int r, int c, int t, void* d) : rows(r), cols(c), type_(t), data((uchar*)d) {
    step.p[0] = (size_t)c * shim_elem_size(t);
    step.p[1] = shim_elem_size(t);
  }
  Mat(Size s, int t, Scalar) : Mat(s.height, s.width, t) { memset(data, 0, step.p[0] * rows); }

  int type() const { return type_; }
  bool empty() const { return data == nullptr; }
  Size size() const { return Size(cols, rows); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    memcpy(m.data, data, step.p[0] * rows);
    return m;
  }
  template <class T> T& at(int y, int x) { return *(T*)(data + y * step.p[0] + x * sizeof(T)); }
  template <class T> const T& at(int y, int x) const { return *(const T*)(data + y * step.p[0] + x * sizeof(T)); }
  Mat operator*(double) const { return *this; }  // only used by display code (resimg_*10)
};

template <class T> struct Mat_ : Mat {
  Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_64F) {}
  static Mat_ eye(int r, int c) { return Mat_(r, c); }
  Mat_& operator<<(double) { return *this; }
  Mat_& operator,(double) { return *this; }
};

struct TermCriteria {
  enum { COUNT = 1, EPS = 2 };
  TermCriteria(int, int, double) {}
};
enum { RANSAC = 8, FM_RANSAC = 8, OPTFLOW_USE_INITIAL_FLOW = 4, INTER_LINEAR = 1, WINDOW_AUTOSIZE = 1, COLOR_RGBA2GRAY = 11 };

// features2d: FAST-9/16 (OpenCV's default type), defined in cv_fast.cpp
void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmax_suppression);

// highgui: display code paths are never enabled (display_=false)
inline void namedWindow(const char*, int) {}
inline void imshow(const char*, const Mat&) {}
inline int waitKey(int) { return 0; }

// calib3d / imgproc: only reachable with lens distortion, which the synthetic
// cameras do not have.  initUndistortRectifyMap is called unconditionally by
// the PinholeCamera ctor (pinhole_camera.cpp:34) => no-op.
inline void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) {}
inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&) { abort(); }
inline void remap(const Mat&, Mat&, const Mat&, const Mat&, int) { abort(); }

}  // namespace cv
