// TEST INFRASTRUCTURE ONLY. cv::FAST for the reference build (oracle/_ref):
// the reference ships no OpenCV binaries, so its single cv::FAST call site
// (feature_detection.cpp:91-94) is served by the restated FAST-9/16 in
// oracle/svo_oracle.c, which tests pin against python cv2 4.13 golden vectors.
#include <opencv2/opencv.hpp>
#include "../svo_oracle.h"

namespace cv {
void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmax)
{
  const int cap = img.rows * img.cols;
  std::vector<int> xs(cap), ys(cap), sc(cap);
  // the oracle takes a dense image (stride == cols), as cv::Mat levels are here
  const int n = svo_oracle_fast(img.data, img.cols, img.rows, threshold, nonmax ? 1 : 0, cap, xs.data(), ys.data(), sc.data());
  kps.clear();
  kps.reserve(n);
  for (int i = 0; i < n; ++i) kps.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1.f, (float)sc[i]));
}
}  // namespace cv
