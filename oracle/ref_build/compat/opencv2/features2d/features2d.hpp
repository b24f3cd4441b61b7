// opencv2/features2d/features2d.hpp — COMPAT LAYER (see core/core.hpp). Implemented in ../cv_impl.cpp.
#pragma once
#include <opencv2/core/core.hpp>
#include <vector>
namespace cv {
// cv::FAST(img, kps, t, nonmaxSuppression) = FAST_t<16> (ORBextractor.cc:832,850) -> orc_fast
void FAST(InputArray image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true);
// Only referenced by the dead ComputeKeyPointsOld (ORBextractor.cc:902-1072); must link, is never called.
class KeyPointsFilter {
public:
    static void retainBest(std::vector<KeyPoint> &keypoints, int npoints);
};
}  // namespace cv
