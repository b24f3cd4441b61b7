// opencv2/imgproc/imgproc.hpp — COMPAT LAYER (see core/core.hpp). Implemented in ../cv_impl.cpp.
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
// 8UC1, INTER_LINEAR only (ORBextractor.cc:1184)  -> orc_resize_linear
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
// 8UC1, Size(7,7), sigma 2, BORDER_REFLECT_101 only (ORBextractor.cc:1130) -> orc_gaussian_blur7
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
// 8UC1, BORDER_REFLECT_101 [+ BORDER_ISOLATED] only (ORBextractor.cc:1186-1192) -> orc_border_reflect101 semantics
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
}  // namespace cv
