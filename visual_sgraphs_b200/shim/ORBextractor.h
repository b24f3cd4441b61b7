// ORBextractor.h — drop-in replacement for the reference's orb_slam3/include/ORBextractor.h
// (snt-arg/visual_sgraphs).  Same namespace (VS_GRAPHS — the fork's name; ORB_SLAM3 is offered as an
// alias), same constructor, same operator()(image, mask, keypoints, descriptors, vLappingArea), same
// inline getters and the same public mvImagePyramid member, so Frame / Tracking / LocalMapping compile
// and behave unchanged.  All image work runs on the GPU through the C ABI in include/vsg_cuda.h; there
// is no CPU implementation behind this class.
//
// Differences a maintainer should know about:
//   * ExtractorNode (ORBextractor.h:29-40) is not declared: the oct-tree lives in a CUDA kernel.
//   * mvImagePyramid is filled after every call (levels downloaded, 19-px BORDER_REFLECT_101 frame
//     rebuilt on the host) because Frame::ComputeStereoMatches reads it (Frame.cc:964,1054,1069).
//     Call SetPyramidDownload(false) on extractors whose pyramid is never read (monocular / RGB-D).
#ifndef VSG_SHIM_ORBEXTRACTOR_H
#define VSG_SHIM_ORBEXTRACTOR_H

#include <vector>

#include "cv_compat.h"

struct vsg_extractor;

namespace VS_GRAPHS {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    // Returns the number of keypoints outside vLappingArea (monoIndex), or -1 for an empty image.
    int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint> &_keypoints,
                   cv::OutputArray _descriptors, std::vector<int> &vLappingArea);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;

    // --- additions ---
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }
    void SetDevice(int device);                 // before the first call; default 0
    vsg_extractor *Handle() { return mpHandle; }  // for callers that want the device-side pyramid (stereo)
    // Rectification on the device: map1 / map2 are this camera's CV_32FC1 maps of cv::initUndistortRectifyMap
    // (Settings.cc:571-574, width x height = the rectified size).  From then on operator() takes the UNRECTIFIED image
    // and does what System::TrackStereo's cv::remap(im, imToFeed, M1, M2, cv::INTER_LINEAR) (System.cc:284-292) plus the
    // reference's operator() would; mvImagePyramid[0] is the rectified image.
    void SetRectification(const float *map1, const float *map2, int width, int height);

protected:
    void EnsureHandle();

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    vsg_extractor *mpHandle = nullptr;
    int mnDevice = 0;
    bool mbDownloadPyramid = true;
    int mnRectWidth = 0, mnRectHeight = 0;   // > 0: SetRectification was called
    std::vector<cv::Mat> mvPyramidStorage;   // padded buffers the mvImagePyramid ROIs point into
};

}  // namespace VS_GRAPHS

namespace ORB_SLAM3 {
using VS_GRAPHS::ORBextractor;
}

#endif
