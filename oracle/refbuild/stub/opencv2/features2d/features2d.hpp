/* STUB OpenCV — TEST INFRASTRUCTURE ONLY (see opencv2/core/core.hpp). features2d: cv::FAST (ORBextractor.cc:809,814)
 * and KeyPointsFilter::retainBest (only in the dead ComputeKeyPointsOld, :1006,1024). */
#ifndef CORB_REFSTUB_OPENCV_FEATURES2D_HPP
#define CORB_REFSTUB_OPENCV_FEATURES2D_HPP
#include "opencv2/core/core.hpp"
namespace cv {
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
class KeyPointsFilter {
public:
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};
}
#endif
