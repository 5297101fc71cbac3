/* STUB OpenCV — TEST INFRASTRUCTURE ONLY (see opencv2/core/core.hpp). imgproc: the three calls of
 * ORBextractor.cc (:1086,:1120,:1122,:1127) + undistortPoints (Frame.cc:426, only reached with distortion). */
#ifndef CORB_REFSTUB_OPENCV_IMGPROC_HPP
#define CORB_REFSTUB_OPENCV_IMGPROC_HPP
#include "opencv2/core/core.hpp"
namespace cv {
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs,
                     InputArray R = noArray(), InputArray P = noArray());
}
#endif
