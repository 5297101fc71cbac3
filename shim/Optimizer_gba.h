// Flat form of the BA graph shared by shim/Optimizer_gba.cc and its tests (see there).
#ifndef CORB_SHIM_OPTIMIZER_GBA_H
#define CORB_SHIM_OPTIMIZER_GBA_H
#include <stdint.h>

#include <vector>

#include "corb_b200.h"

namespace corb_shim {

void quat_from_pose(const cv::Mat& Tcw, double q_xyzw[4], double t[3]);  // Converter::toSE3Quat
cv::Mat pose_from_quat(const double q_xyzw[4], const double t[3]);       // Converter::toCvMat(SE3Quat)

struct FlatBA {
    std::vector<ORB_SLAM2::KeyFrame*> kf;   // dense pose index -> keyframe
    std::vector<ORB_SLAM2::MapPoint*> mp;   // dense point index -> map point
    std::vector<bool> not_included;         // vbNotIncludedMP, per entry of vpMP
    std::vector<double> pose_q, pose_t, pose_cam, point_xyz, edge_obs, edge_inv_sigma2;
    std::vector<uint8_t> pose_fixed, point_fixed;
    std::vector<int32_t> edge_pose, edge_point;
    void build(const std::vector<ORB_SLAM2::KeyFrame*>& vpKFs, const std::vector<ORB_SLAM2::MapPoint*>& vpMP);
    corb_ba_problem problem();
    void write_back(const unsigned long nLoopKF);
};

}  // namespace corb_shim
#endif
