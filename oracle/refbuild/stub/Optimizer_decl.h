/* STAND-IN for corbslam_client/include/Optimizer.h (:39-46) — TEST INFRASTRUCTURE ONLY. The real header includes g2o's Sim3
 * types (Eigen, absent here); shim/Optimizer_gba.cc only needs the declaration of the one member it defines. */
#ifndef CORB_REF_OPTIMIZER_DECL_H
#define CORB_REF_OPTIMIZER_DECL_H
#ifdef __cplusplus
#include <vector>
namespace ORB_SLAM2 {
class Optimizer {
public:
    void static BundleAdjustment(const std::vector<KeyFrame*>& vpKF, const std::vector<MapPoint*>& vpMP, int nIterations = 5,
                                 bool* pbStopFlag = NULL, const unsigned long nLoopKF = 0, const bool bRobust = true);
};
}  // namespace ORB_SLAM2
#endif
#endif
