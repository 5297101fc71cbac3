// Flat form of the BA graph shared by shim/Optimizer_gba.cc and its tests (see there).
#ifndef CORB_SHIM_OPTIMIZER_GBA_H
#define CORB_SHIM_OPTIMIZER_GBA_H
#include <stdint.h>

#include <stdlib.h>

#include <new>
#include <vector>

#include "corb_b200.h"

namespace corb_shim {

// Allocator of the edge arrays: page-locked memory from the library (corb_host_alloc), so that corb_ba_solve uploads them with
// the DMA engines in place; plain malloc when the library cannot provide it (no GPU in the process: the CPU tests of the seam).
template <typename T>
struct PageLocked {
    typedef T value_type;
    PageLocked() {}
    template <typename U> PageLocked(const PageLocked<U>&) {}
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T) + 16;  // 16-byte header in front: who owns the block
        void* p = nullptr;
        const bool locked = corb_host_alloc(bytes, &p) == CORB_OK && p;
        if (!locked) p = malloc(bytes);
        if (!p) throw std::bad_alloc();
        *static_cast<int*>(p) = locked ? 1 : 0;
        return reinterpret_cast<T*>(static_cast<char*>(p) + 16);
    }
    void deallocate(T* q, size_t) {
        void* p = reinterpret_cast<char*>(q) - 16;
        if (*static_cast<int*>(p)) corb_host_free(p); else free(p);
    }
    template <typename U> bool operator==(const PageLocked<U>&) const { return true; }
    template <typename U> bool operator!=(const PageLocked<U>&) const { return false; }
};

void quat_from_pose(const cv::Mat& Tcw, double q_xyzw[4], double t[3]);  // Converter::toSE3Quat
cv::Mat pose_from_quat(const double q_xyzw[4], const double t[3]);       // Converter::toCvMat(SE3Quat)

struct FlatBA {
    std::vector<ORB_SLAM2::KeyFrame*> kf;   // dense pose index -> keyframe
    std::vector<ORB_SLAM2::MapPoint*> mp;   // dense point index -> map point
    std::vector<bool> not_included;         // vbNotIncludedMP, per entry of vpMP
    std::vector<double> pose_q, pose_t, pose_cam, point_xyz;
    std::vector<uint8_t> pose_fixed, point_fixed;
    std::vector<double, PageLocked<double> > edge_obs, edge_inv_sigma2;  // the bulk of the upload: 40 B per observation
    std::vector<int32_t, PageLocked<int32_t> > edge_pose, edge_point;
    void build(const std::vector<ORB_SLAM2::KeyFrame*>& vpKFs, const std::vector<ORB_SLAM2::MapPoint*>& vpMP);
    corb_ba_problem problem();
    void write_back(const unsigned long nLoopKF);
};

}  // namespace corb_shim
#endif
