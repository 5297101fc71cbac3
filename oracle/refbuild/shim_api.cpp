/* C ABI over Optimizer::BundleAdjustment of shim/Optimizer_gba.cc — TEST INFRASTRUCTURE ONLY (libshim.so). The KeyFrame /
 * MapPoint objects are the stand-ins of stub/ref_world.h filled from a flat problem; poses travel as float32 4x4 matrices
 * (what KeyFrame::GetPose returns), so the float32 <-> SE3Quat conversions of the seam are exercised. */
#include <stdint.h>
#include <string.h>

#include <memory>
#include <vector>

#include "Frame.h"
#include "Optimizer_gba.h"

using namespace ORB_SLAM2;

namespace {
struct World {
    std::vector<std::unique_ptr<KeyFrame> > kfs;
    std::vector<std::unique_ptr<MapPoint> > mps;
    std::vector<KeyFrame*> vpKFs;
    std::vector<MapPoint*> vpMP;
    Cache cache;
    corb_shim::FlatBA flat;
};
}  // namespace

extern "C" {

/* n_kf keyframes: mnId, Tcw (16 floats, row major), fixed / bad flags, intrinsics (fx, fy, cx, cy, mbf); their keypoints are
 * implied by the observations. n_mp map points: xyz (float), fixed / bad flags (a NULL entry: flag 2). Observations
 * (kf index, mp index, u, v, u_right, octave). inv_sigma2 = mvInvLevelSigma2 table. */
void* shim_world_create(int n_kf, const uint64_t* kf_id, const float* kf_Tcw, const uint8_t* kf_flags, const float* kf_cam, int n_mp,
                        const float* mp_xyz, const uint8_t* mp_flags, int n_obs, const int32_t* obs_kf, const int32_t* obs_mp,
                        const float* obs_uvr, const int32_t* obs_octave, const float* inv_sigma2, int n_levels) {
    World* w = new World;
    for (int i = 0; i < n_kf; i++) {
        w->kfs.emplace_back(new KeyFrame());
        KeyFrame* k = w->kfs.back().get();
        k->mnId = kf_id[i];
        k->Tcw = cv::Mat(4, 4, CV_32F);
        memcpy(k->Tcw.data, kf_Tcw + 16 * i, 64);
        k->fixed = kf_flags[i] & 1; k->bad = (kf_flags[i] & 2) != 0;
        k->fx = kf_cam[5 * i]; k->fy = kf_cam[5 * i + 1]; k->cx = kf_cam[5 * i + 2]; k->cy = kf_cam[5 * i + 3]; k->mbf = kf_cam[5 * i + 4];
        k->mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + n_levels);
        k->mpCacher = &w->cache;
        w->vpKFs.push_back(k);
    }
    for (int i = 0; i < n_mp; i++) {
        if (mp_flags[i] & 4) { w->vpMP.push_back(0); continue; }
        w->mps.emplace_back(new MapPoint());
        MapPoint* p = w->mps.back().get();
        p->mnId = i;
        p->mWorldPos = cv::Mat(3, 1, CV_32F);
        memcpy(p->mWorldPos.data, mp_xyz + 3 * i, 12);
        p->fixed = mp_flags[i] & 1; p->bad = (mp_flags[i] & 2) != 0;
        p->cache = &w->cache;
        w->vpMP.push_back(p);
    }
    for (int e = 0; e < n_obs; e++) {
        KeyFrame* k = w->vpKFs[obs_kf[e]];
        MapPoint* p = w->vpMP[obs_mp[e]];
        if (!p) continue;
        cv::KeyPoint kp;
        kp.pt.x = obs_uvr[3 * e]; kp.pt.y = obs_uvr[3 * e + 1]; kp.octave = obs_octave[e];
        k->mvKeysUn.push_back(kp);
        k->mvuRight.push_back(obs_uvr[3 * e + 2]);
        p->obs[k] = k->mvKeysUn.size() - 1;
        p->nObs++;
    }
    return w;
}
void shim_world_destroy(void* h) { delete (World*)h; }

/* the flatten alone (no GPU): sizes, then the arrays */
void shim_ba_flatten(void* h, int32_t* n_poses, int32_t* n_points, int32_t* n_edges) {
    World* w = (World*)h;
    w->flat = corb_shim::FlatBA();
    w->flat.build(w->vpKFs, w->vpMP);
    *n_poses = (int32_t)w->flat.kf.size(); *n_points = (int32_t)w->flat.mp.size(); *n_edges = (int32_t)w->flat.edge_pose.size();
}
void shim_ba_flat_get(void* h, double* pose_q, double* pose_t, uint8_t* pose_fixed, double* pose_cam, double* point_xyz, uint8_t* point_fixed,
                      int32_t* edge_pose, int32_t* edge_point, double* edge_obs, double* edge_inv_sigma2, uint64_t* pose_kf_id,
                      int32_t* point_mp_index) {
    corb_shim::FlatBA& f = ((World*)h)->flat;
#define CP(dst, v) if (!(v).empty()) memcpy(dst, (v).data(), (v).size() * sizeof((v)[0]))
    CP(pose_q, f.pose_q); CP(pose_t, f.pose_t); CP(pose_fixed, f.pose_fixed); CP(pose_cam, f.pose_cam); CP(point_xyz, f.point_xyz);
    CP(point_fixed, f.point_fixed); CP(edge_pose, f.edge_pose); CP(edge_point, f.edge_point); CP(edge_obs, f.edge_obs);
    CP(edge_inv_sigma2, f.edge_inv_sigma2);
#undef CP
    for (size_t i = 0; i < f.kf.size(); i++) pose_kf_id[i] = f.kf[i]->mnId;
    for (size_t i = 0; i < f.mp.size(); i++) point_mp_index[i] = (int32_t)f.mp[i]->mnId;
}
/* results computed elsewhere (the oracle, in the CPU test) are put back and written into the objects like Optimizer.cc:216-263 */
void shim_ba_flat_set_and_write_back(void* h, const double* pose_q, const double* pose_t, const double* point_xyz, uint64_t nLoopKF) {
    corb_shim::FlatBA& f = ((World*)h)->flat;
    memcpy(f.pose_q.data(), pose_q, f.pose_q.size() * 8);
    memcpy(f.pose_t.data(), pose_t, f.pose_t.size() * 8);
    memcpy(f.point_xyz.data(), point_xyz, f.point_xyz.size() * 8);
    f.write_back(nLoopKF);
}
/* the whole seam: Optimizer::BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust) on the GPU */
void shim_ba_run(void* h, int nIterations, uint64_t nLoopKF, int bRobust) {
    World* w = (World*)h;
    Optimizer::BundleAdjustment(w->vpKFs, w->vpMP, nIterations, 0, nLoopKF, bRobust != 0);
}
/* read the objects back: per keyframe the current pose and mTcwGBA (16 floats each, zeros when empty), mnBAGlobalForKF; per
 * map point the position, mPosGBA, mnBAGlobalForKF, UpdateNormalAndDepth calls; and the cache queues' lengths */
void shim_world_read(void* h, float* kf_Tcw, float* kf_TcwGBA, uint64_t* kf_gba, float* mp_xyz, float* mp_posGBA, uint64_t* mp_gba,
                     int32_t* mp_normal_updates, int32_t* cache_counts) {
    World* w = (World*)h;
    for (size_t i = 0; i < w->vpKFs.size(); i++) {
        KeyFrame* k = w->vpKFs[i];
        memcpy(kf_Tcw + 16 * i, k->Tcw.data, 64);
        if (!k->mTcwGBA.empty()) memcpy(kf_TcwGBA + 16 * i, k->mTcwGBA.data, 64); else memset(kf_TcwGBA + 16 * i, 0, 64);
        kf_gba[i] = k->mnBAGlobalForKF;
    }
    for (size_t i = 0; i < w->vpMP.size(); i++) {
        MapPoint* p = w->vpMP[i];
        memset(mp_xyz + 3 * i, 0, 12); memset(mp_posGBA + 3 * i, 0, 12); mp_gba[i] = 0; mp_normal_updates[i] = 0;
        if (!p) continue;
        memcpy(mp_xyz + 3 * i, p->mWorldPos.data, 12);
        if (!p->mPosGBA.empty()) memcpy(mp_posGBA + 3 * i, p->mPosGBA.data, 12);
        mp_gba[i] = p->mnBAGlobalForKF;
        mp_normal_updates[i] = p->nNormalUpdates;
    }
    cache_counts[0] = (int32_t)w->cache.updatedKFs.size();
    cache_counts[1] = (int32_t)w->cache.updatedMPs.size();
}
void shim_quat_from_pose(const float* Tcw16, double* q, double* t) {
    cv::Mat T(4, 4, CV_32F, (void*)Tcw16);
    corb_shim::quat_from_pose(T, q, t);
}
void shim_pose_from_quat(const double* q, const double* t, float* Tcw16) {
    cv::Mat T = corb_shim::pose_from_quat(q, t);
    memcpy(Tcw16, T.data, 64);
}

} // extern "C"
