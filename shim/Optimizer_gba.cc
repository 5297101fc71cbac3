// Replacement body of Optimizer::BundleAdjustment (corbslam_client/src/Optimizer.cc:54-270): the same static member with the
// same signature; the g2o graph (new VertexSE3Expmap / VertexSBAPointXYZ / Edge*SE3ProjectXYZ per element, :80-207) becomes the
// flat corb_ba_problem arrays, optimizer.optimize() becomes corb_ba_solve on the GPU, and the results are written back exactly
// like :219-263. GlobalBundleAdjustemnt (:43-51) keeps calling it.
//
// Pieces of Eigen / g2o arithmetic that sit on this seam are restated (fp64, their published formulas):
//   Converter::toSE3Quat (Converter.cc:37-47) = g2o::SE3Quat(R, t): Eigen's Quaterniond(Matrix3d) + normalizeRotation (se3quat.h:58-60,280-288)
//   Converter::toCvMat(SE3Quat) (Converter.cc:49-53) = to_homogeneous_matrix: Eigen's Quaterniond::toRotationMatrix (se3quat.h:270-278)
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#ifndef CORB_SHIM_OPTIMIZER_DECLARED  // in the reference tree this is simply #include "Optimizer.h"
#include "Optimizer.h"
#endif
#include "shim_common.h"
#include "Optimizer_gba.h"

using namespace std;

namespace corb_shim {

void quat_from_pose(const cv::Mat& T, double q[4], double t[3]) {
    double m[3][3];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) m[i][j] = T.at<float>(i, j);
        t[i] = T.at<float>(i, 3);
    }
    double x, y, z, w;
    double tr = m[0][0] + m[1][1] + m[2][2];
    if (tr > 0) {
        double s = sqrt(tr + 1.0);
        w = 0.5 * s;
        s = 0.5 / s;
        x = (m[2][1] - m[1][2]) * s; y = (m[0][2] - m[2][0]) * s; z = (m[1][0] - m[0][1]) * s;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
        double v[3];
        v[i] = 0.5 * s;
        s = 0.5 / s;
        w = (m[k][j] - m[j][k]) * s;
        v[j] = (m[j][i] + m[i][j]) * s;
        v[k] = (m[k][i] + m[i][k]) * s;
        x = v[0]; y = v[1]; z = v[2];
    }
    if (w < 0) { x = -x; y = -y; z = -z; w = -w; }  // normalizeRotation
    const double n = sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}

cv::Mat pose_from_quat(const double q[4], const double t[3]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    const double R[3][3] = {{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}};
    cv::Mat T(4, 4, CV_32F);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T.at<float>(i, j) = (float)R[i][j];  // Converter::toCvMat(Matrix4d): double -> float
        T.at<float>(i, 3) = (float)t[i];
        T.at<float>(3, i) = 0.f;
    }
    T.at<float>(3, 3) = 1.f;
    return T;
}

// Optimizer.cc:80-207 without g2o: dense indices instead of the sparse vertex ids (mnId, mnId + maxKFid + 1)
void FlatBA::build(const vector<ORB_SLAM2::KeyFrame*>& vpKFs, const vector<ORB_SLAM2::MapPoint*>& vpMP) {
    using namespace ORB_SLAM2;
    long unsigned int maxKFid = 0;
    // keyframe vertices, ascending mnId like g2o's active vertex order (sparse_optimizer.cpp:166-190)
    vector<KeyFrame*> kfs;
    for (size_t i = 0; i < vpKFs.size(); i++) {
        KeyFrame* pKF = vpKFs[i];
        if (pKF && pKF->mnId > maxKFid) maxKFid = pKF->mnId;  // :85-86
        if (pKF->isBad()) continue;                            // :87-88
        kfs.push_back(pKF);
    }
    sort(kfs.begin(), kfs.end(), [](KeyFrame* a, KeyFrame* b) { return a->mnId < b->mnId; });
    map<long unsigned int, int> kfIndex;  // allKFId (:96) -> dense index
    for (size_t i = 0; i < kfs.size(); i++) {
        KeyFrame* pKF = kfs[i];
        if (kfIndex.count(pKF->mnId)) continue;  // a second vertex with the same id is refused by addVertex
        kfIndex[pKF->mnId] = (int)kf.size();
        kf.push_back(pKF);
        double q[4], t[3];
        quat_from_pose(pKF->GetPose(), q, t);  // :90
        pose_q.insert(pose_q.end(), q, q + 4);
        pose_t.insert(pose_t.end(), t, t + 3);
        pose_fixed.push_back(pKF->mnId == 1 || pKF->getFixed());  // :92
        const double cam[5] = {pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf};  // float members widened (:162-165, :186-190)
        pose_cam.insert(pose_cam.end(), cam, cam + 5);
    }
    not_included.assign(vpMP.size(), true);  // vbNotIncludedMP (:63-64, :106)
    {   // one page-locked block per edge array (PageLocked): an upper bound of the observations is reserved up front
        size_t n_obs = 0;
        for (size_t i = 0; i < vpMP.size(); i++)
            if (vpMP[i] && !vpMP[i]->isBad()) n_obs += (size_t)vpMP[i]->Observations();
        edge_pose.reserve(n_obs); edge_point.reserve(n_obs); edge_inv_sigma2.reserve(n_obs); edge_obs.reserve(3 * n_obs);
    }
    for (size_t i = 0; i < vpMP.size(); i++) {
        MapPoint* pMP = vpMP[i];
        if (!pMP || pMP->isBad()) continue;  // :107-109
        const size_t e0 = edge_pose.size();
        const map<KeyFrame*, size_t> observations = pMP->GetObservations();
        for (map<KeyFrame*, size_t>::const_iterator mit = observations.begin(); mit != observations.end(); mit++) {
            KeyFrame* pKF = mit->first;
            if (pKF->isBad() || pKF->mnId > maxKFid) continue;           // :130-131
            map<long unsigned int, int>::const_iterator it = kfIndex.find(pKF->mnId);
            if (it == kfIndex.end()) continue;                           // :133-134
            const cv::KeyPoint& kpUn = pKF->mvKeysUn[mit->second];
            const float ur = pKF->mvuRight[mit->second];
            edge_pose.push_back(it->second);
            edge_point.push_back((int32_t)mp.size());
            edge_obs.push_back(kpUn.pt.x);
            edge_obs.push_back(kpUn.pt.y);
            edge_obs.push_back(ur < 0 ? -1.0 : (double)ur);              // mono (:140-169) or stereo edge (:170-195)
            edge_inv_sigma2.push_back(pKF->mvInvLevelSigma2[kpUn.octave]);  // float widened (:151-152, :178-180)
        }
        if (edge_pose.size() == e0) continue;  // nEdges == 0: removeVertex, stays "not included" (:198-202)
        not_included[i] = false;
        const cv::Mat X = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) point_xyz.push_back(X.at<float>(k));  // Converter::toVector3d (:113)
        point_fixed.push_back(pMP->getFixed());                          // :118
        mp.push_back(pMP);
    }
}

corb_ba_problem FlatBA::problem() {
    corb_ba_problem p;
    p.n_poses = (int32_t)kf.size(); p.n_points = (int32_t)mp.size(); p.n_edges = (int32_t)edge_pose.size();
    p.pose_q = pose_q.data(); p.pose_t = pose_t.data(); p.pose_fixed = pose_fixed.data(); p.pose_cam = pose_cam.data();
    p.point_xyz = point_xyz.data(); p.point_fixed = point_fixed.data();
    p.edge_pose = edge_pose.data(); p.edge_point = edge_point.data(); p.edge_obs = edge_obs.data(); p.edge_inv_sigma2 = edge_inv_sigma2.data();
    return p;
}

// Optimizer.cc:216-263
void FlatBA::write_back(const unsigned long nLoopKF) {
    using namespace ORB_SLAM2;
    for (size_t i = 0; i < kf.size(); i++) {
        KeyFrame* pKF = kf[i];
        if (pKF->isBad() || pKF->getFixed()) continue;  // :222-223
        const cv::Mat T = pose_from_quat(&pose_q[4 * i], &pose_t[3 * i]);
        if (nLoopKF == 0) {
            pKF->SetPose(T);
            if (pKF->mpCacher) pKF->mpCacher->addUpdateKeyframe(pKF);
        } else {
            pKF->mTcwGBA.create(4, 4, CV_32F);
            T.copyTo(pKF->mTcwGBA);
            pKF->mnBAGlobalForKF = nLoopKF;
        }
    }
    for (size_t i = 0; i < mp.size(); i++) {
        MapPoint* pMP = mp[i];
        if (pMP->isBad() || pMP->getFixed()) continue;  // :245-246
        cv::Mat X(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) X.at<float>(k) = (float)point_xyz[3 * i + k];
        if (nLoopKF == 0) {
            pMP->SetWorldPos(X);
            if (pMP->getCache()) pMP->getCache()->addUpdateMapPoint(pMP);
            pMP->UpdateNormalAndDepth();
        } else {
            pMP->mPosGBA.create(3, 1, CV_32F);
            X.copyTo(pMP->mPosGBA);
            pMP->mnBAGlobalForKF = nLoopKF;
        }
    }
}

}  // namespace corb_shim

namespace ORB_SLAM2 {

void Optimizer::BundleAdjustment(const vector<KeyFrame*>& vpKFs, const vector<MapPoint*>& vpMP, int nIterations, bool* pbStopFlag,
                                 const unsigned long nLoopKF, const bool bRobust) {
    corb_shim::FlatBA flat;
    flat.build(vpKFs, vpMP);
    corb_ba_problem prob = flat.problem();
    corb_ba_result res;
    // single GPU here; a server with more than one client attached passes its ncclAllReduce hook and its landmark shard
    // (tests/host_harness/ba_nccl.cpp)
    const int rc = corb_ba_solve(&prob, nIterations, (const volatile uint8_t*)pbStopFlag, bRobust ? 1 : 0, corb_shim::device(), &res, 0, 0);
    if (rc != CORB_OK && rc != CORB_ERR_STOPPED) corb_shim::check(rc, "corb_ba_solve");
    flat.write_back(nLoopKF);
}

}  // namespace ORB_SLAM2
