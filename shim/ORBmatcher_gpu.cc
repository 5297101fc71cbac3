// Replacement bodies for five members of ORBmatcher (the class and every other member stay the reference's):
//   int SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)             corbslam_client/src/ORBmatcher.cc:162-291
//   int SearchByBoWInServer(KeyFrame*, KeyFrame*, vector<MapPoint*>&)  :294-423
//   int SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)          :657-790
//   int SearchByProjection(Frame&, const vector<MapPoint*>&, float)    :44-131
//   int SearchByProjection(Frame&, const Frame&, float, bool)          :1470-1614
// A maintainer deletes those five bodies from ORBmatcher.cc and adds this file to the target; the
// build test does exactly that with a temporary copy (oracle/refbuild/Makefile, target libshim.so).
// The shim flattens the pointer graph (DBoW2::FeatureVector -> CSR, MapPoint* -> liveness bytes, Frame::mGrid -> CSR), calls
// the C ABI and maps the returned indices back to MapPoint*.
#include <string.h>

#include <vector>

#include "ORBmatcher.h"
#include "shim_common.h"

using namespace std;

namespace ORB_SLAM2 {

namespace {

struct FlatSide {  // one side of a SearchByBoW call
    vector<uint32_t> nodes, idx;
    vector<int32_t> off;
    vector<uint8_t> valid;
    vector<float> ang;
    corb_bow_side side;
    FlatSide(const cv::Mat& desc, const DBoW2::FeatureVector& fv, const vector<MapPoint*>* mps, const vector<cv::KeyPoint>& keys) {
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            nodes.push_back(it->first);
            off.push_back((int32_t)idx.size());
            idx.insert(idx.end(), it->second.begin(), it->second.end());
        }
        off.push_back((int32_t)idx.size());
        if (mps) {
            valid.resize(mps->size());
            for (size_t i = 0; i < mps->size(); ++i) valid[i] = (*mps)[i] && !(*mps)[i]->isBad();  // :196-200, :691-695
        }
        ang.resize(keys.size());
        for (size_t i = 0; i < keys.size(); ++i) ang[i] = keys[i].angle;
        side.desc = desc.data;
        side.n = desc.rows;
        side.fv_nodes = nodes.data(); side.fv_off = off.data(); side.fv_idx = idx.data(); side.fv_n = (int32_t)nodes.size();
        side.valid = mps ? valid.data() : 0;
        side.angles = ang.data();
    }
};

struct FlatFrame {  // what SearchByProjection reads from the current Frame
    vector<float> x, y, ang;
    vector<int32_t> oct, goff, gidx;
    vector<uint8_t> taken;
    corb_frame_view v;
    explicit FlatFrame(const Frame& F) {
        const int N = F.N;
        x.resize(N); y.resize(N); ang.resize(N); oct.resize(N); taken.assign(N, 0);
        for (int i = 0; i < N; i++) {
            const cv::KeyPoint& k = F.mvKeysUn[i];
            x[i] = k.pt.x; y[i] = k.pt.y; oct[i] = k.octave; ang[i] = k.angle;
            MapPoint* t = F.mvpMapPoints[i].getMapPoint();
            taken[i] = t && t->Observations() > 0;  // :81-84, :1554-1557
        }
        goff.assign(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0);
        for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
            for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
                const vector<size_t>& c = F.mGrid[ix][iy];
                for (size_t e = 0; e < c.size(); e++) gidx.push_back((int32_t)c[e]);
                goff[ix * FRAME_GRID_ROWS + iy + 1] = (int32_t)gidx.size();
            }
        memset(&v, 0, sizeof(v));
        v.n = N; v.x = x.data(); v.y = y.data(); v.octave = oct.data(); v.angle = ang.data();
        v.desc = F.mDescriptors.data; v.u_right = F.mvuRight.data(); v.taken = taken.data();
        v.grid_off = goff.data(); v.grid_idx = gidx.data();
        v.min_x = Frame::mnMinX; v.min_y = Frame::mnMinY; v.max_x = Frame::mnMaxX; v.max_y = Frame::mnMaxY;
        v.grid_w_inv = Frame::mfGridElementWidthInv; v.grid_h_inv = Frame::mfGridElementHeightInv;
        v.scale_factors = F.mvScaleFactors.data(); v.n_levels = (int32_t)F.mvScaleFactors.size();
        v.fx = Frame::fx; v.fy = Frame::fy; v.cx = Frame::cx; v.cy = Frame::cy; v.mbf = F.mbf; v.mb = F.mb;
        if (!F.mTcw.empty())
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++) v.Tcw[4 * r + c] = F.mTcw.at<float>(r, c);
    }
};

}  // namespace

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) {
    const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    FlatSide A(pKF->mDescriptors, pKF->mFeatVec, &vpMapPointsKF, pKF->mvKeysUn), B(F.mDescriptors, F.mFeatVec, 0, F.mvKeys);
    vector<int32_t> match(max(F.N, 1));
    int32_t n = 0;
    corb_shim::check(corb_bow_match(corb_shim::matcher(), CORB_BOW_KF_FRAME, &A.side, &B.side, mfNNratio, mbCheckOrientation,
                                    match.data(), &n), "corb_bow_match");
    vpMapPointMatches = vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    for (int b = 0; b < F.N; ++b)
        if (match[b] >= 0) vpMapPointMatches[b] = vpMapPointsKF[match[b]];
    return n;
}

int ORBmatcher::SearchByBoWInServer(KeyFrame* pKF, KeyFrame* F, vector<MapPoint*>& vpMapPointMatches) {
    const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    FlatSide A(pKF->mDescriptors, pKF->mFeatVec, &vpMapPointsKF, pKF->mvKeysUn), B(F->mDescriptors, F->mFeatVec, 0, F->mvKeys);
    vector<int32_t> match(max(F->N, 1));
    int32_t n = 0;
    corb_shim::check(corb_bow_match(corb_shim::matcher(), CORB_BOW_KF_SERVER, &A.side, &B.side, mfNNratio, mbCheckOrientation,
                                    match.data(), &n), "corb_bow_match");
    vpMapPointMatches = vector<MapPoint*>(F->N, static_cast<MapPoint*>(NULL));
    for (int b = 0; b < F->N; ++b)
        if (match[b] >= 0) vpMapPointMatches[b] = vpMapPointsKF[match[b]];
    return n;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) {
    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    FlatSide A(pKF1->mDescriptors, pKF1->mFeatVec, &vpMapPoints1, pKF1->mvKeysUn), B(pKF2->mDescriptors, pKF2->mFeatVec, &vpMapPoints2, pKF2->mvKeysUn);
    vector<int32_t> match(max((int)vpMapPoints1.size(), 1));
    int32_t n = 0;
    corb_shim::check(corb_bow_match(corb_shim::matcher(), CORB_BOW_KF_KF, &A.side, &B.side, mfNNratio, mbCheckOrientation, match.data(),
                                    &n), "corb_bow_match");
    vpMatches12 = vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    for (size_t a = 0; a < vpMapPoints1.size(); ++a)
        if (match[a] >= 0) vpMatches12[a] = vpMapPoints2[match[a]];  // :737
    return n;
}

int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th) {
    FlatFrame ff(F);
    const int M = (int)vpMapPoints.size();
    vector<uint8_t> in_view(M), blocks(M), desc(32 * (size_t)max(M, 1));
    vector<float> proj(3 * (size_t)max(M, 1)), vcos(max(M, 1));
    vector<int32_t> level(max(M, 1));
    for (int i = 0; i < M; i++) {
        MapPoint* p = vpMapPoints[i];
        in_view[i] = p->mbTrackInView && !p->isBad();  // :54-58
        if (!in_view[i]) continue;
        blocks[i] = p->Observations() > 0;
        proj[3 * i] = p->mTrackProjX; proj[3 * i + 1] = p->mTrackProjY; proj[3 * i + 2] = p->mTrackProjXR;
        level[i] = p->mnTrackScaleLevel;
        vcos[i] = p->mTrackViewCos;
        const cv::Mat d = p->GetDescriptor();
        memcpy(&desc[32 * (size_t)i], d.data, 32);
    }
    vector<int32_t> match(max(F.N, 1));
    int32_t n = 0;
    corb_shim::check(corb_search_by_projection_map(corb_shim::matcher(), &ff.v, M, in_view.data(), blocks.data(), proj.data(), level.data(),
                                                   vcos.data(), desc.data(), th, mfNNratio, match.data(), &n),
                     "corb_search_by_projection_map");
    for (int idx = 0; idx < F.N; idx++)
        if (match[idx] >= 0) F.mvpMapPoints[idx] = LightMapPoint(vpMapPoints[match[idx]]);  // :124-125
    return n;
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    FlatFrame ff(CurrentFrame);
    const int M = LastFrame.N;
    vector<uint8_t> valid(max(M, 1), 0), blocks(max(M, 1), 0), desc(32 * (size_t)max(M, 1));
    vector<float> xyz(3 * (size_t)max(M, 1)), ang(max(M, 1));
    vector<int32_t> oct(max(M, 1));
    for (int i = 0; i < M; i++) {
        MapPoint* p = LastFrame.mvpMapPoints[i].getMapPoint();
        oct[i] = LastFrame.mvKeys[i].octave;
        ang[i] = LastFrame.mvKeysUn[i].angle;
        if (!p || LastFrame.mvbOutlier[i]) continue;  // :1496-1500
        valid[i] = 1;
        blocks[i] = p->Observations() > 0;
        const cv::Mat X = p->GetWorldPos();
        for (int k = 0; k < 3; k++) xyz[3 * i + k] = X.at<float>(k);
        const cv::Mat d = p->GetDescriptor();
        memcpy(&desc[32 * (size_t)i], d.data, 32);
    }
    float Tlw[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) Tlw[4 * r + c] = LastFrame.mTcw.at<float>(r, c);
    vector<int32_t> match(max(CurrentFrame.N, 1));
    int32_t n = 0;
    corb_shim::check(corb_search_by_projection_last(corb_shim::matcher(), &ff.v, M, valid.data(), blocks.data(), xyz.data(), desc.data(),
                                                    oct.data(), ang.data(), Tlw, th, bMono, mbCheckOrientation, match.data(), &n),
                     "corb_search_by_projection_last");
    // a current feature that was assigned and then dropped by the rotation check ends up NULL in the reference (:1606), also
    // when it held a MapPoint before the call; the library reports -1 for it, so previously held points of untouched
    // features are kept and everything the call assigned is written
    for (int i2 = 0; i2 < CurrentFrame.N; i2++)
        if (match[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = LightMapPoint(LastFrame.mvpMapPoints[match[i2]].getMapPoint());
    return n;
}

}  // namespace ORB_SLAM2
