/* C ABI over the REFERENCE'S OWN CLASSES — TEST INFRASTRUCTURE ONLY (oracle/refbuild).
 *
 * Every function below marshals flat arrays into the reference's objects and calls the unmodified reference code:
 *   ORBextractor::operator()                       corbslam_client/src/ORBextractor.cc:1043-1105
 *   Frame::Frame (stereo) -> ExtractORB x2, ComputeStereoMatches, AssignFeaturesToGrid   src/Frame.cc:60-124,230-253,470-644
 *   Frame::GetFeaturesInArea                       src/Frame.cc:331-384
 *   ORBmatcher::SearchByBoW x2 / SearchByBoWInServer / SearchByProjection x2 / DescriptorDistance
 *                                                  src/ORBmatcher.cc:44-131,162-423,657-790,1470-1614,1792-1808
 *   ORBVocabulary::loadFromTextFile / transform / score   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, ScoringObject.cpp
 *   PnPsolver::SetRansacParameters / iterate       src/PnPsolver.cc:163-300
 * The argument layouts mirror oracle/*_oracle.cpp so that tests hand the same arrays to both. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "Frame.h"
#include "ORBVocabulary.h"
#include "ORBextractor.h"
#include "ORBmatcher.h"
#define private public /* PnPsolver::mnIterations / mRansacMinInliers / mRansacMaxIts are read back for comparison */
#include "PnPsolver.h"
#undef private
#include "Thirdparty/DBoW2/DUtils/Random.h"

using namespace ORB_SLAM2;

extern "C" void ref_alloc_scope(int delta);

namespace {

struct ExtractorScope { /* see ref_alloc.cpp */
    ExtractorScope() { ref_alloc_scope(1); }
    ~ExtractorScope() { ref_alloc_scope(-1); }
};

struct ref_keypoint { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(cv::KeyPoint) == sizeof(ref_keypoint), "cv::KeyPoint layout");

/* FeatureVector from CSR (node ids ascending, feature indices in addFeature order) */
void fill_featvec(DBoW2::FeatureVector& fv, const uint32_t* nodes, const int32_t* off, const uint32_t* idx, int n_nodes) {
    for (int g = 0; g < n_nodes; g++)
        for (int e = off[g]; e < off[g + 1]; e++) fv.addFeature(nodes[g], idx[e]);
}

/* MapPoints of one side: valid == 0 -> no MapPoint (NULL on even indices, a bad one on odd indices), else a live one */
struct Points {
    std::vector<std::unique_ptr<MapPoint> > store;
    std::vector<MapPoint*> ptr;
    void init(int n, const uint8_t* valid) {
        ptr.assign(n, 0);
        for (int i = 0; i < n; i++) {
            const bool live = !valid || valid[i];
            if (!live && (i & 1) == 0) continue;
            store.emplace_back(new MapPoint());
            MapPoint* p = store.back().get();
            p->mnId = i;
            p->bad = !live;
            p->nObs = 1;
            ptr[i] = p;
        }
    }
};

void fill_keyframe(KeyFrame& kf, const uint8_t* desc, int n, const uint32_t* nodes, const int32_t* off, const uint32_t* idx, int n_nodes,
                   const float* angles, Points& pts) {
    kf.N = n;
    kf.mDescriptors = cv::Mat(n, 32, CV_8U, (void*)desc);
    fill_featvec(kf.mFeatVec, nodes, off, idx, n_nodes);
    kf.mvKeysUn.resize(n);
    kf.mvKeys.resize(n);
    for (int i = 0; i < n; i++) kf.mvKeysUn[i].angle = kf.mvKeys[i].angle = angles ? angles[i] : 0.f;
    kf.mvpMapPoints.resize(n);
    for (int i = 0; i < n; i++) kf.mvpMapPoints[i] = LightMapPoint(pts.ptr[i]);
}

struct ref_frame_view { /* layout of corb_frame_view / oracle_frame_view */
    int32_t n;
    const float* x; const float* y; const int32_t* octave; const float* angle;
    const uint8_t* desc;
    const float* u_right;
    const uint8_t* taken;
    const int32_t* grid_off;
    const int32_t* grid_idx;
    float min_x, min_y, max_x, max_y, grid_w_inv, grid_h_inv;
    const float* scale_factors; int32_t n_levels;
    float fx, fy, cx, cy, mbf, mb;
    float Tcw[12];
};

const long unsigned int kPlaceholderId = 1ul << 40;

/* a reference Frame from the flattened view (public members only; the grid goes into the real mGrid) */
void fill_frame(Frame& F, const ref_frame_view* v, std::vector<std::unique_ptr<MapPoint> >& store) {
    F.N = v->n;
    F.mvKeysUn.resize(v->n);
    for (int i = 0; i < v->n; i++) {
        cv::KeyPoint& kp = F.mvKeysUn[i];
        kp.pt.x = v->x[i]; kp.pt.y = v->y[i]; kp.octave = v->octave[i]; kp.angle = v->angle[i];
    }
    F.mvKeys = F.mvKeysUn;
    F.mvuRight.assign(v->u_right, v->u_right + v->n);
    F.mDescriptors = cv::Mat(v->n, 32, CV_8U, (void*)v->desc);
    F.mvScaleFactors.assign(v->scale_factors, v->scale_factors + v->n_levels);
    F.mnScaleLevels = v->n_levels;
    F.mvpMapPoints.assign(v->n, LightMapPoint());
    for (int i = 0; i < v->n; i++)
        if (v->taken && v->taken[i]) { /* the feature already holds a MapPoint with observations */
            store.emplace_back(new MapPoint());
            store.back()->mnId = kPlaceholderId + i;
            store.back()->nObs = 1;
            F.mvpMapPoints[i] = LightMapPoint(store.back().get());
        }
    F.mvbOutlier.assign(v->n, false);
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
        for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            const int c = ix * FRAME_GRID_ROWS + iy;
            F.mGrid[ix][iy].assign(v->grid_idx + v->grid_off[c], v->grid_idx + v->grid_off[c + 1]);
        }
    F.mTcw = cv::Mat::eye(4, 4, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) F.mTcw.at<float>(r, c) = v->Tcw[4 * r + c];
    F.mbf = v->mbf; F.mb = v->mb;
    Frame::fx = v->fx; Frame::fy = v->fy; Frame::cx = v->cx; Frame::cy = v->cy;
    Frame::invfx = 1.0f / v->fx; Frame::invfy = 1.0f / v->fy;
    Frame::mnMinX = v->min_x; Frame::mnMinY = v->min_y; Frame::mnMaxX = v->max_x; Frame::mnMaxY = v->max_y;
    Frame::mfGridElementWidthInv = v->grid_w_inv; Frame::mfGridElementHeightInv = v->grid_h_inv;
}

} // namespace

extern "C" {

/* ------------------------------------------------------------------ ORBextractor */
void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    return new ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_orb_destroy(void* h) { delete (ORBextractor*)h; }
int ref_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, ref_keypoint* kps, uint8_t* desc, int cap) {
    ExtractorScope scope;
    ORBextractor& ex = *(ORBextractor*)h;
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    ex(image, cv::Mat(), keys, descriptors);
    const int n = (int)keys.size();
    if (n > cap) return -n;
    if (n) {
        memcpy(kps, keys.data(), sizeof(ref_keypoint) * n);
        for (int i = 0; i < n; i++) memcpy(desc + 32 * i, descriptors.ptr(i), 32);
    }
    return n;
}
int ref_orb_levels(void* h) { return ((ORBextractor*)h)->GetLevels(); }
void ref_orb_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
    ORBextractor& ex = *(ORBextractor*)h;
    const int n = ex.GetLevels();
    std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}
/* mvImagePyramid[level] (ROI inside its bordered buffer): pointer to pixel (0,0), size and row stride */
const uint8_t* ref_orb_pyramid(void* h, int level, int* w, int* hgt, int* stride) {
    const cv::Mat& m = ((ORBextractor*)h)->mvImagePyramid[level];
    *w = m.cols; *hgt = m.rows; *stride = (int)m.step;
    return m.data;
}

/* ------------------------------------------------------------------ Frame (stereo constructor)
 * Frame(imLeft, imRight, ..., extractorLeft, extractorRight, voc, K, distCoef, bf, thDepth): two extractor threads,
 * UndistortKeyPoints (no distortion), ComputeStereoMatches, image bounds, AssignFeaturesToGrid. */
void* ref_frame_stereo(void* exL, void* exR, const uint8_t* left, const uint8_t* right, int w, int h, int stride, float fx, float fy,
                       float cx, float cy, float bf, float thDepth) {
    ExtractorScope scope;
    cv::Mat L(h, w, CV_8UC1, (void*)left, (size_t)stride), R(h, w, CV_8UC1, (void*)right, (size_t)stride);
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
    cv::Mat dist = cv::Mat::zeros(4, 1, CV_32F);
    Frame::mbInitialComputations = true; /* image bounds and grid pitch of THIS image size */
    /* Frame.cc:87 runs ComputeStereoMatches() before `mb = mbf/fx` is assigned at :114, and no constructor initialises
     * mb: the reference reads an indeterminate member there (maxD = mbf/mb decides the disparity range). The storage is
     * therefore pre-loaded with the value every later reader sees, which is what the oracle and the CUDA path define. */
    static const size_t mb_offset = [] { Frame d; return (size_t)((char*)&d.mb - (char*)&d); }();
    void* mem = ::operator new(sizeof(Frame));
    memset(mem, 0, sizeof(Frame));
    const float mb = bf / fx;
    memcpy((char*)mem + mb_offset, &mb, sizeof(float));
    return new (mem) Frame(L, R, 0.0, (ORBextractor*)exL, (ORBextractor*)exR, (ORBVocabulary*)0, K, dist, bf, thDepth);
}
void ref_frame_destroy(void* f) { delete (Frame*)f; }
void ref_frame_counts(void* f, int* n_left, int* n_right) { *n_left = (int)((Frame*)f)->mvKeys.size(); *n_right = (int)((Frame*)f)->mvKeysRight.size(); }
void ref_frame_results(void* f, ref_keypoint* kl, uint8_t* dl, ref_keypoint* kr, uint8_t* dr, float* u_right, float* depth) {
    Frame& F = *(Frame*)f;
    const int nl = (int)F.mvKeys.size(), nr = (int)F.mvKeysRight.size();
    if (nl) memcpy(kl, F.mvKeys.data(), sizeof(ref_keypoint) * nl);
    if (nr) memcpy(kr, F.mvKeysRight.data(), sizeof(ref_keypoint) * nr);
    for (int i = 0; i < nl; i++) memcpy(dl + 32 * i, F.mDescriptors.ptr(i), 32);
    for (int i = 0; i < nr; i++) memcpy(dr + 32 * i, F.mDescriptorsRight.ptr(i), 32);
    for (int i = 0; i < nl; i++) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}
/* mGrid as CSR in the C-ABI's cell order (cell = ix * 48 + iy): off[64*48+1], idx[N]; plus bounds and grid pitch */
void ref_frame_grid(void* f, int32_t* off, int32_t* idx, float* bounds6) {
    Frame& F = *(Frame*)f;
    int k = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
        for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            off[ix * FRAME_GRID_ROWS + iy] = k;
            for (size_t e = 0; e < F.mGrid[ix][iy].size(); e++) idx[k++] = (int32_t)F.mGrid[ix][iy][e];
        }
    off[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
    bounds6[0] = Frame::mnMinX; bounds6[1] = Frame::mnMinY; bounds6[2] = Frame::mnMaxX; bounds6[3] = Frame::mnMaxY;
    bounds6[4] = Frame::mfGridElementWidthInv; bounds6[5] = Frame::mfGridElementHeightInv;
}
/* Frame::ComputeStereoMatches alone on given keypoints (the pyramids are those of the extractors' last call) */
void ref_stereo_matches(void* exL, void* exR, const ref_keypoint* kl, const uint8_t* dl, int nl, const ref_keypoint* kr, const uint8_t* dr,
                        int nr, float mbf, float mb, float* u_right, float* depth) {
    Frame F;
    ORBextractor* L = (ORBextractor*)exL;
    F.mpORBextractorLeft = L; F.mpORBextractorRight = (ORBextractor*)exR;
    F.N = nl;
    F.mvKeys.resize(nl); F.mvKeysRight.resize(nr);
    if (nl) memcpy(F.mvKeys.data(), kl, sizeof(ref_keypoint) * nl);
    if (nr) memcpy(F.mvKeysRight.data(), kr, sizeof(ref_keypoint) * nr);
    F.mDescriptors = cv::Mat(nl, 32, CV_8U, (void*)dl);
    F.mDescriptorsRight = cv::Mat(nr, 32, CV_8U, (void*)dr);
    F.mvScaleFactors = L->GetScaleFactors(); F.mvInvScaleFactors = L->GetInverseScaleFactors();
    F.mbf = mbf; F.mb = mb;
    F.ComputeStereoMatches();
    for (int i = 0; i < nl; i++) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

/* ------------------------------------------------------------------ ORBmatcher */
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, (void*)a), cv::Mat(1, 32, CV_8U, (void*)b));
}

/* variant 0: SearchByBoW(KeyFrame*, Frame&, ...)   1: SearchByBoWInServer(KeyFrame*, KeyFrame*, ...)
 * variant 2: SearchByBoW(KeyFrame*, KeyFrame*, ...)  -  same arguments and `match` meaning as oracle_search_by_bow */
int ref_search_by_bow(int variant, const uint8_t* descA, int nA, const uint8_t* descB, int nB, const uint32_t* fvA_nodes,
                      const int32_t* fvA_off, const uint32_t* fvA_idx, int n_fvA, const uint32_t* fvB_nodes, const int32_t* fvB_off,
                      const uint32_t* fvB_idx, int n_fvB, const uint8_t* validA, const uint8_t* validB, const float* anglesA,
                      const float* anglesB, float nnratio, int check_ori, int32_t* match) {
    ORBmatcher matcher(nnratio, check_ori != 0);
    Points ptsA, ptsB;
    ptsA.init(nA, validA);
    KeyFrame kfA;
    fill_keyframe(kfA, descA, nA, fvA_nodes, fvA_off, fvA_idx, n_fvA, anglesA, ptsA);
    std::vector<MapPoint*> out;
    int nm;
    if (variant == 0) {
        Frame F;
        F.N = nB;
        F.mDescriptors = cv::Mat(nB, 32, CV_8U, (void*)descB);
        fill_featvec(F.mFeatVec, fvB_nodes, fvB_off, fvB_idx, n_fvB);
        F.mvKeys.resize(nB);
        for (int i = 0; i < nB; i++) F.mvKeys[i].angle = anglesB ? anglesB[i] : 0.f;
        F.mvKeysUn = F.mvKeys;
        nm = matcher.SearchByBoW(&kfA, F, out);
    } else {
        ptsB.init(nB, variant == 2 ? validB : 0);
        KeyFrame kfB;
        fill_keyframe(kfB, descB, nB, fvB_nodes, fvB_off, fvB_idx, n_fvB, anglesB, ptsB);
        nm = variant == 1 ? matcher.SearchByBoWInServer(&kfA, &kfB, out) : matcher.SearchByBoW(&kfA, &kfB, out);
    }
    for (size_t i = 0; i < out.size(); i++) match[i] = out[i] ? (int32_t)out[i]->mnId : -1;
    return nm;
}

/* SearchByProjection(Frame& Current, const Frame& Last, th, bMono): arguments of oracle_search_by_projection_last */
int ref_search_by_projection_last(const ref_frame_view* cur, int n_last, const uint8_t* last_valid, const uint8_t* last_blocks,
                                  const float* last_xyz, const uint8_t* last_mp_desc, const int32_t* last_octave, const float* last_angle,
                                  const float* Tlw, float th, int mono, int check_ori, int32_t* match) {
    std::vector<std::unique_ptr<MapPoint> > store;
    Frame Cur, Last;
    fill_frame(Cur, cur, store);
    Last.N = n_last;
    Last.mvpMapPoints.assign(n_last, LightMapPoint());
    Last.mvbOutlier.assign(n_last, false);
    Last.mvKeys.resize(n_last);
    for (int i = 0; i < n_last; i++) {
        Last.mvKeys[i].octave = last_octave[i];
        Last.mvKeys[i].angle = last_angle[i];
        if (!last_valid[i]) {
            if ((i & 1) == 0) continue; /* no MapPoint; odd indices: a MapPoint flagged as outlier */
            Last.mvbOutlier[i] = true;
        }
        store.emplace_back(new MapPoint());
        MapPoint* p = store.back().get();
        p->mnId = i;
        p->nObs = (!last_blocks || last_blocks[i]) ? 1 : 0;
        p->mWorldPos = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) p->mWorldPos.at<float>(k) = last_xyz[3 * i + k];
        p->mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(last_mp_desc + 32 * (size_t)i)).clone();
        Last.mvpMapPoints[i] = LightMapPoint(p);
    }
    Last.mvKeysUn = Last.mvKeys;
    Last.mTcw = cv::Mat::eye(4, 4, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) Last.mTcw.at<float>(r, c) = Tlw[4 * r + c];
    ORBmatcher matcher(0.9f, check_ori != 0);
    const int nm = matcher.SearchByProjection(Cur, Last, th, mono != 0);
    for (int i = 0; i < cur->n; i++) {
        MapPoint* p = Cur.mvpMapPoints[i].getMapPoint();
        match[i] = (p && p->mnId < kPlaceholderId) ? (int32_t)p->mnId : -1;
    }
    return nm;
}

/* SearchByProjection(Frame& F, const vector<MapPoint*>&, th): arguments of oracle_search_by_projection_map */
int ref_search_by_projection_map(const ref_frame_view* Fv, int n_mp, const uint8_t* in_view, const uint8_t* blocks, const float* proj,
                                 const int32_t* level, const float* view_cos, const uint8_t* mp_desc, float th, float nnratio,
                                 int32_t* match) {
    std::vector<std::unique_ptr<MapPoint> > store;
    Frame F;
    fill_frame(F, Fv, store);
    std::vector<MapPoint*> mps(n_mp);
    for (int i = 0; i < n_mp; i++) {
        store.emplace_back(new MapPoint());
        MapPoint* p = store.back().get();
        p->mnId = i;
        p->mbTrackInView = in_view[i] != 0;
        p->nObs = (!blocks || blocks[i]) ? 1 : 0;
        p->mTrackProjX = proj[3 * i]; p->mTrackProjY = proj[3 * i + 1]; p->mTrackProjXR = proj[3 * i + 2];
        p->mnTrackScaleLevel = level[i];
        p->mTrackViewCos = view_cos[i];
        p->mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)i)).clone();
        mps[i] = p;
    }
    ORBmatcher matcher(nnratio, true);
    const int nm = matcher.SearchByProjection(F, mps, th);
    for (int i = 0; i < Fv->n; i++) {
        MapPoint* p = F.mvpMapPoints[i].getMapPoint();
        match[i] = (p && p->mnId < kPlaceholderId) ? (int32_t)p->mnId : -1;
    }
    return nm;
}

/* Frame::GetFeaturesInArea on a view (returns the count; out sized >= n) */
int ref_features_in_area(const ref_frame_view* Fv, float x, float y, float r, int minLevel, int maxLevel, int32_t* out) {
    std::vector<std::unique_ptr<MapPoint> > store;
    Frame F;
    fill_frame(F, Fv, store);
    std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size(); i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}

/* ------------------------------------------------------------------ ORBVocabulary (DBoW2) */
void* ref_voc_load_text(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return 0; }
    return v;
}
void ref_voc_destroy(void* v) { delete (ORBVocabulary*)v; }
void ref_voc_info(void* v, int* k, int* L, int* n_words) {
    ORBVocabulary& voc = *(ORBVocabulary*)v;
    *k = voc.getBranchingFactor(); *L = voc.getDepthLevels(); *n_words = (int)voc.size();
}
/* transform(features, BowVector&, FeatureVector&, levelsup) exactly as Frame::ComputeBoW calls it (Frame.cc:399-406).
 * Outputs: the two std::maps flattened in key order. Returns #bow entries; *n_fv = #feature-vector nodes. */
int ref_voc_transform(void* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_words, double* bow_vals, uint32_t* fv_nodes,
                      int32_t* fv_off, uint32_t* fv_idx, int* n_fv) {
    ORBVocabulary& voc = *(ORBVocabulary*)v;
    cv::Mat D(n, 32, CV_8U, (void*)desc);
    std::vector<cv::Mat> feats = Converter::toDescriptorVector(D);
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    voc.transform(feats, bv, fv, levelsup);
    int m = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++m) { bow_words[m] = it->first; bow_vals[m] = it->second; }
    int g = 0, e = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++g) {
        fv_nodes[g] = it->first;
        fv_off[g] = e;
        for (size_t j = 0; j < it->second.size(); j++) fv_idx[e++] = it->second[j];
    }
    fv_off[g] = e;
    *n_fv = g;
    return m;
}
double ref_voc_score(void* v, int n1, const uint32_t* w1, const double* v1, int n2, const uint32_t* w2, const double* v2) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; i++) a.insert(a.end(), std::make_pair(w1[i], v1[i]));
    for (int i = 0; i < n2; i++) b.insert(b.end(), std::make_pair(w2[i], v2[i]));
    return ((ORBVocabulary*)v)->score(a, b);
}

/* ------------------------------------------------------------------ PnPsolver
 * The reference consumes the process-global rand() stream (DUtils::Random::RandomInt). ref_rand_draws(seed, ...) returns
 * the values RandomInt will return for the first `iters` RANSAC iterations after ref_srand(seed): pick k of an
 * iteration draws from N - k available indices (PnPsolver.cc:236-247). */
void ref_srand(unsigned seed) { srand(seed); }
void ref_rand_draws(unsigned seed, int N, int iters, int32_t* draws) {
    srand(seed);
    for (int it = 0; it < iters; it++)
        for (int k = 0; k < 4; k++) draws[4 * it + k] = DUtils::Random::RandomInt(0, N - k - 1);
}
struct RefPnp {
    Frame F;
    std::vector<std::unique_ptr<MapPoint> > store;
    std::vector<MapPoint*> mps;
    std::unique_ptr<PnPsolver> solver;
};
/* N matches: keypoint (x, y, octave), MapPoint position; sigma2 = mvLevelSigma2 table */
void* ref_pnp_create(int N, const float* p2d, const int32_t* octave, const float* p3d, const float* level_sigma2, int n_levels, float fx,
                     float fy, float cx, float cy) {
    RefPnp* r = new RefPnp;
    r->F.N = N;
    r->F.mvKeysUn.resize(N);
    r->F.mvLevelSigma2.assign(level_sigma2, level_sigma2 + n_levels);
    r->F.mvpMapPoints.assign(N, LightMapPoint());
    r->mps.resize(N);
    for (int i = 0; i < N; i++) {
        r->F.mvKeysUn[i].pt.x = p2d[2 * i]; r->F.mvKeysUn[i].pt.y = p2d[2 * i + 1]; r->F.mvKeysUn[i].octave = octave[i];
        r->store.emplace_back(new MapPoint());
        MapPoint* p = r->store.back().get();
        p->mnId = i;
        p->mWorldPos = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) p->mWorldPos.at<float>(k) = p3d[3 * i + k];
        r->mps[i] = p;
    }
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
    r->solver.reset(new PnPsolver(r->F, r->mps));
    return r;
}
void ref_pnp_destroy(void* h) { delete (RefPnp*)h; }
void ref_pnp_set_ransac(void* h, double probability, int minInliers, int maxIterations, int minSet, float epsilon, float th2,
                        int* out_min_inliers, int* out_max_its) {
    PnPsolver& s = *((RefPnp*)h)->solver;
    s.SetRansacParameters(probability, minInliers, maxIterations, minSet, epsilon, th2);
    *out_min_inliers = s.mRansacMinInliers;
    *out_max_its = s.mRansacMaxIts;
}
int ref_pnp_iterations(void* h) { return ((RefPnp*)h)->solver->mnIterations; }
/* returns 1 when a pose came back (Tcw16 row-major 4x4 float), 0 for an empty Mat */
int ref_pnp_iterate(void* h, int nIterations, int* no_more, uint8_t* inliers, int* n_inliers, float* Tcw16) {
    RefPnp* r = (RefPnp*)h;
    bool bNoMore = false;
    std::vector<bool> vb;
    int nInl = 0;
    cv::Mat T = r->solver->iterate(nIterations, bNoMore, vb, nInl);
    *no_more = bNoMore ? 1 : 0;
    *n_inliers = nInl;
    for (size_t i = 0; i < r->mps.size(); i++) inliers[i] = i < vb.size() && vb[i] ? 1 : 0;
    if (T.empty()) return 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) Tcw16[4 * i + j] = T.at<float>(i, j);
    return 1;
}

/* ------------------------------------------------------------------ the stub's gemm, for pinning against cv2.gemm vectors */
void ref_cv_gemm_f32(const float* A, int ar, int ac, const float* B, int br, int bc, double alpha, const float* Cm, int cr, int cc,
                     double beta, int flags, float* D) {
    cv::Mat a(ar, ac, CV_32F, (void*)A), b(br, bc, CV_32F, (void*)B), c, d;
    if (Cm) c = cv::Mat(cr, cc, CV_32F, (void*)Cm);
    cv::gemm(a, b, alpha, c, beta, d, flags);
    for (int i = 0; i < d.rows; i++)
        for (int j = 0; j < d.cols; j++) D[i * d.cols + j] = d.at<float>(i, j);
}
/* `-R.t()*t` and `R*x+t` evaluated through the stub's cv::MatExpr exactly as ORBmatcher.cc:1483,1488,1503 write them */
void ref_expr_minus_Rt_t(const float* R9, const float* t3, float* out3) {
    cv::Mat R(3, 3, CV_32F, (void*)R9), t(3, 1, CV_32F, (void*)t3);
    const cv::Mat r = -R.t() * t;
    for (int i = 0; i < 3; i++) out3[i] = r.at<float>(i);
}
void ref_expr_Rx_plus_t(const float* R9, const float* x3, const float* t3, float* out3) {
    cv::Mat R(3, 3, CV_32F, (void*)R9), x(3, 1, CV_32F, (void*)x3), t(3, 1, CV_32F, (void*)t3);
    const cv::Mat r = R * x + t;
    for (int i = 0; i < 3; i++) out3[i] = r.at<float>(i);
}

} // extern "C"
