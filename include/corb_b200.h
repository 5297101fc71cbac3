/* libcorb_b200.so - C ABI of the B200-native (sm_100a) hot path of CORB-SLAM.
 *
 * Every entry point replaces one seam of the reference (paths relative to the reference tree):
 *   ORB front end      corbslam_client/include/ORBextractor.h:45-112, src/ORBextractor.cc:410-470,1043-1132
 *   Hamming / BoW      corbslam_client/include/ORBmatcher.h:41-67,  src/ORBmatcher.cc:162-423,657-790,1792-1808
 *   DBoW2              corbslam_client/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259,1338-1424,
 *                      ScoringObject.cpp:23-68, BowVector.cpp:34-84, FeatureVector.cpp:31-45
 *   EPnP-RANSAC        corbslam_client/include/PnPsolver.h:64-78,   src/PnPsolver.cc:66-383,420-962
 *   global BA          corbslam_client/include/Optimizer.h:42-46, src/Optimizer.cc:54-270 and the g2o slice under it
 *
 * Plain pointers and sizes only. All functions return a corb_status (0 = ok); corb_last_error() gives the text of
 * the calling thread's last failure. There is no CPU fallback: without a CUDA device every compute call fails.
 * Handles are independent (own stream, own device buffers); distinct handles may be used from distinct threads
 * concurrently (the reference runs the left and right extractor on two threads, Frame.cc:78-81).
 */
#ifndef CORB_B200_H
#define CORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CORB_API __attribute__((visibility("default")))

typedef enum {
    CORB_OK = 0,
    CORB_ERR_INVALID = 1,     /* bad argument */
    CORB_ERR_CUDA = 2,        /* CUDA runtime/driver failure (includes "no device") */
    CORB_ERR_UNSUPPORTED = 3, /* geometry the reference itself cannot process (e.g. image smaller than the FAST grid) */
    CORB_ERR_CAPACITY = 4,    /* an internal fixed-size buffer would overflow; nothing was truncated silently */
    CORB_ERR_STOPPED = 5,     /* the caller's stop flag ended an iterative solve early (results are still valid) */
    CORB_ERR_IO = 6
} corb_status;

CORB_API const char* corb_last_error(void);
CORB_API int corb_version(void);
/* number of CUDA devices visible, or 0 */
CORB_API int corb_device_count(void);

/* ------------------------------------------------------------------------------------------------ ORB extractor */

/* cv::KeyPoint layout (pt.x, pt.y, size, angle, response, octave, class_id), 28 bytes */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} corb_keypoint;

typedef struct corb_orb corb_orb;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)  [ORBextractor.cc:410-470;
 * created from the settings file at Tracking.cc:112-121]. `device` = CUDA ordinal (clientId-1 in the 1:1 topology). */
CORB_API int corb_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device,
                             corb_orb** out);
CORB_API void corb_orb_destroy(corb_orb* h);

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * [ORBextractor.h:63-85]. Arrays have corb_orb_levels() entries; any pointer may be NULL. `quota` is
 * mnFeaturesPerLevel, `umax16` the 16-entry disc table (both protected in the reference; exposed for tests). */
CORB_API int corb_orb_levels(const corb_orb* h);
CORB_API float corb_orb_scale_factor(const corb_orb* h);
CORB_API int corb_orb_tables(const corb_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota,
                             int* umax16);
/* size of pyramid level `level` for a w x h input [ComputePyramid, ORBextractor.cc:1111-1112] */
CORB_API int corb_orb_level_size(const corb_orb* h, int level, int w, int hgt, int* lw, int* lh);
/* upper bound of keypoints operator() can return for a w x h input (sum over levels of quota+3, or 4*nIni if larger;
 * DistributeOctTree may exceed the quota by up to 2 per level, ORBextractor.cc:539-763). <0 if unsupported. */
CORB_API int corb_orb_capacity(const corb_orb* h, int w, int hgt);

/* ORBextractor::operator()(image, mask (ignored), keypoints, descriptors)  [ORBextractor.cc:1043-1105; called from
 * Frame::ExtractORB, Frame.cc:247-253]. Host buffers in, host buffers out; blocks until the results are on the host.
 *   img/stride : 8-bit grey image (CV_8UC1), `stride` bytes per row
 *   kps, desc  : caller-owned, at least corb_orb_capacity() entries / x32 bytes
 *   n          : number of keypoints written (0 for an empty image, like the reference's silent return :1046)
 *   pyr_out    : NULL, or corb_orb_levels() host pointers receiving mvImagePyramid[l] (dense rows, lw*lh bytes each;
 *                a NULL entry skips that level) - needed while Frame::ComputeStereoMatches stays on the CPU */
CORB_API int corb_orb_extract(corb_orb* h, const uint8_t* img, int w, int hgt, int stride, corb_keypoint* kps, uint8_t* desc,
                              int* n, uint8_t* const* pyr_out);

/* Split form of the same call so one thread can keep several handles (left/right, or several frames) in flight:
 * _submit enqueues H2D + kernels + D2H on the handle's stream and returns; _wait blocks and fills the outputs. */
CORB_API int corb_orb_extract_submit(corb_orb* h, const uint8_t* img, int w, int hgt, int stride, int want_pyramid);
CORB_API int corb_orb_extract_wait(corb_orb* h, corb_keypoint* kps, uint8_t* desc, int* n, uint8_t* const* pyr_out);

/* Both images of a stereo frame from ONE thread: the two per-frame graphs are launched back to back and both results
 * collected - what Frame::Frame does with two std::threads (Frame.cc:78-81). Arguments as corb_orb_extract, per side. */
CORB_API int corb_orb_extract_pair(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt,
                                   int stride, corb_keypoint* kps_l, uint8_t* desc_l, int* n_l, corb_keypoint* kps_r,
                                   uint8_t* desc_r, int* n_r, uint8_t* const* pyr_l, uint8_t* const* pyr_r);
/* Split form of corb_orb_extract_pair, so that ONE client thread keeps two stereo frames in flight on two handle pairs
 * (frame i + 1 is read over PCIe and extracted while the caller consumes frame i): _submit enqueues everything on the left
 * handle's stream and returns, _wait blocks and fills the outputs (NULL kps / desc: read them in place through
 * corb_orb_host_results). Both images must be non-empty. */
CORB_API int corb_orb_extract_pair_submit(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt,
                                          int stride, int want_pyramid);
CORB_API int corb_orb_extract_pair_wait(corb_orb* hl, corb_orb* hr, corb_keypoint* kps_l, uint8_t* desc_l, int* n_l,
                                        corb_keypoint* kps_r, uint8_t* desc_r, int* n_r, uint8_t* const* pyr_l, uint8_t* const* pyr_r);
CORB_API int corb_orb_extract_pair_device(corb_orb* hl, corb_orb* hr, const uint8_t* d_img_l, const uint8_t* d_img_r, int w,
                                          int hgt, int stride);

/* Frame::ComputeStereoMatches [Frame.cc:470-644] on the results of the last extraction of `left` and `right`, which are
 * still resident in HBM (keypoints, descriptors and both un-blurred pyramids): u_right[i] = mvuRight[i], depth[i] =
 * mvDepth[i] for the n_left left keypoints, -1 where there is no match. mbf, mb as in Frame (mbf = baseline * fx, mb = mbf/fx).
 * Both handles must have completed an extraction of the same image size on the same device. */
CORB_API int corb_stereo_match(corb_orb* left, corb_orb* right, float mbf, float mb, int n_left, float* u_right, float* depth);
/* The heavy part of the stereo Frame constructor [Frame.cc:61-117] in one call: ExtractORB left and right, then
 * ComputeStereoMatches, with nothing but the results crossing PCIe (the pyramids stay on the device). */
CORB_API int corb_frame_stereo(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride,
                               float mbf, float mb, corb_keypoint* kps_l, uint8_t* desc_l, int* n_l, corb_keypoint* kps_r,
                               uint8_t* desc_r, int* n_r, float* u_right, float* depth);

/* Split form of corb_frame_stereo (see corb_orb_extract_pair_submit). */
CORB_API int corb_frame_stereo_submit(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride,
                                      float mbf, float mb);
CORB_API int corb_frame_stereo_wait(corb_orb* hl, corb_orb* hr, corb_keypoint* kps_l, uint8_t* desc_l, int* n_l, corb_keypoint* kps_r,
                                    uint8_t* desc_r, int* n_r, float* u_right, float* depth);

/* Device-resident form: `d_img` is already in HBM (pitch `stride`), results stay in HBM for on-GPU consumers
 * (matcher, stereo). Enqueued on the handle's stream; corb_orb_sync() waits for it. */
CORB_API int corb_orb_extract_device(corb_orb* h, const uint8_t* d_img, int w, int hgt, int stride);
CORB_API int corb_orb_sync(corb_orb* h);
/* device pointers of the last extraction: keypoints [capacity], descriptors [capacity*32], count [1] */
CORB_API int corb_orb_device_results(const corb_orb* h, const corb_keypoint** d_kps, const uint8_t** d_desc,
                                     const int** d_count);
/* Zero-copy form of the host results: the extraction calls accept NULL for `kps` / `desc`; the results of the last
 * completed extraction on `h` can then be read in place from the handle's page-locked result buffer (where the D2H node
 * of the frame graph put them). Valid until the next extraction on the handle. The shim builds std::vector<cv::KeyPoint>
 * straight from these (one pass over the data instead of two). */
CORB_API int corb_orb_host_results(const corb_orb* h, const corb_keypoint** kps, const uint8_t** desc, int* n);
/* device pointer + pitch of pyramid level l (un-blurred if blurred == 0) of the last extraction */
CORB_API int corb_orb_device_level(const corb_orb* h, int level, int blurred, const uint8_t** d_ptr, int* pitch, int* lw,
                                   int* lh);
/* CUDA stream (cudaStream_t) the handle works on, for event timing by the caller */
CORB_API void* corb_orb_stream(const corb_orb* h);
/* number of kernel launches one extraction enqueues (graph nodes that are kernels) */
CORB_API int corb_orb_launches_per_extract(const corb_orb* h);
/* 1 if the FAST kernel stages its tiles with TMA (cp.async.bulk.tensor) for the current plan; 0 if the driver entry point
 * was unavailable or CORB_NO_TMA is set in the environment (plain vector loads then; same results) */
CORB_API int corb_orb_uses_tma(const corb_orb* h);
/* How page-locked host images reach the GPU (the results are the same):
 *   0 (default)  the import kernel reads them in place over PCIe (mapped memory) and every image gets its own launches, so the
 *                left image's pipeline starts while the right one is still arriving: lowest latency of ONE blocking frame
 *   1            copy-engine memcpy nodes into a staging buffer, both images in every launch: the DMA engines sustain about
 *                twice the PCIe rate of SM loads, which is what counts with several frames in flight (submit / wait on
 *                independent handle pairs): 25 k -> 31 k stereo frames/s through host buffers on a B200, 8 in flight
 * Set it on BOTH handles of a pair. CORB_H2D_NODE=1 in the environment makes 1 the default. */
CORB_API int corb_orb_set_host_transfer(corb_orb* h, int mode);
CORB_API int corb_orb_host_transfer(const corb_orb* h);

/* Per-kernel device time (ms, averaged over `reps` eager replays with CUDA events between launches) of one extraction
 * of the image currently resident; *n kernels in launch order, named by corb_orb_kernel_name(). For roofline reports. */
CORB_API int corb_orb_profile(corb_orb* h, int reps, float* ms, int cap, int* n);
CORB_API const char* corb_orb_kernel_name(const corb_orb* h, int i);

/* Stage taps of the last extraction, for stage-wise parity tests against the oracle.
 *   CORB_TAP_PYRAMID / CORB_TAP_BLURRED : out = lw*lh bytes (dense)
 *   CORB_TAP_CANDIDATES                 : out = int32 triplets (x, y, response) relative to (16,16), in the order
 *                                         vToDistributeKeys is built (ORBextractor.cc:789-832); *n = count
 *   CORB_TAP_LEVEL_COUNT                : *n = keypoints kept on that level */
enum { CORB_TAP_PYRAMID = 0, CORB_TAP_BLURRED = 1, CORB_TAP_CANDIDATES = 2, CORB_TAP_LEVEL_COUNT = 3 };
CORB_API int corb_orb_tap(corb_orb* h, int what, int level, void* out, size_t out_bytes, int* n);

/* ------------------------------------------------------------------------------------------------ descriptor matching */

/* Workspace for the matching entry points (device buffers + stream). One per calling thread: Tracking, LoopClosing and
 * the server's fuse thread call SearchByBoW concurrently. */
typedef struct corb_matcher corb_matcher;
CORB_API int corb_matcher_create(int device, corb_matcher** out);
CORB_API void corb_matcher_destroy(corb_matcher* m);

/* ORBmatcher::DescriptorDistance [ORBmatcher.cc:1792-1808], batched: out[i] = Hamming(A[pairs[2i]], B[pairs[2i+1]])
 * over 256-bit descriptors stored as 32-byte rows. Scalar callers keep the inline host popcount. */
CORB_API int corb_hamming_pairs(corb_matcher* m, const uint8_t* A, int nA, const uint8_t* B, int nB, const int32_t* pairs,
                                int n, int32_t* out);

/* ---- ORBmatcher::SearchByProjection -------------------------------------------------------
 * Flattened view of the reference's Frame for the projection matchers (Frame.h:60-200), filled by the shim:
 *   x, y, octave, angle   mvKeysUn[i].pt / .octave / .angle
 *   desc                  mDescriptors (n x 32 bytes);  u_right  mvuRight
 *   taken[i]              mvpMapPoints[i] holds a MapPoint with Observations() > 0: the feature is skipped
 *                         (ORBmatcher.cc:81-84, 1554-1557); NULL = none
 *   grid_off / grid_idx   Frame::mGrid[64][48] (Frame.cc:229-245) as CSR, cell = ix * 48 + iy, indices in push_back order
 *   min/max, grid_*_inv   mnMinX.. / mfGridElementWidthInv.. ;  scale_factors  mvScaleFactors
 *   fx..mb, Tcw           intrinsics, stereo baseline terms, rows 0..2 of mTcw (row-major 3x4) */
typedef struct {
    int32_t n;
    const float* x;
    const float* y;
    const int32_t* octave;
    const float* angle;
    const uint8_t* desc;
    const float* u_right;
    const uint8_t* taken;
    const int32_t* grid_off;
    const int32_t* grid_idx;
    float min_x, min_y, max_x, max_y, grid_w_inv, grid_h_inv;
    const float* scale_factors;
    int32_t n_levels;
    float fx, fy, cx, cy, mbf, mb;
    float Tcw[12];
} corb_frame_view;

/* int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
 * [ORBmatcher.h:55, ORBmatcher.cc:1470-1614]; called by Tracking::TrackWithMotionModel (Tracking.cc:872-882).
 * Per last-frame feature i: last_valid = mvpMapPoints[i] && !mvbOutlier[i]; last_blocks = pMP->Observations() > 0
 * (NULL = all; temporal stereo points have none); last_xyz = pMP->GetWorldPos(); last_mp_desc = pMP->GetDescriptor();
 * last_octave = mvKeys[i].octave; last_angle = mvKeysUn[i].angle; Tlw = LastFrame.mTcw rows 0..2.
 * match[i2] (size cur->n) = index i of the last-frame feature whose MapPoint current feature i2 received, else -1. */
CORB_API int corb_search_by_projection_last(corb_matcher* m, const corb_frame_view* cur, int32_t n_last, const uint8_t* last_valid,
                                            const uint8_t* last_blocks, const float* last_xyz, const uint8_t* last_mp_desc,
                                            const int32_t* last_octave, const float* last_angle, const float* Tlw, float th,
                                            int mono, int check_orientation, int32_t* match, int32_t* nmatches);

/* int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th)
 * [ORBmatcher.h:51, ORBmatcher.cc:44-131]; called by Tracking::SearchLocalPoints (Tracking.cc:1199-1212).
 * Per map point: in_view = mbTrackInView && !isBad(); blocks = Observations() > 0 (NULL = all);
 * proj = (mTrackProjX, mTrackProjY, mTrackProjXR); level = mnTrackScaleLevel; view_cos = mTrackViewCos.
 * match[idx] (size F->n) = index of the map point assigned to frame feature idx, else -1. */
CORB_API int corb_search_by_projection_map(corb_matcher* m, const corb_frame_view* F, int32_t n_mp, const uint8_t* in_view,
                                           const uint8_t* blocks, const float* proj, const int32_t* level, const float* view_cos,
                                           const uint8_t* mp_desc, float th, float nnratio, int32_t* match, int32_t* nmatches);

/* One side of a SearchByBoW call, flattened by the shim:
 *   desc/n            mDescriptors (n x 32 bytes)
 *   fv_*              DBoW2::FeatureVector as CSR: fv_nodes[fv_n] ascending node ids, fv_off[fv_n+1], fv_idx[] feature
 *                     indices in the order FeatureVector::addFeature appended them
 *   valid             per feature: MapPoint* != NULL && !isBad()  (NULL = all valid)
 *   angles            per feature keypoint angle in degrees (mvKeysUn / mvKeys as the variant prescribes); may be NULL
 *                     when check_ori == 0 */
typedef struct {
    const uint8_t* desc;
    int32_t n;
    const uint32_t* fv_nodes;
    const int32_t* fv_off;
    const uint32_t* fv_idx;
    int32_t fv_n;
    const uint8_t* valid;
    const float* angles;
} corb_bow_side;

enum {
    CORB_BOW_KF_FRAME = 0,  /* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)        [ORBmatcher.cc:162-291] */
    CORB_BOW_KF_SERVER = 1, /* ORBmatcher::SearchByBoWInServer(KeyFrame*, KeyFrame*)  [ORBmatcher.cc:294-423] */
    CORB_BOW_KF_KF = 2      /* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...)     [ORBmatcher.cc:657-790] */
};

/* A = the keyframe whose MapPoints are handed out (pKF / pKF1), B = Frame F / KeyFrame F / pKF2.
 * Variants 0,1: match has B->n entries, match[b] = index of the A feature whose MapPoint goes to b, or -1.
 * Variant 2   : match has A->n entries, match[a] = index of the matched B feature, or -1; B->valid is honoured.
 * nnratio / check_ori = ORBmatcher(nnratio, checkOri) [ORBmatcher.cc:41]. *nmatches = the reference's return value. */
CORB_API int corb_bow_match(corb_matcher* m, int variant, const corb_bow_side* A, const corb_bow_side* B, float nnratio,
                            int check_ori, int32_t* match, int32_t* nmatches);
/* The same for ncalls independent (A[i], B[i]) pairs in one launch (relocalisation candidates, Tracking.cc:1405;
 * map-fusion candidates, MapFusion.cpp:691). match[i] / nmatches[i] as above. */
CORB_API int corb_bow_match_batch(corb_matcher* m, int variant, int ncalls, const corb_bow_side* A, const corb_bow_side* B,
                                  float nnratio, int check_ori, int32_t* const* match, int32_t* nmatches);
/* Device-resident form of the batch: every pointer inside A[i], B[i] and match[i] is a device pointer; results stay in
 * HBM (d_nmatches[ncalls] on the device). Enqueued on the matcher's stream; corb_matcher_sync() waits. */
CORB_API int corb_bow_match_batch_device(corb_matcher* m, int variant, int ncalls, const corb_bow_side* A,
                                         const corb_bow_side* B, float nnratio, int check_ori, int32_t* const* d_match,
                                         int32_t* d_nmatches);
CORB_API int corb_matcher_sync(corb_matcher* m);
CORB_API void* corb_matcher_stream(const corb_matcher* m);

/* ------------------------------------------------------------------------------------------------ DBoW2 vocabulary */

typedef struct corb_voc corb_voc;

/* ORBVocabulary::loadFromTextFile [TemplatedVocabulary.h:1338-1424; System.cc:59-68]. The tree is uploaded once
 * (about 35 MB for ORBvoc.txt: 1 082 073 nodes) and is read-only afterwards. Only L1_NORM scoring with TF_IDF or TF
 * weighting is supported (ORBvoc.txt's header is "10 6 0 0"). */
CORB_API int corb_voc_load_text(const char* path, int device, corb_voc** out);
/* Same from arrays: nodes 1..n in DBoW2 id order (parent id, leaf flag, 32-byte descriptor, weight). */
CORB_API int corb_voc_create(int k, int L, int scoring, int weighting, int n, const int32_t* parent, const uint8_t* is_leaf,
                             const uint8_t* desc, const double* weight, int device, corb_voc** out);
CORB_API void corb_voc_destroy(corb_voc* v);
CORB_API int corb_voc_info(const corb_voc* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words);

/* TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup) for n features [:1218-1259]:
 * word id, word weight (idf) and the id of the ancestor node at level L-levelsup, per feature. */
CORB_API int corb_voc_transform_features(corb_voc* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id,
                                         double* weight, uint32_t* node_id);
/* TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) [:1127-1194; Frame.cc:404,
 * KeyFrame.cc:75]: the descent runs on the GPU, the two maps are rebuilt on the host in feature order so the fp64
 * sums are bit-identical. bow_* need n entries, fv_nodes n, fv_off n+1, fv_idx n. */
CORB_API int corb_voc_transform(corb_voc* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_words, double* bow_vals,
                                int* n_bow, uint32_t* fv_nodes, int32_t* fv_off, uint32_t* fv_idx, int* n_fv);
/* ORBVocabulary::score(v1, v2) = L1Scoring::score [ScoringObject.cpp:23-68; KeyFrameDatabase.cc:128,238,345], one
 * query against ncand candidates in one launch. BowVectors are sorted (word id, value) lists. */
CORB_API int corb_bow_score_batch(corb_voc* v, const uint32_t* q_words, const double* q_vals, int nq, int ncand,
                                  const uint32_t* const* c_words, const double* const* c_vals, const int32_t* c_n,
                                  double* scores);

/* ------------------------------------------------------------------------------------------------ EPnP-RANSAC (PnPsolver) */

/* PnPsolver::SetRansacParameters(probability, minInliers, maxIterations, minSet, epsilon, th2) [PnPsolver.cc:163-198]:
 * the adjusted mRansacMinInliers and mRansacMaxIts for N correspondences (host scalar logic, no device needed). */
CORB_API int corb_pnp_ransac_params(int N, double probability, int min_inliers, int max_iterations, int min_set, float epsilon,
                                    int* out_min_inliers, int* out_max_its);

/* One PnPsolver object flattened by the shim [constructor PnPsolver.cc:66-151]: the matches with a live MapPoint
 * (pMP && !pMP->isBad()), in mvKeyPointIndices order. The reference draws its minimal sets from the process-global
 * rand() stream (DUtils::Random::RandomInt, Random.cpp:47-50), so its result is not a function of its arguments; here
 * the draws are an argument: draws[4 * it + k] is the value RandomInt(0, vAvailableIndices.size() - 1) returns for pick k
 * of RANSAC iteration `it` (0-based over the life of the solver object), i.e. 0 <= draws[4 * it + k] < n - k. The shim
 * draws them with the same RandomInt calls, for iterations [0, it_end), it_end = max(max_its, iterations_done +
 * n_iterations) - the loop condition of iterate() (:223). The call is stateless: iterations [0, iterations_done) are
 * recomputed (the best-so-far state of the reference object is a function of them).
 * Deviation: drawing for all iterations up front consumes more of the process-global rand() stream than the reference, which
 * draws four values per EXECUTED iteration and stops at an early Refine() return; a second solver in the same process
 * therefore starts at a later stream position than it would in the reference process (same distribution, different draws). */
typedef struct {
    int32_t n;               /* N = mvP2D.size() */
    const float* p2d;        /* [n][2] mvP2D: undistorted keypoint positions */
    const float* p3d;        /* [n][3] mvP3Dw: MapPoint world positions */
    const float* max_err;    /* [n] mvMaxError = mvLevelSigma2[octave] * th2 (:194-197) */
    float fx, fy, cx, cy;    /* :107-110 */
    int32_t min_inliers;     /* adjusted mRansacMinInliers */
    int32_t max_its;         /* adjusted mRansacMaxIts */
    int32_t iterations_done; /* mnIterations before this call */
    int32_t n_iterations;    /* the nIterations argument of iterate() */
    const int32_t* draws;    /* [it_end][4] */
} corb_pnp_problem;

typedef struct {
    int32_t status;     /* 0 = empty cv::Mat; 1 = refined pose (early return, :262-273); 2 = best pose when the iterations
                           are exhausted (:279-291) */
    int32_t no_more;    /* bNoMore */
    int32_t n_inliers;  /* nInliers */
    int32_t iterations; /* mnIterations after the call */
    float Tcw[16];      /* row-major 4x4 CV_32F, valid when status != 0 */
} corb_pnp_result;

/* PnPsolver::iterate(nIterations, bNoMore, vbInliers, nInliers) [PnPsolver.cc:206-300] for a batch of solver objects -
 * the relocalisation / map-fusion candidates of Tracking.cc:1432-1440 and MapFusion.cpp:720-728 - in one call: every
 * RANSAC hypothesis of every candidate is evaluated at once (EPnP on 4 points: a team of 8 lanes per hypothesis in small
 * batches, one thread in large ones), inlier counting is one warp per hypothesis, Refine() one warp per best-so-far record;
 * the sequential bookkeeping of the reference (best so far, the first iteration whose Refine() succeeds) is resolved
 * afterwards from the per-hypothesis results, so the outcome equals the sequential loop on the same draws. inliers[c]
 * (nullable) receives n bytes: vbInliers over the flattened correspondences (the shim scatters them through
 * mvKeyPointIndices). One handle must not be used from two threads at once. */
CORB_API int corb_pnp_iterate_batch(corb_matcher* m, int n_problems, const corb_pnp_problem* problems, corb_pnp_result* results,
                                    uint8_t* const* inliers);

/* ------------------------------------------------------------------------------------------------ global bundle adjustment */

/* Optimizer::BundleAdjustment flattened by the shim [Optimizer.cc:54-270]: the pointer graph of KeyFrames, MapPoints
 * and observations becomes dense-index SoA arrays (the sparse vertex ids mnId / mnId+maxKFid+1 are only local indices).
 *   pose_q/pose_t   world->camera pose of each keyframe: unit quaternion (x,y,z,w) and translation, fp64
 *                   (Converter::toSE3Quat of the float32 4x4, Converter.cc:37-47); updated in place
 *   pose_fixed      mnId == 1 || getFixed()  (Optimizer.cc:94)
 *   pose_cam        fx, fy, cx, cy, bf of the keyframe (:164-167, :186-190)
 *   point_xyz       MapPoint world positions (:113), updated in place; point_fixed = getFixed() (:120)
 *   edge_*          one entry per observation: keyframe index, point index, (u, v, u_right) with u_right < 0 for a
 *                   monocular observation (mvuRight < 0, :140), information = invSigma2 of the keypoint octave
 * A (pose, point) pair may appear at most once (MapPoint::GetObservations is a map). Points without edges are ignored. */
typedef struct {
    int32_t n_poses, n_points, n_edges;
    double* pose_q;
    double* pose_t;
    const uint8_t* pose_fixed;
    const double* pose_cam;
    double* point_xyz;
    const uint8_t* point_fixed;
    const int32_t* edge_pose;
    const int32_t* edge_point;
    const double* edge_obs;
    const double* edge_inv_sigma2;
} corb_ba_problem;

typedef struct {
    int32_t iterations;       /* LM iterations executed (what g2o's optimize() returns) */
    int32_t n_trials;         /* LM trials = linear solves */
    int32_t stopped;          /* 1 if *stop ended the solve */
    int32_t solver_failures;  /* trials whose reduced system was not positive definite */
    double chi2_initial, chi2_final, lambda_initial, lambda_final;
    uint8_t trial_accepted[256]; /* accept/reject of the first 256 trials */
    double trial_chi2[256];
    double ms_total, ms_solve;   /* wall time of the whole call / device time inside the reduced-camera solves */
    int64_t reduced_blocks;      /* 6x6 blocks in the envelope of the reduced camera system */
    int32_t border_poses;        /* keyframes ordered last because they carry long-range (loop/fusion) links */
    int32_t max_active_rows;     /* widest front of the block-skyline factorisation */
    double ms_setup;             /* host structure build + uploads before the first LM iteration */
    int32_t band_chunks;         /* independent chunks the band was cut into (separator keyframes join the border) */
    int32_t separator_poses;     /* keyframes ordered last only to decouple the chunks */
    int32_t schur_pair_lists;    /* 1: Schur rows from the sorted pair lists built once per call; 0: per-iteration edge walk (fallback) */
    int32_t reserved0;
} corb_ba_result;

/* corb_ba_solve keeps its device buffers in a per-device arena between calls (a global BA allocates ~40 buffers; the
 * cudaMalloc / cudaFree pairs cost as much as the optimisation). This returns the arena of `device` to the driver;
 * a call that is running on the device keeps its memory and the arena is released when it ends. */
CORB_API int corb_ba_release_cache(int device);
/* Page-locked host memory for the caller's flat arrays. corb_ba_solve uploads page-locked edge arrays with the DMA engines in
 * place (40 B per observation: 0.8 ms instead of ~4 ms per million observations out of pageable memory, which is most of the
 * set-up of a global BA); the C++ shim flattens the graph into such buffers (shim/Optimizer_gba.h). Pageable arrays work too. */
CORB_API int corb_host_alloc(size_t bytes, void** out);
CORB_API int corb_host_free(void* p);

/* All-reduce hook for landmark-sharded BA: `buf` is a DEVICE pointer to n doubles, reduced in place over
 * all ranks with op 0 = sum, 1 = min, 2 = max, ordered on CUDA stream `stream` (a cudaStream_t). The C++ server passes
 * a function that calls ncclAllReduce on its communicator; the Python host layer passes torch.distributed.all_reduce. */
typedef int (*corb_allreduce_fn)(void* user, double* d_buf, size_t n, int op, void* stream);

/* Optimizer::BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust) [Optimizer.h:42-46] = g2o
 * BlockSolver_6_3 + Levenberg-Marquardt [block_solver.hpp:354-604, optimization_algorithm_levenberg.cpp:61-189] on the
 * GPU. `stop` (nullable) is polled before every iteration and LM trial like g2o's forceStopFlag. robust != 0 adds the
 * Huber kernels of Optimizer.cc:101-102,155-160,179-184. With allreduce != NULL the problem holds this rank's landmark
 * shard (all poses, a subset of points and their edges) and every rank ends with identical poses; the stop decision is then
 * collective: each rank's flag is sampled where chi2 is reduced and summed over the ranks next to it, and only the reduced
 * value is branched on, so ranks that see the flag at different times still leave the loop together. */
CORB_API int corb_ba_solve(corb_ba_problem* p, int iterations, const volatile uint8_t* stop, int robust, int device,
                           corb_ba_result* result, corb_allreduce_fn allreduce, void* allreduce_user);

/* ------------------------------------------------------------------------------------------------ device-resident BoW record */

/* The matching record of a Frame / KeyFrame kept in HBM: descriptors, keypoint angles, BowVector, FeatureVector. It is what
 * Frame::ComputeBoW [Frame.cc:399-406] / KeyFrame::ComputeBoW [KeyFrame.cc:70-77] produce, built straight from the extractor's
 * device-resident results (no descriptor round trip through the host, no std::map rebuild), and what SearchByBoW and the
 * L1 score then read in place: relocalisation / map-fusion candidates stay on the GPU between calls.
 *   corb_frame_bow        : record <- last extraction of `h` (must be complete) through vocabulary `v`, levelsup as Frame.cc:404
 *   corb_bow_store_fill   : the same from any device arrays (d_kps may be NULL: angles 0)
 *   corb_bow_store_side   : the record as a corb_bow_side of DEVICE pointers for corb_bow_match_batch_device (d_valid: device
 *                           bytes, MapPoint liveness, or NULL)
 *   corb_bow_store_download: BowVector / FeatureVector to the host (key order of the std::maps; fp64 bits of DBoW2)
 *   corb_bow_score_stores : voc.score(query, cand[i]) for records [KeyFrameDatabase.cc:238], one launch */
typedef struct corb_bow_store corb_bow_store;
CORB_API int corb_bow_store_create(int device, int capacity, corb_bow_store** out);
CORB_API void corb_bow_store_destroy(corb_bow_store* s);
CORB_API int corb_bow_store_fill(corb_bow_store* s, corb_voc* v, const corb_keypoint* d_kps, const uint8_t* d_desc, int n, int levelsup,
                                 void* stream);
CORB_API int corb_frame_bow(corb_orb* h, corb_voc* v, int levelsup, corb_bow_store* s);
/* corb_frame_bow / corb_bow_store_fill return once the work is enqueued; every call that reads a record waits for its fill
 * first (corb_bow_store_side / _download, corb_bow_match_stores, corb_bow_score_stores), and so does this one explicitly.
 * A record that several threads will read must be synced once by the thread that filled it before it is shared. */
CORB_API int corb_bow_store_sync(corb_bow_store* s);
CORB_API int corb_bow_store_features(const corb_bow_store* s);  /* features of the last fill (known when it is enqueued) */
CORB_API int corb_bow_store_side(const corb_bow_store* s, const uint8_t* d_valid, corb_bow_side* side, int* n_bow);
CORB_API int corb_bow_store_download(const corb_bow_store* s, uint32_t* bow_words, double* bow_vals, int* n_bow, uint32_t* fv_nodes,
                                     int32_t* fv_off, uint32_t* fv_idx, int* n_fv);
/* SearchByBoW (variant as corb_bow_match) between records: validA[i] / validB[i] are HOST byte masks (MapPoint liveness, may be
 * NULL), match[i] / nmatches are HOST outputs; the descriptors and feature vectors never leave the GPU. */
CORB_API int corb_bow_match_stores(corb_matcher* m, int variant, int ncalls, const corb_bow_store* const* A, const uint8_t* const* validA,
                                   const corb_bow_store* const* B, const uint8_t* const* validB, float nnratio, int check_ori,
                                   int32_t* const* match, int32_t* nmatches);
CORB_API int corb_bow_score_stores(corb_voc* v, const corb_bow_store* query, int ncand, const corb_bow_store* const* cands, double* scores);

/* ------------------------------------------------------------------------------------------------ keyframe payload */

/* Binary form of the extractor-produced part of a KeyFrame on the wire - mvKeys, mvKeysUn, mvuRight, mvDepth, mDescriptors,
 * mBowVec, mFeatVec - instead of the decimal text boost::archive::text_oarchive writes for the same members
 * [KeyFrame.h:61-87 (serialize), SerializeObject.h:34-61 (cv::Mat / cv::KeyPoint), DataDriver.cc:40-238]. One little-endian
 * blob: 40-byte header (magic "CKF\1", version, flags, counts, CRC-32 of the body) + SoA arrays; bit-exact round trip,
 * ~76 bytes per keypoint (61 without the BoW / feature vectors) instead of >= 236 characters. It travels inside the existing srv `DATA` string (version tag = the
 * magic; a boost text archive never starts with it), so the ROS srv/msg surface is unchanged. Host code (no GPU needed).
 * keys_un may be NULL or equal keys (rectified stereo: mvKeysUn == mvKeys, Frame.cc:410-414) and is then not stored. */
CORB_API size_t corb_kf_payload_bound(int n, int n_bow, int n_fv, int n_fv_idx);
CORB_API int corb_kf_payload_encode(const corb_keypoint* keys, const corb_keypoint* keys_un, const float* u_right, const float* depth,
                                    const uint8_t* desc, int n, const uint32_t* bow_words, const double* bow_vals, int n_bow,
                                    const uint32_t* fv_nodes, const int32_t* fv_off, const uint32_t* fv_idx, int n_fv, uint8_t* out,
                                    size_t cap, size_t* written);
CORB_API int corb_kf_payload_info(const uint8_t* buf, size_t len, int* n, int* n_bow, int* n_fv, int* n_fv_idx, int* same_un);
CORB_API int corb_kf_payload_decode(const uint8_t* buf, size_t len, corb_keypoint* keys, corb_keypoint* keys_un, float* u_right,
                                    float* depth, uint8_t* desc, uint32_t* bow_words, double* bow_vals, uint32_t* fv_nodes,
                                    int32_t* fv_off, uint32_t* fv_idx);

#ifdef __cplusplus
}
#endif
#endif /* CORB_B200_H */
