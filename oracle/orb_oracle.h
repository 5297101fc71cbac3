/* CPU ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Plain C++17 restatement (no OpenCV, no Eigen) of the reference's ORB front end:
 *   corbslam_client/src/ORBextractor.cc:77-146,410-853,1034-1132  (reference, read-only)
 * plus closed-form models of the un-vendored OpenCV primitives it calls (FAST-9/16, resize
 * INTER_LINEAR u8, GaussianBlur 7x7 sigma 2 u8, fastAtan2, cvRound), pinned to cv2 4.13.0
 * semantics by tests/test_oracle_primitives.py and the fixtures in tests/golden/.
 *
 * PARITY STATUS: PINNED. The reference ships no golden vectors; the OpenCV primitive
 * models are pinned against cv2 4.13.0 (tests/golden/), and the composition (cell grid, quadtree, ordering,
 * orientation, descriptors, stereo matching) is pinned byte for byte against the reference's own
 * ORBextractor.cc / Frame.cc compiled unmodified into oracle/_ref/libref.so (oracle/refbuild,
 * tests/test_ref_cpu.py). The one canonical choice: quadtree ties are split in creation order, where the
 * reference compares heap addresses (ORBextractor.cc:591,627,684) - libref runs with a creation-ordered
 * node allocator; with plain malloc ~1 % of a frame's keypoints change.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * use this code. The product (libcorb_b200.so) never links or calls it.
 */
#ifndef CORB_ORB_ORACLE_H
#define CORB_ORB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} oracle_keypoint; /* 28 B, field order of cv::KeyPoint */

typedef struct oracle_orb oracle_orb;

/* ORBextractor::ORBextractor (ORBextractor.cc:410-470) */
oracle_orb* oracle_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void oracle_orb_destroy(oracle_orb*);
int oracle_orb_levels(const oracle_orb*);
void oracle_orb_tables(const oracle_orb*, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                       int* quota, int* umax16);
/* level size for an input of w x h (ComputePyramid :1111-1112) */
void oracle_orb_level_size(const oracle_orb*, int level, int w, int h, int* lw, int* lh);
int oracle_orb_capacity(const oracle_orb*, int w, int h);

/* ORBextractor::operator() (:1043-1105). kps/desc sized oracle_orb_capacity(). Returns #keypoints, <0 on error. */
int oracle_orb_extract(oracle_orb*, const uint8_t* img, int w, int h, int stride,
                       oracle_keypoint* kps, uint8_t* desc);

/* stage taps for stage-wise parity (valid after oracle_orb_extract) */
const uint8_t* oracle_orb_pyramid(const oracle_orb*, int level, int* w, int* h);   /* un-blurred, dense rows */
const uint8_t* oracle_orb_blurred(const oracle_orb*, int level, int* w, int* h);   /* blurred (only levels with kps) */
/* candidates before DistributeOctTree, coordinates relative to (minBorderX,minBorderY): x,y,response triplets */
int oracle_orb_candidates(const oracle_orb*, int level, const int32_t** xyr);
int oracle_orb_level_count(const oracle_orb*, int level);

/* Frame::ComputeStereoMatches (Frame.cc:470-644) on the last extraction of the two handles; returns #matches kept */
int oracle_stereo_matches(const oracle_orb* L, const oracle_orb* R, const oracle_keypoint* kl, const uint8_t* dl, int nl,
                          const oracle_keypoint* kr, const uint8_t* dr, int nr, float mbf, float mb, float* u_right, float* depth);

/* primitives, exposed for pinning against cv2 */
void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void oracle_gaussian7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
/* FAST-9/16 score map: 0 where not computable (3 px rim); score = max arc-min - 1 (>=0 everywhere inside) */
void oracle_fast_score(const uint8_t* src, int w, int h, int sstride, uint8_t* score, int scstride);
/* cv::FAST(img, kps, th, true): returns count; out = (x,y,response) triplets in raster order */
int oracle_fast_detect(const uint8_t* src, int w, int h, int sstride, int th, int32_t* out, int cap);
float oracle_fast_atan2(float y, float x);
int oracle_cv_round_f(float v);
/* IC_Angle (:77-104) on a dense image */
float oracle_ic_angle(const uint8_t* img, int stride, int x, int y, const int* umax16);
/* computeOrbDescriptor (:107-146) on an (already blurred) dense image */
void oracle_brief(const uint8_t* img, int stride, int x, int y, float angle_deg, uint8_t* desc32);

#ifdef __cplusplus
}
#endif
#endif
