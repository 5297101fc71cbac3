/* libcorb_b200.so - C ABI of the B200-native (sm_100a) hot path of CORB-SLAM.
 *
 * Every entry point replaces one seam of the reference (paths relative to the reference tree):
 *   ORB front end      corbslam_client/include/ORBextractor.h:45-112, src/ORBextractor.cc:410-470,1043-1132
 *   Hamming / BoW      corbslam_client/include/ORBmatcher.h:41-67,  src/ORBmatcher.cc:162-423,657-790,1792-1808
 *   DBoW2              corbslam_client/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259,1338-1424,
 *                      ScoringObject.cpp:23-68, BowVector.cpp:34-84, FeatureVector.cpp:31-45
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

/* Device-resident form: `d_img` is already in HBM (pitch `stride`), results stay in HBM for on-GPU consumers
 * (matcher, stereo). Enqueued on the handle's stream; corb_orb_sync() waits for it. */
CORB_API int corb_orb_extract_device(corb_orb* h, const uint8_t* d_img, int w, int hgt, int stride);
CORB_API int corb_orb_sync(corb_orb* h);
/* device pointers of the last extraction: keypoints [capacity], descriptors [capacity*32], count [1] */
CORB_API int corb_orb_device_results(const corb_orb* h, const corb_keypoint** d_kps, const uint8_t** d_desc,
                                     const int** d_count);
/* device pointer + pitch of pyramid level l (un-blurred if blurred == 0) of the last extraction */
CORB_API int corb_orb_device_level(const corb_orb* h, int level, int blurred, const uint8_t** d_ptr, int* pitch, int* lw,
                                   int* lh);
/* CUDA stream (cudaStream_t) the handle works on, for event timing by the caller */
CORB_API void* corb_orb_stream(const corb_orb* h);
/* number of kernel launches one extraction enqueues (graph nodes that are kernels) */
CORB_API int corb_orb_launches_per_extract(const corb_orb* h);

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

#ifdef __cplusplus
}
#endif
#endif /* CORB_B200_H */
