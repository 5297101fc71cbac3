/* C ABI over the reference's own classes — TEST INFRASTRUCTURE ONLY (oracle/refbuild). Step 1: ORBextractor. */
#include <stdint.h>

#include "ORBextractor.h"

using namespace ORB_SLAM2;

extern "C" {

struct ref_keypoint { float x, y, size, angle, response; int32_t octave, class_id; };

void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    return new ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_orb_destroy(void* h) { delete (ORBextractor*)h; }
int ref_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, ref_keypoint* kps, uint8_t* desc, int cap) {
    ORBextractor& ex = *(ORBextractor*)h;
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    ex(image, cv::Mat(), keys, descriptors);
    const int n = (int)keys.size();
    if (n > cap) return -n;
    static_assert(sizeof(cv::KeyPoint) == sizeof(ref_keypoint), "cv::KeyPoint layout");
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

} // extern "C"
