// Drop-in replacement of corbslam_client/src/ORBextractor.cc: the same class (include/ORBextractor.h, untouched), the same
// constructor / operator() / getters / mvImagePyramid, computed by libcorb_b200.so on the GPU.
//   ORBextractor::ORBextractor   ORBextractor.cc:410-470  -> corb_orb_create + corb_orb_tables
//   ORBextractor::operator()     ORBextractor.cc:1043-1105 -> corb_orb_extract (keypoints, descriptors, pyramid levels)
// The header has no slot for the handle, so it lives in a file-static map keyed by the object (two extractors - left and
// right - are used from two threads, Frame.cc:78-81: the map is locked, the handles are independent).
#include <string.h>

#include <map>
#include <mutex>

#include "ORBextractor.h"
#include "shim_common.h"

namespace ORB_SLAM2 {

namespace {
std::mutex g_mu;
std::map<const ORBextractor*, corb_orb*> g_handles;
corb_orb* handle_of(const ORBextractor* e) {
    std::lock_guard<std::mutex> lock(g_mu);
    return g_handles[e];
}
}  // namespace

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    corb_orb* h = nullptr;
    corb_shim::check(corb_orb_create(_nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, corb_shim::device(), &h), "corb_orb_create");
    {
        std::lock_guard<std::mutex> lock(g_mu);
        g_handles[this] = h;  // (~ORBextractor is inline and empty in the header: handles live as long as the process)
    }
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    umax.resize(16);
    corb_shim::check(corb_orb_tables(h, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                                     mnFeaturesPerLevel.data(), umax.data()), "corb_orb_tables");
    mvImagePyramid.resize(nlevels);
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;  // :1046
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);
    corb_orb* h = handle_of(this);
    const int cap = corb_orb_capacity(h, image.cols, image.rows);
    if (cap < 0) corb_shim::check(CORB_ERR_INVALID, "corb_orb_capacity");
    std::vector<uint8_t*> pyr(nlevels);
    for (int l = 0; l < nlevels; ++l) {  // Frame::ComputeStereoMatches reads mvImagePyramid (Frame.cc:477,567,584)
        int w = 0, hgt = 0;
        corb_shim::check(corb_orb_level_size(h, l, image.cols, image.rows, &w, &hgt), "corb_orb_level_size");
        // the reference keeps every level inside a buffer with a 19 px REFLECT_101 frame (:1113-1127) and the stereo matcher's
        // 11 x 11 windows may reach into it: same layout here, the frame is filled from the level on the host
        cv::Mat temp(hgt + 38, w + 38, CV_8UC1);
        mvImagePyramid[l] = temp(cv::Rect(19, 19, w, hgt));
        pyr[l] = 0;
    }
    // dense level buffers for the call, copied into the framed ones (a level is at most 466 KB)
    std::vector<std::vector<uint8_t> > dense(nlevels);
    for (int l = 0; l < nlevels; ++l) {
        dense[l].resize((size_t)mvImagePyramid[l].cols * mvImagePyramid[l].rows);
        pyr[l] = dense[l].data();
    }
    std::vector<corb_keypoint> k((size_t)cap);
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    corb_shim::check(corb_orb_extract(h, image.data, image.cols, image.rows, (int)image.step, k.data(), desc.data, &n, pyr.data()),
                     "corb_orb_extract");
    for (int l = 0; l < nlevels; ++l) {
        cv::Mat& lv = mvImagePyramid[l];
        for (int y = 0; y < lv.rows; ++y) memcpy(lv.ptr(y), dense[l].data() + (size_t)y * lv.cols, lv.cols);
        const int W = lv.cols, H = lv.rows;
        uint8_t* base = lv.data - 19 * lv.step - 19;  // the framed buffer
        for (int y = -19; y < H + 19; ++y) {
            int sy = y < 0 ? -y : (y >= H ? 2 * (H - 1) - y : y);
            uint8_t* row = base + (size_t)(y + 19) * lv.step;
            const uint8_t* src = lv.data + (size_t)sy * lv.step;
            if (y < 0 || y >= H) memcpy(row + 19, src, W);
            for (int x = 0; x < 19; ++x) {
                row[18 - x] = src[x + 1 < W ? x + 1 : W - 1];
                row[19 + W + x] = src[W - 2 - x >= 0 ? W - 2 - x : 0];
            }
        }
    }
    static_assert(sizeof(cv::KeyPoint) == sizeof(corb_keypoint), "cv::KeyPoint layout");
    _keypoints.resize(n);
    if (n) memcpy(_keypoints.data(), k.data(), (size_t)n * sizeof(corb_keypoint));
    if (n == 0) {
        _descriptors.release();  // :1064-1065
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat out = _descriptors.getMat();
        for (int i = 0; i < n; ++i) memcpy(out.ptr(i), desc.ptr(i), 32);
    }
}

}  // namespace ORB_SLAM2
