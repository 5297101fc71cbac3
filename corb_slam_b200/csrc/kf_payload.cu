// Binary KeyFrame payload: the bulk of what a client ships to the server per keyframe is the
// extractor's output - mvKeys, mvKeysUn, mvuRight, mvDepth, mDescriptors, mBowVec, mFeatVec - which KeyFrame::serialize
// (corbslam_client/include/KeyFrame.h:61-87, via SerializeObject.h:34-61) writes through boost::archive::text_oarchive as
// decimal text (DataDriver.cc:40-238): >= 236 characters per keypoint. This is the same content as one little-endian blob behind
// a magic + version tag, exact to the bit (floats are not re-parsed), ~76 bytes per keypoint; the ROS srv/msg surface is
// untouched (the blob travels inside the existing `DATA` string). Host code: nothing here needs the GPU.
#include <string.h>

#include "common.cuh"

namespace {

const uint8_t kMagic[4] = {'C', 'K', 'F', 1};
enum { kFlagSameUn = 1, kFlagNoClassId = 2 };

struct Header {  // 40 bytes
    uint8_t magic[4];
    uint16_t version, flags;
    int32_t n, n_bow, n_fv, n_fv_idx;
    uint32_t body_bytes, crc;
    uint32_t reserved[2];
};
static_assert(sizeof(Header) == 40, "payload header layout");

uint32_t crc32(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

size_t keys_bytes(int n, bool class_id) { return (size_t)n * (5 * 4 + 1 + (class_id ? 4 : 0)); }

uint8_t* put_keys(uint8_t* o, const corb_keypoint* k, int n, bool class_id) {
    float* f = reinterpret_cast<float*>(o);  // SoA: x[n] y[n] size[n] angle[n] response[n] octave[n] (class_id[n])
    for (int i = 0; i < n; i++) { f[i] = k[i].x; f[n + i] = k[i].y; f[2 * n + i] = k[i].size; f[3 * n + i] = k[i].angle; f[4 * n + i] = k[i].response; }
    o += (size_t)n * 20;
    for (int i = 0; i < n; i++) o[i] = (uint8_t)(int8_t)k[i].octave;
    o += n;
    if (class_id) {
        for (int i = 0; i < n; i++) memcpy(o + 4 * (size_t)i, &k[i].class_id, 4);
        o += 4 * (size_t)n;
    }
    return o;
}
const uint8_t* get_keys(const uint8_t* p, corb_keypoint* k, int n, bool class_id) {
    for (int i = 0; i < n; i++) {
        float v[5];
        for (int c = 0; c < 5; c++) memcpy(&v[c], p + 4 * ((size_t)c * n + i), 4);
        k[i].x = v[0]; k[i].y = v[1]; k[i].size = v[2]; k[i].angle = v[3]; k[i].response = v[4];
    }
    p += (size_t)n * 20;
    for (int i = 0; i < n; i++) k[i].octave = (int8_t)p[i];
    p += n;
    if (class_id) {
        for (int i = 0; i < n; i++) memcpy(&k[i].class_id, p + 4 * (size_t)i, 4);
        p += 4 * (size_t)n;
    } else {
        for (int i = 0; i < n; i++) k[i].class_id = -1;
    }
    return p;
}

}  // namespace

extern "C" {

size_t corb_kf_payload_bound(int n, int n_bow, int n_fv, int n_fv_idx) {
    return sizeof(Header) + 2 * keys_bytes(n, true) + (size_t)n * (4 + 4 + 32) + (size_t)n_bow * 12 + (size_t)n_fv * 4 + ((size_t)n_fv + 1) * 4 +
           (size_t)n_fv_idx * 4;
}

int corb_kf_payload_encode(const corb_keypoint* keys, const corb_keypoint* keys_un, const float* u_right, const float* depth,
                           const uint8_t* desc, int n, const uint32_t* bow_words, const double* bow_vals, int n_bow,
                           const uint32_t* fv_nodes, const int32_t* fv_off, const uint32_t* fv_idx, int n_fv, uint8_t* out, size_t cap,
                           size_t* written) {
    CORB_CHECK(n >= 0 && n_bow >= 0 && n_fv >= 0 && out && written, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(n == 0 || (keys && u_right && depth && desc), CORB_ERR_INVALID, "keypoint arrays are NULL");
    CORB_CHECK(n_bow == 0 || (bow_words && bow_vals), CORB_ERR_INVALID, "BowVector arrays are NULL");
    CORB_CHECK(n_fv == 0 || (fv_nodes && fv_off && fv_idx), CORB_ERR_INVALID, "FeatureVector arrays are NULL");
    const int n_fv_idx = n_fv ? fv_off[n_fv] : 0;
    CORB_CHECK(n_fv_idx >= 0, CORB_ERR_INVALID, "FeatureVector offsets are not ascending");
    CORB_CHECK(cap >= corb_kf_payload_bound(n, n_bow, n_fv, n_fv_idx), CORB_ERR_CAPACITY, "output buffer too small");
    bool same_un = !keys_un || keys_un == keys || (n && memcmp(keys, keys_un, sizeof(corb_keypoint) * (size_t)n) == 0) || n == 0;
    bool class_id = false;
    for (int i = 0; i < n && !class_id; i++) {
        class_id = keys[i].class_id != -1 || (!same_un && keys_un[i].class_id != -1);
        CORB_CHECK(keys[i].octave >= -128 && keys[i].octave <= 127, CORB_ERR_INVALID, "octave %d does not fit the payload", keys[i].octave);
    }
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kMagic, 4);
    h.version = 1;
    h.flags = (uint16_t)((same_un ? kFlagSameUn : 0) | (class_id ? 0 : kFlagNoClassId));
    h.n = n; h.n_bow = n_bow; h.n_fv = n_fv; h.n_fv_idx = n_fv_idx;
    uint8_t* body = out + sizeof(Header);
    uint8_t* o = put_keys(body, keys, n, class_id);
    if (!same_un) o = put_keys(o, keys_un, n, class_id);
    memcpy(o, u_right, 4 * (size_t)n); o += 4 * (size_t)n;
    memcpy(o, depth, 4 * (size_t)n); o += 4 * (size_t)n;
    memcpy(o, desc, 32 * (size_t)n); o += 32 * (size_t)n;
    memcpy(o, bow_words, 4 * (size_t)n_bow); o += 4 * (size_t)n_bow;
    memcpy(o, bow_vals, 8 * (size_t)n_bow); o += 8 * (size_t)n_bow;
    memcpy(o, fv_nodes, 4 * (size_t)n_fv); o += 4 * (size_t)n_fv;
    if (n_fv) { memcpy(o, fv_off, 4 * ((size_t)n_fv + 1)); o += 4 * ((size_t)n_fv + 1); }
    memcpy(o, fv_idx, 4 * (size_t)n_fv_idx); o += 4 * (size_t)n_fv_idx;
    h.body_bytes = (uint32_t)(o - body);
    h.crc = crc32(body, h.body_bytes);
    memcpy(out, &h, sizeof(h));
    *written = sizeof(Header) + h.body_bytes;
    return CORB_OK;
}

int corb_kf_payload_info(const uint8_t* buf, size_t len, int* n, int* n_bow, int* n_fv, int* n_fv_idx, int* same_un) {
    CORB_CHECK(buf && len >= sizeof(Header), CORB_ERR_INVALID, "payload shorter than its header");
    Header h;
    memcpy(&h, buf, sizeof(h));
    CORB_CHECK(memcmp(h.magic, kMagic, 4) == 0, CORB_ERR_INVALID, "not a CORB keyframe payload (a boost text archive starts with '22 serialization::archive')");
    CORB_CHECK(h.version == 1, CORB_ERR_UNSUPPORTED, "payload version %d", (int)h.version);
    CORB_CHECK(h.n >= 0 && h.n_bow >= 0 && h.n_fv >= 0 && h.n_fv_idx >= 0 && sizeof(Header) + (size_t)h.body_bytes <= len, CORB_ERR_INVALID,
               "payload truncated");
    if (n) *n = h.n;
    if (n_bow) *n_bow = h.n_bow;
    if (n_fv) *n_fv = h.n_fv;
    if (n_fv_idx) *n_fv_idx = h.n_fv_idx;
    if (same_un) *same_un = (h.flags & kFlagSameUn) != 0;
    return CORB_OK;
}

int corb_kf_payload_decode(const uint8_t* buf, size_t len, corb_keypoint* keys, corb_keypoint* keys_un, float* u_right, float* depth,
                           uint8_t* desc, uint32_t* bow_words, double* bow_vals, uint32_t* fv_nodes, int32_t* fv_off, uint32_t* fv_idx) {
    int n, n_bow, n_fv, n_fv_idx, same;
    int rc = corb_kf_payload_info(buf, len, &n, &n_bow, &n_fv, &n_fv_idx, &same);
    if (rc != CORB_OK) return rc;
    Header h;
    memcpy(&h, buf, sizeof(h));
    const uint8_t* body = buf + sizeof(Header);
    CORB_CHECK(crc32(body, h.body_bytes) == h.crc, CORB_ERR_INVALID, "payload checksum mismatch");
    const bool class_id = !(h.flags & kFlagNoClassId);
    const size_t expect = keys_bytes(n, class_id) * (same ? 1 : 2) + (size_t)n * 40 + (size_t)n_bow * 12 + (size_t)n_fv * 4 +
                          (n_fv ? ((size_t)n_fv + 1) * 4 : 0) + (size_t)n_fv_idx * 4;
    CORB_CHECK(expect == h.body_bytes, CORB_ERR_INVALID, "payload body size does not match its header");
    CORB_CHECK(n == 0 || (keys && u_right && depth && desc), CORB_ERR_INVALID, "output arrays are NULL");
    const uint8_t* p = get_keys(body, keys, n, class_id);
    if (!same) {
        CORB_CHECK(keys_un, CORB_ERR_INVALID, "the payload carries separate undistorted keypoints");
        p = get_keys(p, keys_un, n, class_id);
    } else if (keys_un && n) {
        memcpy(keys_un, keys, sizeof(corb_keypoint) * (size_t)n);
    }
    memcpy(u_right, p, 4 * (size_t)n); p += 4 * (size_t)n;
    memcpy(depth, p, 4 * (size_t)n); p += 4 * (size_t)n;
    memcpy(desc, p, 32 * (size_t)n); p += 32 * (size_t)n;
    if (n_bow) { memcpy(bow_words, p, 4 * (size_t)n_bow); p += 4 * (size_t)n_bow; memcpy(bow_vals, p, 8 * (size_t)n_bow); p += 8 * (size_t)n_bow; }
    if (n_fv) {
        memcpy(fv_nodes, p, 4 * (size_t)n_fv); p += 4 * (size_t)n_fv;
        memcpy(fv_off, p, 4 * ((size_t)n_fv + 1)); p += 4 * ((size_t)n_fv + 1);
        CORB_CHECK(fv_off[0] == 0 && fv_off[n_fv] == n_fv_idx, CORB_ERR_INVALID, "FeatureVector offsets are inconsistent");
        memcpy(fv_idx, p, 4 * (size_t)n_fv_idx);
    }
    return CORB_OK;
}

}  // extern "C"
