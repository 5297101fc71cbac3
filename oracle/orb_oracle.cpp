/* CPU ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.h for scope and parity status).
 * Build with -ffp-contract=off: the float expressions below must not be fused. */
#include "orb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

namespace {

const int8_t kPattern[1024] = {
#include "../include/corb_brief_pattern.inc"
};

const int kPatch = 31, kHalfPatch = 15, kEdge = 19; /* ORBextractor.cc:72-74 */

inline int cv_round_d(double v) { return (int)std::nearbyint(v); } /* cvRound: round-half-even */
inline int cv_round_f(float v) { return (int)std::nearbyintf(v); }

/* ---------------------------------------------------------------- resize INTER_LINEAR, u8
 * OpenCV imgproc resize.cpp (4.x): 11-bit fixed point coefficients, horizontal pass in int32,
 * vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.   Call site ORBextractor.cc:1120. */
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * dw), beta(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cv_round_f((1.f - fx) * 2048);
        alpha[2 * dx + 1] = (short)cv_round_f(fx * 2048);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = (short)cv_round_f((1.f - fy) * 2048);
        beta[2 * dy + 1] = (short)cv_round_f(fy * 2048);
    }
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uint8_t* s0 = src + (size_t)sy0 * sstride;
        const uint8_t* s1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1); /* coefficient is 0 whenever this clamps */
            row0[dx] = s0[sx] * alpha[2 * dx] + s0[sx1] * alpha[2 * dx + 1];
            row1[dx] = s1[sx] * alpha[2 * dx] + s1[sx1] * alpha[2 * dx + 1];
        }
        int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++)
            d[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

/* ---------------------------------------------------------------- GaussianBlur 7x7 sigma 2, u8
 * OpenCV 4.x bit-exact u8 path: 8.8 fixed-point taps {18,34,48,56,48,34,18}, exact accumulation over
 * both axes, one rounding (s + 2^15) >> 16, BORDER_REFLECT_101.   Call site ORBextractor.cc:1086. */
inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}
void gaussian7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    /* horizontal pass into 8.8 fixed point (max 255*256 fits u16), then vertical pass with one rounding */
    std::vector<uint16_t> hbuf((size_t)w * h);
    std::vector<uint8_t> pad(w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstride;
        for (int k = 0; k < 3; k++) { pad[k] = s[reflect101(k - 3, w)]; pad[w + 3 + k] = s[reflect101(w + k, w)]; }
        std::memcpy(pad.data() + 3, s, w);
        const uint8_t* p = pad.data();
        uint16_t* o = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; x++)
            o[x] = (uint16_t)(18 * (p[x] + p[x + 6]) + 34 * (p[x + 1] + p[x + 5]) + 48 * (p[x + 2] + p[x + 4]) + 56 * p[x + 3]);
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* r[7];
        for (int k = 0; k < 7; k++) r[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 18u * ((uint32_t)r[0][x] + r[6][x]) + 34u * ((uint32_t)r[1][x] + r[5][x]) +
                           48u * ((uint32_t)r[2][x] + r[4][x]) + 56u * (uint32_t)r[3][x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

/* ---------------------------------------------------------------- FAST-9/16
 * OpenCV features2d fast.cpp / fast_score.cpp.  Call sites ORBextractor.cc:809,814. */
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline int fast_score_px(const uint8_t* p, int stride) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = (int)p[kRingDy[k] * stride + kRingDx[k]] - v;
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = 0; /* max over the 16 nine-arcs of min(d) and of min(-d) */
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) { mn = std::min(mn, d[s + k]); mx = std::max(mx, d[s + k]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best > 0 ? best - 1 : 0;
}
/* A pixel is a corner at threshold t iff some nine-arc has all |d| > t with one sign  <=>  best >= t+1. The
 * score map stores best-1 (the response cv::FAST reports), so "corner at t" <=> best-1 >= t for t >= 1... careful:
 * best-1 >= t  <=>  best >= t+1. For t == 0 the clamp at 0 would lose the distinction; thresholds >= 1 are required. */
void fast_score(const uint8_t* src, int w, int h, int sstride, uint8_t* score, int scstride) {
    for (int y = 0; y < h; y++) std::memset(score + (size_t)y * scstride, 0, w);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++)
            score[(size_t)y * scstride + x] = (uint8_t)fast_score_px(src + (size_t)y * sstride + x, sstride);
}
struct XYR { int x, y, r; };
/* corner test at threshold th: some nine-arc entirely brighter than v+th or entirely darker than v-th */
inline bool fast_is_corner(const uint8_t* p, const int* ofs, int th) {
    const int v = p[0], hi = v + th, lo = v - th;
    /* opposite-pair rejection (any nine-arc contains one pixel of every antipodal pair) */
    int a = p[ofs[0]], b = p[ofs[8]];
    bool br = a > hi || b > hi, dk = a < lo || b < lo;
    if (!br && !dk) return false;
    a = p[ofs[4]]; b = p[ofs[12]];
    br = br && (a > hi || b > hi); dk = dk && (a < lo || b < lo);
    if (!br && !dk) return false;
    uint32_t mb = 0, md = 0;
    for (int k = 0; k < 16; k++) {
        int q = p[ofs[k]];
        mb |= (uint32_t)(q > hi) << k;
        md |= (uint32_t)(q < lo) << k;
    }
    auto run9 = [](uint32_t m) -> bool {
        m |= m << 16;              /* unroll the ring */
        m &= m >> 1;               /* runs >= 2 */
        m &= m >> 2;               /* runs >= 4 */
        m &= m >> 4;               /* runs >= 8 */
        m &= m >> 1;               /* runs >= 9 */
        return (m & 0xFFFFu) != 0;
    };
    return (br && run9(mb)) || (dk && run9(md));
}
void fast_detect(const uint8_t* src, int w, int h, int sstride, int th, std::vector<XYR>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    int ofs[16];
    for (int k = 0; k < 16; k++) ofs[k] = kRingDy[k] * sstride + kRingDx[k];
    /* masked score map: response where the pixel is a corner at th, else 0 (cv::FAST's row buffers) */
    static thread_local std::vector<uint8_t> sc;
    sc.assign((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t* row = src + (size_t)y * sstride;
        for (int x = 3; x < w - 3; x++)
            if (fast_is_corner(row + x, ofs, th)) sc[(size_t)y * w + x] = (uint8_t)fast_score_px(row + x, sstride);
    }
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *r0 = &sc[(size_t)(y - 1) * w], *r1 = &sc[(size_t)y * w], *r2 = &sc[(size_t)(y + 1) * w];
        for (int x = 3; x < w - 3; x++) {
            int s = r1[x];
            if (!s) continue;
            if (s > r0[x - 1] && s > r0[x] && s > r0[x + 1] && s > r1[x - 1] && s > r1[x + 1] && s > r2[x - 1] && s > r2[x] &&
                s > r2[x + 1])
                out.push_back({x, y, s});
        }
    }
}

/* ---------------------------------------------------------------- fastAtan2 (OpenCV mathfuncs_core, scalar atan_f32) */
float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

float ic_angle(const uint8_t* img, int stride, int x, int y, const int* umax) { /* ORBextractor.cc:77-104 */
    int m01 = 0, m10 = 0;
    const uint8_t* c = img + (size_t)y * stride + x;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int p = c[u + v * stride], m = c[u - v * stride];
            vsum += p - m;
            m10 += u * (p + m);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

void brief(const uint8_t* img, int stride, int x, int y, float angle_deg, uint8_t* desc) { /* :107-146 */
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factorPI;
    /* contract: cos/sin evaluated in double, rounded to float */
    float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const uint8_t* c = img + (size_t)y * stride + x;
    const int8_t* pat = kPattern;
    auto tap = [&](int idx) -> int {
        float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
        float fy = px * b + py * a; /* no FMA: built with -ffp-contract=off */
        float fx = px * a - py * b;
        return c[cv_round_f(fy) * stride + cv_round_f(fx)];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t0 = tap(2 * k), t1 = tap(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ---------------------------------------------------------------- quadtree distribution (:481-763) */
struct Key { int x, y, r; }; /* coordinates relative to (minBorderX, minBorderY); response = FAST score */
struct Node {
    int x0, y0, x1, y1; /* UL.x, UL.y, UR.x, BL.y */
    std::vector<Key> keys;
    bool no_more = false;
    long seq = 0; /* creation sequence number: canonical stand-in for the heap address */
    std::list<Node>::iterator self;
};
void divide(const Node& n, Node c[4]) {
    const int halfX = (int)std::ceil((float)(n.x1 - n.x0) / 2), halfY = (int)std::ceil((float)(n.y1 - n.y0) / 2);
    const int xm = n.x0 + halfX, ym = n.y0 + halfY;
    c[0] = Node{n.x0, n.y0, xm, ym, {}, false, 0, {}};
    c[1] = Node{xm, n.y0, n.x1, ym, {}, false, 0, {}};
    c[2] = Node{n.x0, ym, xm, n.y1, {}, false, 0, {}};
    c[3] = Node{xm, ym, n.x1, n.y1, {}, false, 0, {}};
    for (const Key& k : n.keys) {
        if ((float)k.x < xm) c[(float)k.y < ym ? 0 : 2].keys.push_back(k);
        else c[(float)k.y < ym ? 1 : 3].keys.push_back(k);
    }
    for (int i = 0; i < 4; i++)
        if (c[i].keys.size() == 1) c[i].no_more = true;
}
std::vector<Key> distribute(const std::vector<Key>& in, int minX, int maxX, int minY, int maxY, int N) {
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        Node n{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), maxY - minY, {}, false, seq++, {}};
        nodes.push_back(n);
        ini[i] = &nodes.back();
    }
    for (const Key& k : in) ini[(int)((float)k.x / hX)]->keys.push_back(k);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->no_more = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    bool finish = false;
    typedef std::pair<int, long> SizeSeq; /* (count, creation seq) replaces (count, pointer) */
    std::vector<std::pair<SizeSeq, Node*>> expandable;
    auto push_children = [&](Node c[4], int* nToExpand) {
        for (int i = 0; i < 4; i++) {
            if (c[i].keys.empty()) continue;
            c[i].seq = seq++;
            nodes.push_front(c[i]);
            if (c[i].keys.size() > 1) {
                if (nToExpand) ++*nToExpand;
                expandable.push_back({{(int)c[i].keys.size(), c[i].seq}, &nodes.front()});
                nodes.front().self = nodes.begin();
            }
        }
    };
    while (!finish) {
        int prevSize = (int)nodes.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->no_more) { ++it; continue; }
            Node c[4];
            divide(*it, c);
            push_children(c, &nToExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = (int)nodes.size();
                auto prev = expandable;
                expandable.clear();
                std::sort(prev.begin(), prev.end(),
                          [](const std::pair<SizeSeq, Node*>& a, const std::pair<SizeSeq, Node*>& b) { return a.first < b.first; });
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node c[4];
                    divide(*prev[j].second, c);
                    push_children(c, nullptr);
                    nodes.erase(prev[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            }
        }
    }
    std::vector<Key> out;
    out.reserve(nodes.size());
    for (const Node& n : nodes) {
        const Key* best = &n.keys[0];
        for (size_t k = 1; k < n.keys.size(); k++)
            if (n.keys[k].r > best->r) best = &n.keys[k];
        out.push_back(*best);
    }
    return out;
}

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;
};

} // namespace

struct oracle_orb {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor; /* stored as double like the reference member (ORBextractor.h:97) */
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota;
    int umax[16];
    std::vector<Image> pyr, blur;
    std::vector<std::vector<int32_t>> cand;
    std::vector<int> level_count;
};

extern "C" {

oracle_orb* oracle_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    if (nfeatures < 1 || nlevels < 1 || !(scale_factor > 1.f) || ini_th < 1 || min_th < 1) return nullptr;
    oracle_orb* o = new oracle_orb;
    o->nfeatures = nfeatures; o->nlevels = nlevels; o->ini_th = ini_th; o->min_th = min_th;
    o->scale_factor = scale_factor;
    o->scale.resize(nlevels); o->inv_scale.resize(nlevels); o->sigma2.resize(nlevels); o->inv_sigma2.resize(nlevels);
    o->scale[0] = 1.f; o->sigma2[0] = 1.f;
    for (int i = 1; i < nlevels; i++) {
        o->scale[i] = (float)(o->scale[i - 1] * o->scale_factor);
        o->sigma2[i] = o->scale[i] * o->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        o->inv_scale[i] = 1.0f / o->scale[i];
        o->inv_sigma2[i] = 1.0f / o->sigma2[i];
    }
    o->quota.resize(nlevels);
    float factor = (float)(1.0f / o->scale_factor);
    float desired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        o->quota[l] = cv_round_f(desired);
        sum += o->quota[l];
        desired *= factor;
    }
    o->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    /* umax (:452-469) */
    int v, v0, vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= vmax; ++v) o->umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }
    o->pyr.resize(nlevels); o->blur.resize(nlevels); o->cand.resize(nlevels); o->level_count.assign(nlevels, 0);
    return o;
}
void oracle_orb_destroy(oracle_orb* o) { delete o; }
int oracle_orb_levels(const oracle_orb* o) { return o->nlevels; }
void oracle_orb_tables(const oracle_orb* o, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota,
                       int* umax16) {
    for (int i = 0; i < o->nlevels; i++) {
        if (scale) scale[i] = o->scale[i];
        if (inv_scale) inv_scale[i] = o->inv_scale[i];
        if (sigma2) sigma2[i] = o->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = o->inv_sigma2[i];
        if (quota) quota[i] = o->quota[i];
    }
    if (umax16) std::memcpy(umax16, o->umax, sizeof(o->umax));
}
void oracle_orb_level_size(const oracle_orb* o, int level, int w, int h, int* lw, int* lh) {
    float s = o->inv_scale[level];
    *lw = cv_round_f((float)w * s);
    *lh = cv_round_f((float)h * s);
}
static int level_cap(const oracle_orb* o, int l, int w, int h) {
    int lw, lh;
    oracle_orb_level_size(o, l, w, h, &lw, &lh);
    int width = lw - 2 * (kEdge - 3), height = lh - 2 * (kEdge - 3);
    if (width < 1 || height < 1) return -1;
    int nIni = (int)std::round((float)width / height);
    if (nIni < 1) return -1;
    return std::max(o->quota[l] + 3, 4 * nIni);
}
int oracle_orb_capacity(const oracle_orb* o, int w, int h) {
    int cap = 0;
    for (int l = 0; l < o->nlevels; l++) {
        int c = level_cap(o, l, w, h);
        if (c < 0) return -1;
        cap += c;
    }
    return cap;
}

int oracle_orb_extract(oracle_orb* o, const uint8_t* img, int w, int h, int stride, oracle_keypoint* kps, uint8_t* desc) {
    if (!img || w < 1 || h < 1) return 0; /* empty image: silent return (:1046) */
    if (oracle_orb_capacity(o, w, h) < 0) return -2;
    const int L = o->nlevels;
    /* ComputePyramid (:1107-1132); the 19 px border is never read by anything below */
    for (int l = 0; l < L; l++) {
        Image& im = o->pyr[l];
        oracle_orb_level_size(o, l, w, h, &im.w, &im.h);
        im.px.resize((size_t)im.w * im.h);
        if (l == 0)
            for (int y = 0; y < h; y++) std::memcpy(&im.px[(size_t)y * w], img + (size_t)y * stride, w);
        else
            resize_linear_u8(o->pyr[l - 1].px.data(), o->pyr[l - 1].w, o->pyr[l - 1].h, o->pyr[l - 1].w, im.px.data(), im.w,
                             im.h, im.w);
    }
    /* ComputeKeyPointsOctTree (:765-853) */
    std::vector<std::vector<Key>> all(L);
    const float W = 30;
    std::vector<XYR> cell;
    for (int l = 0; l < L; l++) {
        const Image& im = o->pyr[l];
        const int minBX = kEdge - 3, minBY = minBX, maxBX = im.w - kEdge + 3, maxBY = im.h - kEdge + 3;
        std::vector<Key> cand;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) return -2;
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                const uint8_t* roi = im.px.data() + (size_t)y0 * im.w + x0;
                fast_detect(roi, cw, ch, im.w, o->ini_th, cell);
                if (cell.empty()) fast_detect(roi, cw, ch, im.w, o->min_th, cell);
                for (const XYR& k : cell) cand.push_back({k.x + j * wCell, k.y + i * hCell, k.r});
            }
        }
        o->cand[l].clear();
        for (const Key& k : cand) { o->cand[l].push_back(k.x); o->cand[l].push_back(k.y); o->cand[l].push_back(k.r); }
        all[l] = distribute(cand, minBX, maxBX, minBY, maxBY, o->quota[l]);
        for (Key& k : all[l]) { k.x += minBX; k.y += minBY; }
        o->level_count[l] = (int)all[l].size();
    }
    /* orientation on the un-blurred level, then blur + descriptors per level, concatenated (:1067-1104) */
    int n = 0;
    for (int l = 0; l < L; l++) {
        const Image& im = o->pyr[l];
        Image& bl = o->blur[l];
        bl.w = bl.h = 0; bl.px.clear();
        if (all[l].empty()) continue;
        bl.w = im.w; bl.h = im.h; bl.px.resize(im.px.size());
        gaussian7_u8(im.px.data(), im.w, im.h, im.w, bl.px.data(), im.w);
        const int scaledPatch = (int)(kPatch * o->scale[l]);
        for (const Key& k : all[l]) {
            oracle_keypoint& kp = kps[n];
            kp.angle = ic_angle(im.px.data(), im.w, k.x, k.y, o->umax);
            brief(bl.px.data(), bl.w, k.x, k.y, kp.angle, desc + (size_t)n * 32);
            kp.x = (float)k.x; kp.y = (float)k.y;
            if (l != 0) { kp.x *= o->scale[l]; kp.y *= o->scale[l]; }
            kp.size = (float)scaledPatch;
            kp.response = (float)k.r;
            kp.octave = l;
            kp.class_id = -1;
            n++;
        }
    }
    return n;
}

const uint8_t* oracle_orb_pyramid(const oracle_orb* o, int level, int* w, int* h) {
    *w = o->pyr[level].w; *h = o->pyr[level].h;
    return o->pyr[level].px.data();
}
const uint8_t* oracle_orb_blurred(const oracle_orb* o, int level, int* w, int* h) {
    *w = o->blur[level].w; *h = o->blur[level].h;
    return o->blur[level].px.empty() ? nullptr : o->blur[level].px.data();
}
int oracle_orb_candidates(const oracle_orb* o, int level, const int32_t** xyr) {
    *xyr = o->cand[level].data();
    return (int)(o->cand[level].size() / 3);
}
int oracle_orb_level_count(const oracle_orb* o, int level) { return o->level_count[level]; }

/* Frame::ComputeStereoMatches (corbslam_client/src/Frame.cc:470-644) on the results of the last extraction of the
 * left and right oracle extractors (keypoints/descriptors passed in, un-blurred pyramids taken from the handles).
 * u_right / depth: n_left floats, -1 where no match. */
int oracle_stereo_matches(const oracle_orb* L, const oracle_orb* R, const oracle_keypoint* kl, const uint8_t* dl, int nl,
                          const oracle_keypoint* kr, const uint8_t* dr, int nr, float mbf, float mb, float* u_right, float* depth) {
    const int TH_HIGH = 100, TH_LOW = 50;
    for (int i = 0; i < nl; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = L->pyr[0].h;
    std::vector<std::vector<size_t>> rows(nRows);
    for (int iR = 0; iR < nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * R->scale[kr[iR].octave];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) rows[yi].push_back(iR); /* the reference indexes unchecked (always in range for ORB keypoints) */
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < nl; iL++) {
        const int levelL = kl[iL].octave;
        const float vL = kl[iL].y, uL = kl[iL].x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<size_t>& cand = rows[row];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        for (size_t iC = 0; iC < cand.size(); iC++) {
            const size_t iR = cand[iC];
            if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
            const float uR = kr[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int32_t* pa = (const int32_t*)(dl + (size_t)iL * 32);
                const int32_t* pb = (const int32_t*)(dr + iR * 32);
                int dist = 0;
                for (int k = 0; k < 8; k++) dist += __builtin_popcount((unsigned)(pa[k] ^ pb[k]));
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = L->inv_scale[levelL];
            const float scaleduL = std::round(uL * scaleFactor), scaledvL = std::round(vL * scaleFactor);
            const float scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5, Ls = 5;
            const Image& imL = L->pyr[levelL];
            const Image& imR = R->pyr[levelL];
            const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
            int best = 2147483647, bestincR = 0;
            float vDists[11];
            const float iniu = scaleduR0 + Ls - w, endu = scaleduR0 + Ls + w + 1;
            if (iniu < 0 || endu >= imR.w) continue;
            const float cL = imL.px[(size_t)cv * imL.w + cu];
            for (int incR = -Ls; incR <= Ls; incR++) {
                const float cR = imR.px[(size_t)cv * imR.w + cr + incR];
                double acc = 0; /* cv::norm(NORM_L1) of CV_32F accumulates in double */
                for (int dy = -w; dy <= w; dy++)
                    for (int dx = -w; dx <= w; dx++) {
                        const float a = (float)imL.px[(size_t)(cv + dy) * imL.w + cu + dx] - cL;
                        const float b = (float)imR.px[(size_t)(cv + dy) * imR.w + cr + incR + dx] - cR;
                        acc += std::fabs(a - b);
                    }
                const float dist = (float)acc;
                if (dist < best) { best = (int)dist; bestincR = incR; }
                vDists[Ls + incR] = dist;
            }
            if (bestincR == -Ls || bestincR == Ls) continue;
            const float dist1 = vDists[Ls + bestincR - 1], dist2 = vDists[Ls + bestincR], dist3 = vDists[Ls + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = L->scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = mbf / disparity;
                u_right[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(best, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0; /* the reference reads vDistIdx[0] of an empty vector here (undefined) */
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        u_right[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
        kept--;
    }
    return kept;
}

void oracle_resize_linear_u8(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds) {
    resize_linear_u8(s, sw, sh, ss, d, dw, dh, ds);
}
void oracle_gaussian7_u8(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds) { gaussian7_u8(s, w, h, ss, d, ds); }
void oracle_fast_score(const uint8_t* s, int w, int h, int ss, uint8_t* sc, int scs) { fast_score(s, w, h, ss, sc, scs); }
int oracle_fast_detect(const uint8_t* s, int w, int h, int ss, int th, int32_t* out, int cap) {
    std::vector<XYR> v;
    fast_detect(s, w, h, ss, th, v);
    int n = (int)std::min<size_t>(v.size(), (size_t)cap);
    for (int i = 0; i < n; i++) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].r; }
    return (int)v.size();
}
float oracle_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int oracle_cv_round_f(float v) { return cv_round_f(v); }
float oracle_ic_angle(const uint8_t* img, int stride, int x, int y, const int* umax16) { return ic_angle(img, stride, x, y, umax16); }
void oracle_brief(const uint8_t* img, int stride, int x, int y, float angle_deg, uint8_t* d) { brief(img, stride, x, y, angle_deg, d); }

} // extern "C"
