// ORB front-end kernels for sm_100a: pyramid resize, per-cell FAST-9/16 + NMS + threshold fallback, 7x7 Gaussian,
// quadtree keypoint distribution, intensity-centroid orientation and steered BRIEF.
//
// Semantics follow corbslam_client/src/ORBextractor.cc (reference) and the OpenCV primitives it calls, as pinned by
// the CPU oracle (oracle/orb_oracle.cpp); every kernel is integer/bit exact against it. The design is not a port:
// the per-cell cv::FAST calls become one launch over all cells of all levels, the std::list quadtree becomes a
// level-synchronous array algorithm in one CTA per pyramid level, and descriptors are one warp per keypoint.
#include "common.cuh"
#include "orb_kernels.cuh"

#include <cuda.h>

namespace corb {

// Debug build only (make EXTRA=-DCORB_TIMELINE): every block stamps %globaltimer on entry/exit into a per-kernel
// (min start, max end) table, so tools/timeline.py can draw the concurrency of the per-frame graph without nsys.
#ifdef CORB_TIMELINE
__device__ unsigned long long g_tl[2][64];
struct TlScope {
    int id;
    __device__ __forceinline__ static unsigned long long now() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        return t;
    }
    __device__ __forceinline__ TlScope(int i) : id(i) {
        if (threadIdx.x == 0 && threadIdx.y == 0) atomicMin(&g_tl[0][id], now());
    }
    __device__ __forceinline__ ~TlScope() {
        if (threadIdx.x == 0 && threadIdx.y == 0) atomicMax(&g_tl[1][id], now());
    }
};
#define TL_SCOPE(id) TlScope tl_scope_(id)
}  // namespace corb
extern "C" __attribute__((visibility("default"))) int corb_debug_timeline(unsigned long long* out128, int reset) {
    if (out128 && cudaMemcpyFromSymbol(out128, corb::g_tl, sizeof(corb::g_tl)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long init[2][64];
        for (int i = 0; i < 64; i++) { init[0][i] = ~0ull; init[1][i] = 0; }
        if (cudaMemcpyToSymbol(corb::g_tl, init, sizeof(init)) != cudaSuccess) return 1;
    }
    return 0;
}
namespace corb {
#else
#define TL_SCOPE(id) do { } while (0)
#endif

#ifdef CORB_OCT_TRACE
#define TR_DECL long long tr_[16]; int trn_ = 0
#define TR() do { if (threadIdx.x == 0 && threadIdx.y == 0 && trn_ < 16) tr_[trn_++] = clock64(); } while (0)
#define TR_PRINT(name, cond) do { if (threadIdx.x == 0 && threadIdx.y == 0 && blockIdx.z == 0 && (cond)) { \
        for (int i_ = trn_; i_ < 16; i_++) tr_[i_] = tr_[trn_ - 1]; \
        printf("%s blk %d n=%d: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld total %lld\n", name, (int)blockIdx.x, trn_, \
               tr_[1] - tr_[0], tr_[2] - tr_[1], tr_[3] - tr_[2], tr_[4] - tr_[3], tr_[5] - tr_[4], tr_[6] - tr_[5], tr_[7] - tr_[6], \
               tr_[8] - tr_[7], tr_[9] - tr_[8], tr_[10] - tr_[9], tr_[11] - tr_[10], tr_[12] - tr_[11], tr_[trn_ - 1] - tr_[0]); } } while (0)
#else
#define TR_DECL do { } while (0)
#define TR() do { } while (0)
#define TR_PRINT(name, cond) do { } while (0)
#endif

__device__ __align__(16) const int8_t d_pattern[1024] = {
#include "../../include/corb_brief_pattern.inc"
};
// umax of ORBextractor.cc:452-469 for HALF_PATCH_SIZE = 15 (the host recomputes it and checks equality)
__device__ const int d_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// Programmatic dependent launch: a kernel launched with this attribute may be scheduled while its predecessor in the
// stream is still draining; it blocks in pdl_wait() (griddepcontrol.wait) until the predecessor has completed and its
// writes are visible. Along the serial chains of small kernels (resize_1 .. resize_7, FAST_l -> quadtree_l) this takes
// the ~1.3 us launch gap per edge off the critical path. CORB_PDL=0 launches everything the plain way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the dependent launch be scheduled once every CTA of this grid has got here (it still waits in pdl_wait() for this
// grid to finish); without it the dependent is only released when the last CTA exits.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static bool pdl_enabled() {
    static const bool on = !(getenv("CORB_PDL") && atoi(getenv("CORB_PDL")) == 0);
    return on;
}
template <typename... KArgs, typename... Args>
static void launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------ K0 import
// Copies the caller's image (device memory, or page-locked host memory read over PCIe through its UVA mapping) into
// level 0 of the pitched pyramid. First node of the per-frame graph, so one cudaGraphLaunch is the only driver call
// an extraction needs; its (src, stride) arguments are patched per launch with cudaGraphExecKernelNodeSetParams.
// Row-wise variant for device-resident sources (4 bytes per thread, word stores into the pitched level).
__global__ void __launch_bounds__(256) k_import_rows(const uint8_t* __restrict__ src, int stride, uint8_t* __restrict__ dst, int pitch,
                                                     int w, int h, const uint8_t* __restrict__ src1, int stride1,
                                                     uint8_t* __restrict__ dst1) {
    TL_SCOPE(0);
    pdl_release();
    if (blockIdx.z) { src = src1; stride = stride1; dst = dst1; }  // second image of a stereo pair
    const int y = blockIdx.y;
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= w) return;
    const uint8_t* s = src + (size_t)y * stride + x0;
    uint32_t v;
    if ((((uintptr_t)src | (uintptr_t)stride) & 3) == 0 && x0 + 3 < w) {
        v = *reinterpret_cast<const uint32_t*>(s);
    } else {
        v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (x0 + k < w) v |= (uint32_t)s[k] << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * pitch + x0) = v;  // pitch % 128 == 0 and pitch >= w rounded up to 4
}

// Flat variant for host sources (mapped page-locked memory read over PCIe).
__global__ void __launch_bounds__(256) k_import(const uint8_t* __restrict__ src, int stride, uint8_t* __restrict__ dst, int pitch,
                                                int w, int h, const uint8_t* __restrict__ src1, int stride1,
                                                uint8_t* __restrict__ dst1) {
    TL_SCOPE(0);
    // (no pdl_release here: this kernel lasts as long as the PCIe read, and dependents parked on the SMs for that long
    //  take slots from the other image's pipeline - measured 122 -> 127 us per frame)
    if (blockIdx.z) { src = src1; stride = stride1; dst = dst1; }  // second image of a stereo pair
    // The image is walked as a flat run of w * h bytes, 16 per thread: a contiguous, 16-byte aligned source (the usual
    // case: a cv::Mat or a pinned staging buffer) is read with one 128-bit load per thread, i.e. 512-byte requests per
    // warp on the PCIe link when the source is mapped host memory. (Image rows themselves are rarely aligned: 1242 % 4 = 2.)
    const long long n = (long long)w * h;
    const long long off = ((long long)blockIdx.x * 256 + threadIdx.x) * 16;
    if (off >= n) return;
    int row = (int)(off / w), col = (int)(off - (long long)row * w);
    uint32_t v[4];
    if (stride == w && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && off + 16 <= n) {
        const uint4 q = *reinterpret_cast<const uint4*>(src + off);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        int r = row, c = col;
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (off + k < n) v[k >> 2] |= (uint32_t)src[(size_t)r * stride + c] << (8 * (k & 3));
            if (++c >= w) { c = 0; r++; }
        }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (off + k < n) dst[(size_t)row * pitch + col] = (uint8_t)(v[k >> 2] >> (8 * (k & 3)));
        if (++col >= w) { col = 0; row++; }
    }
}

static dim3 import_grid(int w, int h, int n_img, bool host_src) {
    return host_src ? dim3((unsigned)(((long long)w * h + 16 * 256 - 1) / (16 * 256)), 1, n_img) : dim3((w + 1023) / 1024, h, n_img);
}

void launch_import(const OrbGeom& g, const OrbBuffers& b, const uint8_t* src, int stride, cudaStream_t s, const OrbBuffers* b1,
                   const uint8_t* src1, int stride1, bool host_src) {
    const LevelGeom& L0 = g.lv[0];
    const dim3 grid = import_grid(L0.w, L0.h, b1 ? 2 : 1, host_src);
    if (host_src)
        k_import<<<grid, 256, 0, s>>>(src, stride, b.pyr + L0.img_off, L0.pitch, L0.w, L0.h, src1, stride1, b1 ? b1->pyr + L0.img_off : nullptr);
    else
        k_import_rows<<<grid, 256, 0, s>>>(src, stride, b.pyr + L0.img_off, L0.pitch, L0.w, L0.h, src1, stride1,
                                           b1 ? b1->pyr + L0.img_off : nullptr);
}
void import_launch_dims(int w, int h, int n_img, bool host_src, dim3* grid, dim3* block) {
    *grid = import_grid(w, h, n_img, host_src);
    *block = dim3(256);
}
const void* import_kernel_ptr(bool host_src) { return host_src ? (const void*)k_import : (const void*)k_import_rows; }

// ------------------------------------------------------------------------------------------------ K1 resize
// cv::resize INTER_LINEAR u8 (ORBextractor.cc:1120). Coefficient tables are built on the host exactly as OpenCV
// builds them (double -> float -> short), so the kernel is pure integer arithmetic.
__global__ void __launch_bounds__(256) k_resize(LevelGeom src, LevelGeom dst, const uint8_t* __restrict__ pyr_src,
                                                uint8_t* __restrict__ pyr_dst, const int* __restrict__ xofs,
                                                const short2* __restrict__ alpha, const int* __restrict__ yofs,
                                                const short2* __restrict__ beta, const uint8_t* __restrict__ pyr_src1,
                                                uint8_t* __restrict__ pyr_dst1) {
    TL_SCOPE(dst.level);
    if (blockIdx.z) { pyr_src = pyr_src1; pyr_dst = pyr_dst1; }
    pdl_release();
    pdl_wait();  // level l - 1 comes from the previous launch of the chain
    // a thread produces 4 pixels of two consecutive rows: the x tables (source column, coefficients) are fetched once for both
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int dy = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (dy >= dst.h || dx0 >= dst.w) return;
    const bool two = dy + 1 < dst.h;
    const int sya = yofs[dy], syb = yofs[two ? dy + 1 : dy];
    const short2 ba = beta[dy], bb = beta[two ? dy + 1 : dy];
    const uint8_t* a0p = pyr_src + (size_t)min(max(sya, 0), src.h - 1) * src.pitch;
    const uint8_t* a1p = pyr_src + (size_t)min(max(sya + 1, 0), src.h - 1) * src.pitch;
    const uint8_t* b0p = pyr_src + (size_t)min(max(syb, 0), src.h - 1) * src.pitch;
    const uint8_t* b1p = pyr_src + (size_t)min(max(syb + 1, 0), src.h - 1) * src.pitch;
    uint32_t pa = 0, pb = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int dx = dx0 + k;
        if (dx < dst.w) {
            const int sx = xofs[dx];
            const int sx1 = min(sx + 1, src.w - 1);
            const short2 a = alpha[dx];
            {
                const int r0 = (int)a0p[sx] * a.x + (int)a0p[sx1] * a.y;
                const int r1 = (int)a1p[sx] * a.x + (int)a1p[sx1] * a.y;
                const int v = ((((int)ba.x * (r0 >> 4)) >> 16) + (((int)ba.y * (r1 >> 4)) >> 16) + 2) >> 2;
                pa |= (uint32_t)(v & 0xff) << (8 * k);
            }
            {
                const int r0 = (int)b0p[sx] * a.x + (int)b0p[sx1] * a.y;
                const int r1 = (int)b1p[sx] * a.x + (int)b1p[sx1] * a.y;
                const int v = ((((int)bb.x * (r0 >> 4)) >> 16) + (((int)bb.y * (r1 >> 4)) >> 16) + 2) >> 2;
                pb |= (uint32_t)(v & 0xff) << (8 * k);
            }
        }
    }
    uint8_t* o = pyr_dst + (size_t)dy * dst.pitch + dx0;  // pitch % 128 == 0, dx0 % 4 == 0
    *reinterpret_cast<uint32_t*>(o) = pa;
    if (two) *reinterpret_cast<uint32_t*>(o + dst.pitch) = pb;
}

void launch_resize(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s, const OrbBuffers* b1) {
    const LevelGeom& src = g.lv[level - 1];
    const LevelGeom& dst = g.lv[level];
    dim3 block(32, 8), grid((dst.w + 127) / 128, (dst.h + 15) / 16, b1 ? 2 : 1);
    launch_chain(k_resize, grid, block, 0, s, src, dst, (const uint8_t*)(b.pyr + src.img_off), b.pyr + dst.img_off, b.xofs + dst.xtab_off,
                 b.alpha + dst.xtab_off, b.yofs + dst.ytab_off, b.beta + dst.ytab_off,
                 (const uint8_t*)(b1 ? b1->pyr + src.img_off : nullptr), b1 ? b1->pyr + dst.img_off : nullptr);
}

// ------------------------------------------------------------------------------------------------ K1' fused pyramid
// ComputePyramid (ORBextractor.cc:1107-1132) in ONE launch. Level l is a resize of level l-1, so a chain of per-level
// launches is serial (7 dependent kernels, ~30 us of launch + drain latency for 1 MB of pixels). Here a CTA takes one
// tile of level 0, and produces the pixels it owns on EVERY level, level after level in shared memory: the resize
// arithmetic is the one of k_resize (bit-exact), pixels in the halo between tiles are computed redundantly.
constexpr int kPyrTW = 64, kPyrTH = 32;

void build_pyr_plan(const OrbGeom& g, const int* xofs, const int* yofs, int base, PyrPlan* plan, std::vector<int>* tab) {
    const int L = g.n_levels;
    PyrPlan& p = *plan;
    p.base = base;  // the level that is read; levels base + 1 .. L - 1 are produced
    p.tw = kPyrTW; p.th = kPyrTH;
    p.ntx = (g.lv[base].w + p.tw - 1) / p.tw;
    p.nty = (g.lv[base].h + p.th - 1) / p.th;
    tab->clear();
    int max_w[kMaxLevels] = {0}, max_h[kMaxLevels] = {0};
    for (int axis = 0; axis < 2; axis++) {
        const int nt = axis == 0 ? p.ntx : p.nty, tile = axis == 0 ? p.tw : p.th;
        auto size_of = [&](int l) { return axis == 0 ? g.lv[l].w : g.lv[l].h; };
        auto ofs = [&](int l, int d) {  // left/top source pixel of level-l pixel d in level l-1 (clamped like k_resize)
            const int v = axis == 0 ? xofs[g.lv[l].xtab_off + d] : yofs[g.lv[l].ytab_off + d];
            return std::min(std::max(v, 0), size_of(l - 1) - 1);
        };
        // f[l][d]: where the chain of left/top sources of level-l pixel d ends in level 0
        std::vector<std::vector<int>> f(L), own(L), n0(L), n1(L);
        for (int l = base; l < L; l++) {
            f[l].resize(size_of(l));
            for (int d = 0; d < size_of(l); d++) f[l][d] = l == base ? d : f[l - 1][ofs(l, d)];
            own[l].assign(nt + 1, size_of(l));
            for (int t = nt - 1; t >= 0; t--) {
                int d = own[l][t + 1];
                while (d > 0 && f[l][d - 1] >= t * tile) d--;
                own[l][t] = d;
            }
            own[l][0] = 0;
            n0[l].assign(nt, 1);
            n1[l].assign(nt, 0);
        }
        for (int t = 0; t < nt; t++) {
            for (int l = L - 1; l >= base; l--) {
                int a = own[l][t], b2 = own[l][t + 1] - 1;  // owned, inclusive (empty if a > b2)
                if (l + 1 < L && n0[l + 1][t] <= n1[l + 1][t]) {
                    const int sa = ofs(l + 1, n0[l + 1][t]), sb = std::min(ofs(l + 1, n1[l + 1][t]) + 1, size_of(l) - 1);
                    if (a > b2) { a = sa; b2 = sb; }
                    else { a = std::min(a, sa); b2 = std::max(b2, sb); }
                }
                n0[l][t] = a; n1[l][t] = b2;
                if (a <= b2) {
                    int& m = axis == 0 ? max_w[l] : max_h[l];
                    m = std::max(m, b2 - a + 1);
                }
            }
        }
        for (int l = base; l < L; l++) {
            (axis == 0 ? p.xoff[l] : p.yoff[l]) = (int)tab->size();
            tab->insert(tab->end(), own[l].begin(), own[l].end());
            tab->insert(tab->end(), n0[l].begin(), n0[l].end());
            tab->insert(tab->end(), n1[l].begin(), n1[l].end());
        }
    }
    int buf = 0, tabs = 0;
    for (int l = base; l < L; l++) {
        buf = std::max(buf, ((max_w[l] + 3 + 3) & ~3) * max_h[l]);  // the base level is stored from a 4-aligned column
        if (l > base) tabs += 2 * (max_w[l] + max_h[l]);
    }
    p.buf_bytes = (buf + 15) & ~15;
    p.tab_smem_ints = tabs;
}

__global__ void __launch_bounds__(256) k_pyramid(OrbGeom g, PyrPlan p, uint8_t* __restrict__ pyr, const int* __restrict__ xofs,
                                                 const short2* __restrict__ alpha, const int* __restrict__ yofs,
                                                 const short2* __restrict__ beta, uint8_t* __restrict__ pyr1) {
    extern __shared__ __align__(16) uint8_t psm[];
    if (blockIdx.z) pyr = pyr1;
    TL_SCOPE(1);
    TR_DECL; TR();
    uint8_t* cur = psm;
    uint8_t* nxt = psm + p.buf_bytes;
    int* stab = reinterpret_cast<int*>(psm + 2 * p.buf_bytes);
    __shared__ int rng[kMaxLevels][8];  // per level: x own0, own1, need0, need1, y own0, own1, need0, need1
    __shared__ int lvi[kMaxLevels][6];  // per level: w, h, pitch, img_off, xtab_off, ytab_off (read once from the parameters)
    const int tid = threadIdx.x, tx = blockIdx.x, ty = blockIdx.y;
    const int L = g.n_levels, B = p.base;
    if (tid < L * 8 && (tid >> 3) >= B) {
        const int l = tid >> 3, k = tid & 7, axis = k >> 2, kk = k & 3;
        const int nt = axis ? p.nty : p.ntx, t = axis ? ty : tx;
        const int* sec = p.tab + (axis ? p.yoff[l] : p.xoff[l]);
        rng[l][k] = kk == 0 ? sec[t] : kk == 1 ? sec[t + 1] : kk == 2 ? sec[nt + 1 + t] : sec[2 * nt + 1 + t];
    } else if (tid >= 128 && tid < 128 + L * 6) {
        const int l = (tid - 128) / 6, k = (tid - 128) - 6 * l;
        const LevelGeom& G = g.lv[l];
        lvi[l][k] = k == 0 ? G.w : k == 1 ? G.h : k == 2 ? G.pitch : k == 3 ? G.img_off : k == 4 ? G.xtab_off : G.ytab_off;
    }
    __syncthreads();
    TR();
    // level 0 (4-byte words of the needed region of the 128-byte pitched level-0 image) and the slices of the resize
    // tables this tile needs on every level are fetched in the same phase: one round of global latency for both
    int pw;  // pitch of `cur`
    int px0 = rng[B][2] & ~3, py0 = rng[B][6];  // origin of `cur` in level coordinates
    {
        const int nh = rng[B][7] - py0 + 1;
        const int nwords = ((rng[B][3] - px0) >> 2) + 1;
        pw = nwords * 4;
        const int pitch0 = lvi[B][2];
        const uint8_t* s0 = pyr + lvi[B][3] + (size_t)py0 * pitch0 + px0;
        const int lanes = tid & 31, rows = tid >> 5;
        for (int yy = rows; yy < nh; yy += 8)
            for (int xx = lanes; xx < nwords; xx += 32)
                reinterpret_cast<uint32_t*>(cur)[yy * nwords + xx] = *reinterpret_cast<const uint32_t*>(s0 + (size_t)yy * pitch0 + 4 * xx);
    }
    // start of level l's staged table slice in stab
    auto slice_off = [&](int l) {
        int o = 0;
        for (int k = B + 1; k < l; k++) o += 2 * (max(rng[k][3] - rng[k][2] + 1, 0) + max(rng[k][7] - rng[k][6] + 1, 0));
        return o;
    };
    {
        // warp w stages level w + 1 (and w + 9 ...): the levels' loads are in flight together
        for (int l = B + 1 + (tid >> 5); l < L; l += 8) {
            const int x0 = rng[l][2], nw = rng[l][3] - x0 + 1, y0 = rng[l][6], nh = rng[l][7] - y0 + 1;
            if (nw <= 0 || nh <= 0) continue;
            int* t = stab + slice_off(l);
            const int xo = lvi[l][4] + x0, yo = lvi[l][5] + y0;
            for (int i = tid & 31; i < nw + nh; i += 32) {
                if (i < nw) {
                    t[i] = xofs[xo + i];
                    t[nw + i] = *reinterpret_cast<const int*>(&alpha[xo + i]);
                } else {
                    const int j = i - nw;
                    t[2 * nw + j] = yofs[yo + j];
                    t[2 * nw + nh + j] = *reinterpret_cast<const int*>(&beta[yo + j]);
                }
            }
        }
    }
    __syncthreads();
    TR();
    int toff = 0;
#pragma unroll 1
    for (int l = B + 1; l < L; l++) {
        const int x0 = rng[l][2], nw = rng[l][3] - x0 + 1, y0 = rng[l][6], nh = rng[l][7] - y0 + 1;
        if (nw <= 0 || nh <= 0) break;  // deeper levels need nothing either
        const int ox0 = rng[l][0], ox1 = rng[l][1], oy0 = rng[l][4], oy1 = rng[l][5];
        const int sw = lvi[l - 1][0], sh = lvi[l - 1][1];
        const int dpitch = lvi[l][2];
        const int* t = stab + toff;
        toff += 2 * (nw + nh);
        uint8_t* dst = pyr + lvi[l][3];
        // thread = (column, row group); four rows per batch so that their shared-memory loads overlap
        const int nwp = (nw + 31) & ~31;
        const int xx = tid % nwp, yg = tid / nwp, ystep = 256 / nwp;
        if (xx < nw && yg < ystep) {
            const int sx = t[xx];
            const int sx1 = min(sx + 1, sw - 1);
            const int aw = t[nw + xx];
            const int a0 = (short)(aw & 0xffff), a1 = aw >> 16;
            const int dx = x0 + xx;
            const bool own_x = dx >= ox0 && dx < ox1;
            for (int yb = yg; yb < nh; yb += 4 * ystep) {
                int v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int yy = min(yb + u * ystep, nh - 1);
                    const int sy = t[2 * nw + yy];
                    const int sy0 = min(max(sy, 0), sh - 1), sy1 = min(max(sy + 1, 0), sh - 1);
                    const int bw = t[2 * nw + nh + yy];
                    const int b0 = (short)(bw & 0xffff), b1 = bw >> 16;
                    const uint8_t* r0p = cur + (sy0 - py0) * pw - px0;
                    const uint8_t* r1p = cur + (sy1 - py0) * pw - px0;
                    const int r0 = (int)r0p[sx] * a0 + (int)r0p[sx1] * a1;
                    const int r1 = (int)r1p[sx] * a0 + (int)r1p[sx1] * a1;
                    v[u] = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int yy = yb + u * ystep;
                    if (yy < nh) {
                        nxt[yy * nw + xx] = (uint8_t)v[u];
                        const int dy = y0 + yy;
                        if (own_x && dy >= oy0 && dy < oy1) dst[(size_t)dy * dpitch + dx] = (uint8_t)v[u];
                    }
                }
            }
        }
        __syncthreads();
        TR();
        { uint8_t* tmp = cur; cur = nxt; nxt = tmp; }
        pw = nw; px0 = x0; py0 = y0;
    }
    TR_PRINT("pyr", blockIdx.x == 3 && blockIdx.y == 3);
}

cudaError_t prepare_pyramid(const PyrPlan& p) {
    const int smem = 2 * p.buf_bytes + 4 * p.tab_smem_ints + 16;
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k_pyramid, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

void launch_pyramid(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1, bool tail) {
    const PyrPlan& p = tail ? b.pyr_tail : b.pyr_plan;
    const int smem = 2 * p.buf_bytes + 4 * p.tab_smem_ints + 16;
    k_pyramid<<<dim3(p.ntx, p.nty, b1 ? 2 : 1), 256, smem, s>>>(g, p, b.pyr, b.xofs, b.alpha, b.yofs, b.beta, b1 ? b1->pyr : nullptr);
}

// ------------------------------------------------------------------------------------------------ K2 FAST per cell
constexpr int kRoiPitch = 96;  // bytes per ROI row in shared memory = TMA box width: multiple of 16 and >= kCellRoiMax + 15,
                               // because the innermost TMA coordinate must be 16-byte aligned (measured: any other x traps
                               // with 'illegal instruction' on sm_100a), so the box starts at iniX & ~15
constexpr int kRoiPitchRaw = kRoiPitch;
constexpr int kBlurBoxW = 96, kBlurBoxH = kBlurTH + 6;   // input box of a blur tile (orb_geom.h): columns x0 - 16 .. x0 + 79 (16-byte aligned start), rows y0 - 3 .. y0 + kBlurTH + 2
constexpr int kRoiTmaBytes = kRoiPitch * kCellRoiMax;  // one 96 x 66 box per cell

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives, sm_90+/sm_100a PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
constexpr int kScDim = 62;     // valid area (<= 60) + 1 px zero border each side
constexpr int kScPitch = 64;

__device__ __forceinline__ bool run9(uint32_t m) {
    m |= m << 16;
    m &= m >> 1;
    m &= m >> 2;
    m &= m >> 4;
    m &= m >> 1;
    return (m & 0xffffu) != 0;
}

// FAST-9/16 response (max over the 16 nine-arcs of the arc minimum of |ring - centre|, minus 1) if the pixel is a
// corner at threshold `th`, else 0.  == cv::FAST's cornerScore for every pixel cv::FAST(th) reports.
// Any nine-arc of the 16-ring covers at least two of the four compass pixels (0, 4, 8, 12), so most pixels are rejected
// after four loads and compares.
__device__ __forceinline__ int fast_score_dev(const uint8_t* c, int th) {
    constexpr int P = kRoiPitch;
    const int v = c[0];
    int r[16];
    // (no compass-point early exit here: the callers' pre-tests have removed the bulk, half of what is left are corners, and
    // the nine-arc test below implies it)
    r[0] = c[3 * P]; r[4] = c[3]; r[8] = c[-3 * P]; r[12] = c[-3];
    r[1] = c[3 * P + 1];   r[2] = c[2 * P + 2];   r[3] = c[P + 3];
    r[5] = c[-P + 3];      r[6] = c[-2 * P + 2];  r[7] = c[-3 * P + 1];
    r[9] = c[-3 * P - 1];  r[10] = c[-2 * P - 2]; r[11] = c[-P - 3];
    r[13] = c[P - 3];      r[14] = c[2 * P - 2];  r[15] = c[3 * P - 1];
    const int hi = v + th, lo = v - th;
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mb |= (uint32_t)(r[k] > hi) << k;
        md |= (uint32_t)(r[k] < lo) << k;
    }
    const bool bright = run9(mb);
    if (!bright && !run9(md)) return 0;
    // a bright and a dark nine-arc cannot coexist (9 + 9 > 16), so only one sign can produce a positive arc minimum
    int e[16];
#pragma unroll
    for (int k = 0; k < 16; k++) e[k] = bright ? r[k] - v : v - r[k];
    int m2[16], m4[16], m8[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m2[k] = min(e[k], e[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) m8[k] = min(m4[k], m4[(k + 4) & 15]);
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) best = max(best, min(m8[k], e[(k + 8) & 15]));
    return best - 1;
}

// exclusive scan of one int per thread over the block (thread order); *total gets the block sum
__device__ __forceinline__ int block_scan_values(int v, int* warp_tmp, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? warp_tmp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        warp_tmp[lane] = w;
    }
    __syncthreads();
    const int base = wid ? warp_tmp[wid - 1] : 0;
    *total = warp_tmp[nw - 1];
    __syncthreads();
    return base + inc - v;
}

// One CTA per grid cell of ComputeKeyPointsOctTree (ORBextractor.cc:789-832), all levels in one launch.
// The cell's ROI (wCell+6 x hCell+6) is staged in shared memory; FAST ignores a 3 px rim, so the valid areas of
// neighbouring cells tile the level disjointly and NMS sees zeros outside its own cell, exactly like cv::FAST on the
// ROI. The iniTh -> minTh fallback is decided per cell on the post-NMS count.
__global__ void __launch_bounds__(256, 6) k_fast_cells(OrbGeom g, const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                    const uint8_t* __restrict__ pyr, int* __restrict__ cell_count,
                                                    uint32_t* __restrict__ cand_xy, uint32_t* __restrict__ cand_ro,
                                                    int* __restrict__ level_cand, int* __restrict__ status, int cell_begin,
                                                    const CUtensorMap* __restrict__ tmap_dev, FastImage2 im1,
                                                    const __grid_constant__ CUtensorMap tmap1) {
    __shared__ __align__(128) uint8_t roi_raw[kCellRoiMax * kRoiPitchRaw + 16];
    __shared__ __align__(8) uint64_t tma_bar;
    const uint8_t* roi;
    TR_DECL; TR();
    __shared__ __align__(16) uint8_t sc[kScDim * kScPitch];
    __shared__ uint32_t row_ini[128];  // two ballot words per valid row (<= 60 rows)
    __shared__ uint16_t surv[60 * 60];  // pixels that pass the compass pre-test (y << 6 | x)
    __shared__ int n_surv;
    __shared__ int row_off[66];
    const int tid = threadIdx.x;
    const int cell = blockIdx.x + cell_begin;
    int l = 0;
    while (l + 1 < g.n_levels && cell >= g.lv[l + 1].cell_base) l++;
    TL_SCOPE(16 + l);
    pdl_release();  // the quadtree launch behind this one may be scheduled; it waits for this grid in pdl_wait()
    const CUtensorMap* tm = &tmap;
    if (blockIdx.z) {  // second image of a stereo pair
        pyr = im1.pyr; cell_count = im1.cell_count; cand_xy = im1.cand_xy; cand_ro = im1.cand_ro; level_cand = im1.level_cand;
        status = im1.status;
        tmap_dev = im1.tmap_dev;
        tm = &tmap1;
    }
    const LevelGeom L = g.lv[l];
    const int c = cell - L.cell_base;
    TR();
    const int ci = c / L.n_cols, cj = c - ci * L.n_cols;
    const int iniX = kBorder + cj * L.w_cell, iniY = kBorder + ci * L.h_cell;
    const int maxX = min(iniX + L.w_cell + 6, L.max_bx), maxY = min(iniY + L.h_cell + 6, L.max_by);
    const int rw = maxX - iniX, rh = maxY - iniY;
    const int vw = rw - 6, vh = rh - 6;
    if (vw <= 0 || vh <= 0) {  // covers the reference's skip rules (:795,:803) and ROIs cv::FAST cannot process
        if (tid == 0) cell_count[cell] = 0;
        return;
    }
    if (use_tma) {
        // stage the ROI with one TMA box load (96 x 66 bytes at (iniX & ~15, iniY) of this level's tensor map; out-of-image
        // bytes are zero filled and never read): a single thread issues it, everybody waits on the mbarrier
        if (tid == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&tma_bar, kRoiTmaBytes);
            tma_load_2d(roi_raw, tmap_dev ? tmap_dev + l : tm, iniX & ~15, iniY, &tma_bar);
        }
        roi = roi_raw + (iniX & 15);
    } else {
        // fallback: 4-byte words from the 4-aligned column at or left of iniX (the pitch is a multiple of 128)
        const int ax = iniX & ~3, shift = iniX - ax;           // shift in 0..3; roi row holds [ax, ax + 4 * nw)
        const int nw = (rw + shift + 3) >> 2;                  // <= 18 words (rw <= 66), within the 96-byte row
        const uint8_t* srow = pyr + L.img_off + (size_t)iniY * L.pitch + ax;
        const int xw = tid & 31;
        if (xw < nw)
            for (int y = tid >> 5; y < rh; y += 8)
                reinterpret_cast<uint32_t*>(roi_raw + y * kRoiPitchRaw)[xw] = *reinterpret_cast<const uint32_t*>(srow + (size_t)y * L.pitch + 4 * xw);
        roi = roi_raw + shift;
    }
    if (use_tma) mbar_wait(&tma_bar, 0);
    TR();
    // warp w owns rows w, w + 8, ...; lane = column (two halves: x = lane, lane + 32) -> no integer divisions, and the
    // raster order needed for the output falls out of ballots (x order) and a scan over rows (y order).
    // The FAST response is expensive (16 ring loads, a min/max network) but most pixels fail the four-compass-point
    // pre-test, and a warp pays for the full evaluation as soon as one lane needs it. So: pass 1 runs the pre-test
    // on every pixel and compacts the survivors into a shared list, pass 2 evaluates the list densely.
    // cv::FAST(iniTh) first; only a cell without any corner is redone at minTh (:809-815) - a corner at iniTh
    // beats every neighbour that is not a corner at iniTh, so the iniTh score map alone decides the iniTh result.
    const int lane = tid & 31, warp = tid >> 5;
    int th = g.ini_th;
    const uint32_t* rowm = row_ini;
    for (int attempt = 0; attempt < 2; attempt++) {
        for (int i = tid; i < kScDim * kScPitch / 4; i += 256) reinterpret_cast<uint32_t*>(sc)[i] = 0;
        if (tid == 0) n_surv = 0;
        if (tid < 128) row_ini[tid] = 0;
        __syncthreads();
        for (int y = warp; y < vh; y += 8) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int x = lane + 32 * half;
                bool pass = false;
                if (x < vw) {
                    const uint8_t* c = roi + (y + 3) * kRoiPitch + (x + 3);
                    const int v = c[0], r0 = c[3 * kRoiPitch], r4 = c[3], r8 = c[-3 * kRoiPitch], r12 = c[-3];
                    const int hi = v + th, lo = v - th;
                    pass = (r0 > hi) + (r4 > hi) + (r8 > hi) + (r12 > hi) >= 2 || (r0 < lo) + (r4 < lo) + (r8 < lo) + (r12 < lo) >= 2;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&n_surv, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (pass) surv[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(y << 6 | x);
                }
            }
        }
        __syncthreads();
        TR();
        const int ns = n_surv;
        for (int i = tid; i < ns; i += 256) {
            const int y = surv[i] >> 6, x = surv[i] & 63;
            sc[(y + 1) * kScPitch + (x + 1)] = (uint8_t)fast_score_dev(roi + (y + 3) * kRoiPitch + (x + 3), th);
        }
        __syncthreads();
        TR();
        // strict 8-neighbour maxima of the score tile (non-corners are 0). Only pixels of the survivor list can have a
        // non-zero score, so only those are tested; their bits go into two mask words per row.
        bool any_local = false;
        for (int i = tid; i < ns; i += 256) {
            const int y = surv[i] >> 6, x = surv[i] & 63;
            const uint8_t* q = sc + (y + 1) * kScPitch + (x + 1);
            const int s = q[0];
            if (s != 0 && s > q[-kScPitch - 1] && s > q[-kScPitch] && s > q[-kScPitch + 1] && s > q[-1] && s > q[1] &&
                s > q[kScPitch - 1] && s > q[kScPitch] && s > q[kScPitch + 1]) {
                atomicOr(&row_ini[y * 2 + (x >> 5)], 1u << (x & 31));
                any_local = true;
            }
        }
        const int any = __syncthreads_or(any_local);
        TR();
        if (any || g.min_th >= g.ini_th) break;
        th = g.min_th;
    }
    if (warp == 0) {  // exclusive scan of the per-row counts (vh <= 60 rows: two per lane)
        const int y0 = 2 * lane, y1 = 2 * lane + 1;
        const int c0 = y0 < vh ? __popc(rowm[y0 * 2]) + __popc(rowm[y0 * 2 + 1]) : 0;
        const int c1 = y1 < vh ? __popc(rowm[y1 * 2]) + __popc(rowm[y1 * 2 + 1]) : 0;
        int inc = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (y0 < vh) row_off[y0] = inc - c0 - c1;
        if (y1 < vh) row_off[y1] = inc - c1;
        if (lane == 31) {
            row_off[64] = inc;
            // reserve this cell's stretch of the level's candidate list (a strict 8-neighbour maximum has no
            // 8-neighbour that is one too, so a cell holds at most `slot` corners and the list cannot overflow)
            row_off[65] = inc > 0 && inc <= L.slot ? atomicAdd(&level_cand[l], inc) : 0;
        }
    }
    __syncthreads();
    const int total = row_off[64];
    if (total > L.slot) {  // impossible; never truncate silently
        if (tid == 0) { atomicExch(status, 101); cell_count[cell] = 0; }
        return;
    }
    const int base = L.cand_base + row_off[65];
    const uint32_t order0 = 0xffffffu - (uint32_t)(c * L.slot);
    for (int y = warp; y < vh; y += 8) {
        const uint32_t m0 = rowm[y * 2], m1 = rowm[y * 2 + 1];
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t m = half ? m1 : m0;
            if (m >> lane & 1u) {
                const int x = lane + 32 * half;
                const int pos = row_off[y] + (half ? __popc(m0) : 0) + __popc(m & lt);
                // coordinates relative to (minBorderX, minBorderY): FAST's ROI coordinate + j*wCell (:822-823)
                cand_xy[base + pos] = (uint32_t)(x + 3 + cj * L.w_cell) | (uint32_t)(y + 3 + ci * L.h_cell) << 16;
                cand_ro[base + pos] = (uint32_t)sc[(y + 1) * kScPitch + (x + 1)] << 24 | (order0 - (uint32_t)pos);
            }
        }
    }
    if (tid == 0) cell_count[cell] = total;
    TR();
    TR_PRINT("fast", cell == 100 || cell == 900 || cell == 1200);
}

// ---- The same per-cell FAST, kFastG cells per CTA (64 threads = 2 warps per cell). One cell per CTA leaves every fixed cost -
// the launch slot, the TMA round trip, eight block barriers, the scan - to be paid for ~1000 pixels (4 per thread): a cell CTA
// lives ~10 us whatever its size, and the 2 462 cell CTAs of a stereo pair were ~28 of the ~46 us of GPU time a frame occupies.
// Four cells share those costs here (16 pixels per thread, a quarter of the CTAs); each cell keeps its own ROI, score tile,
// survivor list, masks and TMA barrier, and the iniTh -> minTh fallback is still decided per cell. Same results bit for bit.
constexpr int kFastG = 4;
constexpr int kFastTpg = 256 / kFastG;  // threads per cell
struct FastCellSmem {
    uint8_t roi_raw[kCellRoiMax * kRoiPitchRaw + 16];
    uint8_t sc[kScDim * kScPitch];
    uint32_t row_ini[128];
    uint16_t surv[60 * 60];
    int row_off[66];
    int n_surv, any, active, pad;
    uint64_t tma_bar;
    uint8_t tail[112];  // keeps every cell's ROI 128-byte aligned (the TMA destination)
};
static_assert(sizeof(FastCellSmem) % 128 == 0, "per-cell shared block must keep the ROI 128-byte aligned");

__global__ void __launch_bounds__(256, 3) k_fast_cells_g(OrbGeom g, const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                      const uint8_t* __restrict__ pyr, int* __restrict__ cell_count,
                                                      uint32_t* __restrict__ cand_xy, uint32_t* __restrict__ cand_ro,
                                                      int* __restrict__ level_cand, int* __restrict__ status, int cell_begin, int cell_end,
                                                      const CUtensorMap* __restrict__ tmap_dev, FastImage2 im1,
                                                      const __grid_constant__ CUtensorMap tmap1) {
    extern __shared__ __align__(128) uint8_t fsm_raw[];
    const int tid = threadIdx.x, grp = tid / kFastTpg, tig = tid - grp * kFastTpg;
    const int lane = tid & 31, wg = tig >> 5;
    constexpr int kWpg = kFastTpg / 32;
    FastCellSmem& S = reinterpret_cast<FastCellSmem*>(fsm_raw)[grp];
    const int cell = cell_begin + blockIdx.x * kFastG + grp;
    const bool in_range = cell < cell_end;
    int l = 0;
    {
        const int cl = in_range ? cell : cell_end - 1;
        while (l + 1 < g.n_levels && cl >= g.lv[l + 1].cell_base) l++;
    }
    TL_SCOPE(16 + l);
    pdl_release();  // the quadtree launch behind this one may be scheduled; it waits for this grid in pdl_wait()
    const CUtensorMap* tm = &tmap;
    if (blockIdx.z) {  // second image of a stereo pair
        pyr = im1.pyr; cell_count = im1.cell_count; cand_xy = im1.cand_xy; cand_ro = im1.cand_ro; level_cand = im1.level_cand;
        status = im1.status;
        tmap_dev = im1.tmap_dev;
        tm = &tmap1;
    }
    const LevelGeom L = g.lv[l];
    const int c = (in_range ? cell : cell_end - 1) - L.cell_base;
    const int ci = c / L.n_cols, cj = c - ci * L.n_cols;
    const int iniX = kBorder + cj * L.w_cell, iniY = kBorder + ci * L.h_cell;
    const int maxX = min(iniX + L.w_cell + 6, L.max_bx), maxY = min(iniY + L.h_cell + 6, L.max_by);
    const int rw = maxX - iniX, rh = maxY - iniY;
    const int vw = rw - 6, vh = rh - 6;
    const bool live = in_range && vw > 0 && vh > 0;  // covers the reference's skip rules (:795,:803) and ROIs cv::FAST cannot process
    if (in_range && !live && tig == 0) cell_count[cell] = 0;
    const uint8_t* roi;
    if (use_tma) {
        if (tig == 0) mbar_init(&S.tma_bar, 1);
        __syncthreads();
        if (tig == 0 && live) {
            mbar_expect_tx(&S.tma_bar, kRoiTmaBytes);
            tma_load_2d(S.roi_raw, tmap_dev ? tmap_dev + l : tm, iniX & ~15, iniY, &S.tma_bar);
        }
        roi = S.roi_raw + (iniX & 15);
    } else {
        const int ax = iniX & ~3, shift = iniX - ax;
        const int nw = (rw + shift + 3) >> 2;
        const uint8_t* srow = pyr + L.img_off + (size_t)iniY * L.pitch + ax;
        const int xw = tig & 31;
        if (live && xw < nw)
            for (int y = tig >> 5; y < rh; y += kWpg)
                reinterpret_cast<uint32_t*>(S.roi_raw + y * kRoiPitchRaw)[xw] = *reinterpret_cast<const uint32_t*>(srow + (size_t)y * L.pitch + 4 * xw);
        roi = S.roi_raw + shift;
    }
    if (use_tma && live) mbar_wait(&S.tma_bar, 0);
    if (tig == 0) S.active = live;
    int th = g.ini_th;
    for (int attempt = 0; attempt < 2; attempt++) {
        const bool run = live && (attempt == 0 || S.active);
        if (run) {
            for (int i = tig; i < kScDim * kScPitch / 4; i += kFastTpg) reinterpret_cast<uint32_t*>(S.sc)[i] = 0;
            if (tig == 0) { S.n_surv = 0; S.any = 0; }
            for (int i = tig; i < 128; i += kFastTpg) S.row_ini[i] = 0;
        }
        __syncthreads();
        if (run) {
            // Pre-test, four pixels per lane with byte-wise SIMD. Any nine-arc of the ring contains pixel 0 or 8 and pixel 4 or
            // 12, and every pixel of the arc differs from the centre by more than th, so
            //     (|r0 - v| > th or |r8 - v| > th) and (|r4 - v| > th or |r12 - v| > th)
            // holds for every corner: a superset filter (it keeps 7.9 % of the pixels of a bench frame; the exact score below
            // rejects the rest). |a - b| is one VABSDIFF4; "byte > th" is the carry into bit 7 of (byte & 0x7f) + (0x7f - th),
            // or bit 7 of the byte itself. The ROI row pitch is a multiple of 4 and the cell's byte offset is the same in every
            // row, so unaligned quads are two aligned words and one funnel shift with a cell-uniform amount.
            const int nq = (vw + 3) >> 2, n_items = vh * nq;
            const uint32_t inv = (65536u + nq - 1) / nq;  // item / nq == item * inv >> 16 for item < 1100, nq <= 16
            const int a3 = (int)(roi - S.roi_raw) + 3;     // byte offset of ROI column 3 (valid x = 0) in a raw row
            const int off0 = a3 & 3, sh0 = 8 * off0;
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(S.roi_raw);
            constexpr int kRowW = kRoiPitch / 4;
            const uint32_t th4 = (uint32_t)(0x7f - th) * 0x01010101u;
            for (int ib = 0; ib < n_items; ib += kFastTpg) {
                const int item = ib + tig;
                uint32_t m = 0;
                int y = 0, q = 0;
                if (item < n_items) {
                    y = (int)(((uint32_t)item * inv) >> 16);
                    q = item - y * nq;
                    const int cw = (a3 >> 2) + q;
                    const uint32_t* r3 = rw + (y + 3) * kRowW + cw;
                    const uint32_t wm = r3[-1], w0 = r3[0], w1 = r3[1], w2 = r3[2];
                    const uint32_t* ru = rw + y * kRowW + cw;        // ring pixel 8: three rows up
                    const uint32_t* rd = rw + (y + 6) * kRowW + cw;  // ring pixel 0: three rows down
                    const uint32_t v = __funnelshift_rc(w0, w1, sh0);
                    const uint32_t p8 = __funnelshift_rc(ru[0], ru[1], sh0), p0 = __funnelshift_rc(rd[0], rd[1], sh0);
                    const uint32_t p12 = __funnelshift_rc(wm, w0, sh0 + 8);  // bytes off0 - 3 .. off0 of (w0 : wm)
                    const uint32_t p4 = off0 <= 1 ? __funnelshift_rc(w0, w1, sh0 + 24) : __funnelshift_rc(w1, w2, sh0 - 8);
                    const uint32_t d0 = __vabsdiffu4(p0, v), d8 = __vabsdiffu4(p8, v), d4 = __vabsdiffu4(p4, v), d12 = __vabsdiffu4(p12, v);
                    const uint32_t g08 = ((d0 & 0x7f7f7f7fu) + th4) | ((d8 & 0x7f7f7f7fu) + th4) | d0 | d8;
                    const uint32_t g4c = ((d4 & 0x7f7f7f7fu) + th4) | ((d12 & 0x7f7f7f7fu) + th4) | d4 | d12;
                    m = g08 & g4c & 0x80808080u;
                    const int left = vw - 4 * q;  // valid pixels of this quad
                    if (left < 4) m &= (1u << (8 * left)) - 1u;
                }
                if (m) {  // the survivor list is a set (scores land by position, NMS sets mask bits): any order will do
                    const int base = atomicAdd(&S.n_surv, __popc(m));
                    const uint32_t code = (uint32_t)(y << 6 | 4 * q);
#pragma unroll
                    for (int k = 0; k < 4; k++)  // predicated stores instead of a divergent bit loop
                        if (m & (0x80u << (8 * k))) S.surv[base + __popc(m & ((0x80u << (8 * k)) - 1u))] = (uint16_t)(code + k);
                }
            }
        }
        __syncthreads();
        if (run) {
            const int ns = S.n_surv;
            for (int i = tig; i < ns; i += kFastTpg) {
                const int y = S.surv[i] >> 6, x = S.surv[i] & 63;
                S.sc[(y + 1) * kScPitch + (x + 1)] = (uint8_t)fast_score_dev(roi + (y + 3) * kRoiPitch + (x + 3), th);
            }
        }
        __syncthreads();
        if (run) {
            const int ns = S.n_surv;
            bool any_local = false;
            for (int i = tig; i < ns; i += kFastTpg) {
                const int y = S.surv[i] >> 6, x = S.surv[i] & 63;
                const uint8_t* q = S.sc + (y + 1) * kScPitch + (x + 1);
                const int sv = q[0];
                if (sv != 0 && sv > q[-kScPitch - 1] && sv > q[-kScPitch] && sv > q[-kScPitch + 1] && sv > q[-1] && sv > q[1] &&
                    sv > q[kScPitch - 1] && sv > q[kScPitch] && sv > q[kScPitch + 1]) {
                    atomicOr(&S.row_ini[y * 2 + (x >> 5)], 1u << (x & 31));
                    any_local = true;
                }
            }
            if (any_local) S.any = 1;
        }
        __syncthreads();
        // a cell without any corner is redone at minTh (:809-815); the other cells of the CTA sit the second round out
        const bool retry = run && !S.any && g.min_th < g.ini_th && attempt == 0;
        const int again = __syncthreads_or(retry);
        if (tig == 0) S.active = retry;
        if (!again) break;
        th = g.min_th;
        __syncthreads();
    }
    if (live && wg == 0) {  // exclusive scan of the per-row counts (vh <= 60 rows: two per lane)
        const uint32_t* rowm = S.row_ini;
        const int y0 = 2 * lane, y1 = 2 * lane + 1;
        const int c0 = y0 < vh ? __popc(rowm[y0 * 2]) + __popc(rowm[y0 * 2 + 1]) : 0;
        const int c1 = y1 < vh ? __popc(rowm[y1 * 2]) + __popc(rowm[y1 * 2 + 1]) : 0;
        int inc = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (y0 < vh) S.row_off[y0] = inc - c0 - c1;
        if (y1 < vh) S.row_off[y1] = inc - c1;
        if (lane == 31) {
            S.row_off[64] = inc;
            S.row_off[65] = inc > 0 && inc <= L.slot ? atomicAdd(&level_cand[l], inc) : 0;
        }
    }
    __syncthreads();
    if (!live) return;
    const int total = S.row_off[64];
    if (total > L.slot) {  // impossible; never truncate silently
        if (tig == 0) { atomicExch(status, 101); cell_count[cell] = 0; }
        return;
    }
    const int base = L.cand_base + S.row_off[65];
    const uint32_t order0 = 0xffffffu - (uint32_t)(c * L.slot);
    if (tig < vh) {  // a cell keeps a handful of corners: one thread per row walks that row's mask bits (raster order inside the row)
        const int y = tig;
        int pos = S.row_off[y];
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t m = S.row_ini[y * 2 + half];
            while (m) {
                const int x = __ffs(m) - 1 + 32 * half;
                m &= m - 1;
                // coordinates relative to (minBorderX, minBorderY): FAST's ROI coordinate + j*wCell (:822-823)
                cand_xy[base + pos] = (uint32_t)(x + 3 + cj * L.w_cell) | (uint32_t)(y + 3 + ci * L.h_cell) << 16;
                cand_ro[base + pos] = (uint32_t)S.sc[(y + 1) * kScPitch + (x + 1)] << 24 | (order0 - (uint32_t)pos);
                pos++;
            }
        }
    }
    if (tig == 0) cell_count[cell] = total;
}

// cuTensorMapEncodeTiled is a driver-API symbol; it is resolved at run time through the runtime so that the library
// does not link libcuda (and still loads on a machine without a driver, where every compute call fails loudly).
bool encode_tma_maps(const OrbGeom& g, uint8_t* pyr, TmaMaps* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    for (int l = 0; l < g.n_levels; l++) {
        const LevelGeom& L = g.lv[l];
        const cuuint64_t dims[2] = {(cuuint64_t)L.w, (cuuint64_t)L.h};
        const cuuint64_t strides[1] = {(cuuint64_t)L.pitch};
        const cuuint32_t box[2] = {(cuuint32_t)kRoiPitch, (cuuint32_t)kCellRoiMax};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = ((EncodeFn)fn)(&out->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, pyr + L.img_off, dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
        const cuuint32_t bbox[2] = {(cuuint32_t)kBlurBoxW, (cuuint32_t)kBlurBoxH};
        const CUresult rb = ((EncodeFn)fn)(&out->mb[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, pyr + L.img_off, dims, strides, bbox, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rb != CUDA_SUCCESS) return false;
    }
    return true;
}

// level < 0: every level (one launch each); else only that level (lets a level start as soon as it is resized)
static FastImage2 fast_image2(const OrbBuffers* b1, bool dev_maps) {
    FastImage2 im = {};
    if (b1) {
        im.pyr = b1->pyr; im.cell_count = b1->cell_count; im.cand_xy = b1->cand_xy; im.cand_ro = b1->cand_ro;
        im.level_cand = b1->level_cand; im.status = b1->status;
        im.tmap_dev = dev_maps ? b1->tma_dev : nullptr;
    }
    return im;
}

static bool fast_grouped() {  // CORB_FAST_GROUP=0: one cell per CTA (the A/B switch; same results)
    static const bool on = [] {
        const char* e = getenv("CORB_FAST_GROUP");
        if (e && atoi(e) == 0) return false;
        return cudaFuncSetAttribute(k_fast_cells_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFastG * sizeof(FastCellSmem))) == cudaSuccess;
    }();
    return on;
}

void launch_fast_cells(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s, const OrbBuffers* b1) {
    // one launch per level: the level's tensor map travels as a __grid_constant__ parameter (the form TMA expects)
    const int l0 = level < 0 ? 0 : level, l1 = level < 0 ? g.n_levels : level + 1;
    for (int l = l0; l < l1; l++) {
        const int n = g.lv[l].n_cols * g.lv[l].n_rows;
        if (fast_grouped())
            k_fast_cells_g<<<dim3((n + kFastG - 1) / kFastG, 1, b1 ? 2 : 1), 256, kFastG * sizeof(FastCellSmem), s>>>(
                g, b.tma_maps->m[l], b.use_tma, b.pyr, b.cell_count, b.cand_xy, b.cand_ro, b.level_cand, b.status, g.lv[l].cell_base,
                g.lv[l].cell_base + n, nullptr, fast_image2(b1, false), (b1 ? b1 : &b)->tma_maps->m[l]);
        else
            k_fast_cells<<<dim3(n, 1, b1 ? 2 : 1), 256, 0, s>>>(g, b.tma_maps->m[l], b.use_tma, b.pyr, b.cell_count, b.cand_xy, b.cand_ro,
                                                                b.level_cand, b.status, g.lv[l].cell_base, nullptr, fast_image2(b1, false),
                                                                (b1 ? b1 : &b)->tma_maps->m[l]);
    }
}

// every cell of every level in one launch; the per-level tensor maps are read from their device copy
void launch_fast_all(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1) {
    k_fast_cells<<<dim3(g.n_cells, 1, b1 ? 2 : 1), 256, 0, s>>>(g, b.tma_maps->m[0], b.use_tma, b.pyr, b.cell_count, b.cand_xy, b.cand_ro,
                                                                b.level_cand, b.status, 0, b.tma_dev, fast_image2(b1, true),
                                                                b.tma_maps->m[0]);
}

// ------------------------------------------------------------------------------------------------ K5 Gaussian 7x7
// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on u8 (ORBextractor.cc:1086): fixed-point taps
// {18,34,48,56,48,34,18}/256 on both axes, exact accumulation, one rounding (s + 2^15) >> 16.

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

constexpr int kBlurThreads = 128;  // 64 x 32 outputs per CTA: 16 per thread, so the per-thread prologue is paid once per 16 pixels
__global__ void __launch_bounds__(kBlurThreads) k_blur(OrbGeom g, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur,
                                              const uint8_t* __restrict__ pyr1, uint8_t* __restrict__ blur1,
                                              const CUtensorMap* __restrict__ maps, const CUtensorMap* __restrict__ maps1) {
    if (blockIdx.z) { pyr = pyr1; blur = blur1; maps = maps1; }
    // input box of the tile: [kBlurBoxH][kBlurBoxW], pixel (x0 - 16 + xx, y0 - 3 + yy); one TMA box load (maps != NULL) or
    // 32-bit loads. What lies outside the image arrives as zeros and is rewritten with the REFLECT_101 pixels below.
    __shared__ __align__(128) uint8_t in[kBlurBoxH][kBlurBoxW];
    __shared__ __align__(16) uint16_t hb[kBlurTH + 6][kBlurTW];
    __shared__ __align__(8) uint64_t bar;
    TL_SCOPE(48);
    const int tid = threadIdx.x;
    int l = 0;
    while (l + 1 < g.n_levels && (int)blockIdx.x >= g.lv[l + 1].blur_tile_base) l++;
    const LevelGeom L = g.lv[l];
    const int t = blockIdx.x - L.blur_tile_base;
    const int ty = t / L.blur_tiles_x, tx = t - ty * L.blur_tiles_x;
    const int x0 = tx * kBlurTW, y0 = ty * kBlurTH;
    if (maps) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, kBlurBoxW * kBlurBoxH);
            tma_load_2d(&in[0][0], maps + l, x0 - 16, y0 - 3, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        const uint8_t* src = pyr + L.img_off;
        for (int i = tid; i < kBlurBoxH * (kBlurBoxW / 4); i += kBlurThreads) {
            const int yy = i / (kBlurBoxW / 4), xw = i - yy * (kBlurBoxW / 4);
            const int gy = y0 - 3 + yy, gx = x0 - 16 + 4 * xw;
            uint32_t v = 0;
            if (gy >= 0 && gy < L.h && gx >= 0 && gx < L.pitch) v = *reinterpret_cast<const uint32_t*>(src + (size_t)gy * L.pitch + gx);
            reinterpret_cast<uint32_t*>(&in[yy][0])[xw] = v;
        }
        __syncthreads();
    }
    // tiles on the image border: the frame of 3 pixels outside the image is the reflection of pixels inside this box
    const bool edge = x0 < 3 || y0 < 3 || x0 + kBlurTW + 3 > L.w || y0 + kBlurTH + 3 > L.h;
    if (edge) {
        // only the frame itself is visited: the (at most) three columns left and right of the image and the three rows above and
        // below it, as far as they fall into this box (corners are written by both strips with the same value)
        constexpr int kColItems = 6 * kBlurBoxH, kRowItems = 6 * (kBlurTW + 6);
        for (int i = tid; i < kColItems + kRowItems; i += kBlurThreads) {
            int gx, gy;
            if (i < kColItems) {
                const int c = i / kBlurBoxH;
                gy = y0 - 3 + (i - c * kBlurBoxH);
                gx = c < 3 ? c - 3 : L.w + c - 3;
            } else {
                const int k = i - kColItems, r = k / (kBlurTW + 6);
                gx = x0 - 3 + (k - r * (kBlurTW + 6));
                gy = r < 3 ? r - 3 : L.h + r - 3;
            }
            const int yy = gy - (y0 - 3), xx = gx - (x0 - 16);
            if (yy >= 0 && yy < kBlurBoxH && xx >= 13 && xx < 13 + kBlurTW + 6 && gy >= -3 && gy <= L.h + 2 && gx >= -3 && gx <= L.w + 2 &&
                (gy < 0 || gy >= L.h || gx < 0 || gx >= L.w)) {
                const int sy = reflect101(gy, L.h) - (y0 - 3), sx = reflect101(gx, L.w) - (x0 - 16);
                in[yy][xx] = in[sy][sx];  // sources are inside the image, destinations outside: no overlap
            }
        }
        __syncthreads();
    }
    // horizontal pass, 4 outputs per item from 3 words: output pixel k takes the bytes p[k + 1 .. k + 7] of the 12 loaded ones as two
    // byte windows (funnel shifts) and two dot products with the taps packed as bytes (DP4A): 18 34 48 56 | 48 34 18 0
    constexpr uint32_t kTapA = 18u | 34u << 8 | 48u << 16 | 56u << 24, kTapB = 48u | 34u << 8 | 18u << 16;
    for (int i = tid; i < (kBlurTH + 6) * (kBlurTW / 4); i += kBlurThreads) {
        const int yy = i / (kBlurTW / 4), xq = i - yy * (kBlurTW / 4);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&in[yy][12 + 4 * xq]);  // pixels x0 - 4 + 4 xq .. + 11
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        const uint32_t v0 = __dp4a(__funnelshift_r(w1, w2, 8), kTapB, __dp4a(__funnelshift_r(w0, w1, 8), kTapA, 0u));
        const uint32_t v1 = __dp4a(__funnelshift_r(w1, w2, 16), kTapB, __dp4a(__funnelshift_r(w0, w1, 16), kTapA, 0u));
        const uint32_t v2 = __dp4a(__funnelshift_r(w1, w2, 24), kTapB, __dp4a(__funnelshift_r(w0, w1, 24), kTapA, 0u));
        const uint32_t v3 = __dp4a(w2, kTapB, __dp4a(w1, kTapA, 0u));
        *reinterpret_cast<uint2*>(&hb[yy][4 * xq]) = make_uint2(v0 | v1 << 16, v2 | v3 << 16);
    }
    __syncthreads();
    // vertical pass: a thread owns 2 columns x 8 rows; the 14 rows it needs slide through registers (one 32-bit load per row)
    const int lx = (tid & 31) * 2, ly = (tid >> 5) * 8;
    const int gx = x0 + lx;
    if (gx < L.w) {
        uint32_t lo[14], hi[14];
#pragma unroll
        for (int r = 0; r < 14; r++) {
            const uint32_t h2 = *reinterpret_cast<const uint32_t*>(&hb[ly + r][lx]);
            lo[r] = h2 & 0xffffu;
            hi[r] = h2 >> 16;
        }
        uint8_t* out = blur + L.img_off + (size_t)(y0 + ly) * L.pitch + gx;
        const int rows = min(8, L.h - (y0 + ly));
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (r < rows) {
                const uint32_t a = 18u * (lo[r] + lo[r + 6]) + 34u * (lo[r + 1] + lo[r + 5]) + 48u * (lo[r + 2] + lo[r + 4]) + 56u * lo[r + 3];
                const uint32_t c = 18u * (hi[r] + hi[r + 6]) + 34u * (hi[r + 1] + hi[r + 5]) + 48u * (hi[r + 2] + hi[r + 4]) + 56u * hi[r + 3];
                // the pitch is a multiple of 128 and gx is even: a 16-bit store; pixel L.w (odd widths) falls into the row padding
                *reinterpret_cast<uint16_t*>(out) = (uint16_t)(((a + 32768u) >> 16) | (((c + 32768u) >> 16) << 8));
            }
            out += L.pitch;
        }
    }
}

void launch_blur(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1) {
    const CUtensorMap* m0 = b.use_tma && b.tma_dev ? b.tma_dev + kMaxLevels : nullptr;
    const CUtensorMap* m1 = b1 && b1->use_tma && b1->tma_dev ? b1->tma_dev + kMaxLevels : nullptr;
    if (b1 && !m1) m0 = nullptr;  // both images take the same path
    k_blur<<<dim3(g.blur_tiles, 1, b1 ? 2 : 1), kBlurThreads, 0, s>>>(g, b.pyr, b.blur, b1 ? b1->pyr : nullptr, b1 ? b1->blur : nullptr, m0, b1 ? m1 : nullptr);
}

// ------------------------------------------------------------------------------------------------ K3 quadtree
// DistributeOctTree (ORBextractor.cc:539-763) as a level-synchronous array algorithm, one CTA per pyramid level.
//
// The reference keeps a std::list of nodes; every split pushes the non-empty children to the front (n1..n4) and
// erases the parent. Storing nodes *in list order* makes the list implicit: after a step in which the processed
// parents create T children in creation order 0..T-1, child j sits at position T-1-j and every unprocessed node
// keeps its relative order behind them. A step therefore is: count keys per (node, quadrant) with shared-memory
// atomics, scan, scatter. The near-quota phase (:660-733) sorts the freshly created expandable nodes by
// (count, address) and splits from the back until the list reaches N; nodes created in one step have addresses
// in creation order (canonical tie-break), i.e. reverse list position, so its processing order is
// (count descending, position ascending) and the early stop is the first prefix whose size reaches N.
#ifndef CORB_OCT_THREADS
#define CORB_OCT_THREADS 256  // measured on B200, latency of a frame alone: 128 -> 79.0, 256 -> 70.3, 512 -> 72.1, 1024 -> 74.9 us
#endif
constexpr int kOctThreads = CORB_OCT_THREADS;

struct OctLayout {
    int cntA, cntB, bndA, bndB, cc, procpos, scan, crank, krank, fresh, candf, warp_tmp, sh, scan64, warp_tmp64, keys_xy, keys_node, keys_ro, total;
    int fast;  // closed-form path tabulated (n_ini small enough), with the arrays below
    int f_hist, f_code, f_owner, f_flag, f_stat, f_lut;
};
// Closed-form path (see k_octtree): quadtree depths 0..kOctDh are tabulated from one histogram of the keys' depth-kOctDh
// cells; the cell index of depth d is root * 4^d + sum q_i * 4^(d-i).
constexpr int kOctDh = 5;
constexpr int kOctFastMaxIni = 8;
constexpr int kOctSmemLimit = 200 * 1024;
// f_flag doubles as the bins of the counting sort of the near-quota rounds: (warps + 1) x 256 ints
__host__ __device__ inline int kOctFlagInts(int TC) { return TC + 4 > (CORB_OCT_THREADS / 32 + 1) * 256 ? TC + 4 : (CORB_OCT_THREADS / 32 + 1) * 256; }
__host__ __device__ inline int oct_depth_off(int n_ini, int d) { return n_ini * (((1 << (2 * d)) - 1) / 3); }
__host__ __device__ inline OctLayout oct_layout(int NC, int key_cap, int n_lut, int n_ini) {
    OctLayout o;
    int p = 0;
    // every array starts on a 16-byte boundary and is padded to a multiple of 16 bytes (vectorised scans read whole int4)
#define OCT_ALLOC(field, bytes) do { o.field = p; p += ((bytes) + 15) & ~15; } while (0)
    OCT_ALLOC(cntA, 4 * NC);
    OCT_ALLOC(cntB, 4 * NC);
    OCT_ALLOC(bndA, 8 * NC);
    OCT_ALLOC(bndB, 8 * NC);
    OCT_ALLOC(cc, 16 * NC);
    OCT_ALLOC(procpos, 4 * NC);
    OCT_ALLOC(scan, 4 * NC);
    OCT_ALLOC(crank, 4 * NC);
    OCT_ALLOC(krank, 4 * NC);
    OCT_ALLOC(fresh, 2 * ((NC + 3) & ~3));  // fresh A and B
    OCT_ALLOC(candf, NC);
    OCT_ALLOC(warp_tmp, 4 * 36);
    OCT_ALLOC(sh, 4 * 8);
    OCT_ALLOC(scan64, 8 * NC);
    OCT_ALLOC(warp_tmp64, 8 * 36);
    OCT_ALLOC(keys_xy, 4 * key_cap);
    OCT_ALLOC(keys_node, 2 * key_cap);
    OCT_ALLOC(keys_ro, 4 * key_cap);
    o.f_hist = o.f_code = o.f_owner = o.f_flag = o.f_stat = o.f_lut = 0;
    const int HC = n_ini << (2 * kOctDh), TC = oct_depth_off(n_ini, kOctDh + 1);
    {   // the closed-form tables are optional: keep them only while the whole layout stays inside kOctSmemLimit
        const int extra = 4 * (HC + 4) + 4 * kOctFlagInts(TC) + 64 + 2 * (n_lut + 2) + 4 * 16;
        o.fast = n_ini <= kOctFastMaxIni && p + extra <= kOctSmemLimit;
    }
    if (o.fast) {
        OCT_ALLOC(f_hist, 4 * (HC + 4));
        OCT_ALLOC(f_flag, 4 * kOctFlagInts(TC));
        OCT_ALLOC(f_stat, 4 * 16);
        OCT_ALLOC(f_lut, 2 * (n_lut + 2));
        o.f_code = o.keys_node;  // a key's depth-Dh cell is replaced in place by the list position of its node
        o.f_owner = o.f_flag;    // the cell -> node map is built after the last use of the flag / bin array (2 HC <= its size)
    }
#undef OCT_ALLOC
    o.total = p;
    return o;
}

// entries of a level's path tables: xs over the window width (even-padded), ys over the window height
__host__ __device__ inline int oct_lut_entries(const LevelGeom& L) {
    return ((L.max_bx - kBorder + 1) & ~1) + ((L.max_by - kBorder + 1) & ~1);
}

// Host: xs[x] = root << 2Dh | x-bits of the path on the even bit positions, ys[y] = y-bits on the odd positions, so the
// depth-Dh cell of a key is xs[x] | ys[y]. Same float and integer arithmetic as DistributeOctTree / DivideNode
// (ORBextractor.cc:481-489, 543-559): root = (int)(x / hX), halves = ceil(extent / 2).
void build_oct_lut(const LevelGeom& L, uint16_t* out) {
    const int W = L.max_bx - kBorder, H = L.max_by - kBorder, Wp = (W + 1) & ~1;
    for (int x = 0; x < W; x++) {
        const int ni = (int)((float)x / L.h_x);
        int x0 = (int)(L.h_x * (float)ni), x1 = (int)(L.h_x * (float)(ni + 1));
        int code = 0;
        for (int d = 0; d < kOctDh; d++) {
            const int xm = x0 + ((x1 - x0 + 1) >> 1);
            const int bit = x < xm ? 0 : 1;
            if (bit) x0 = xm; else x1 = xm;
            code = code * 4 + bit;
        }
        out[x] = (uint16_t)(ni << (2 * kOctDh) | code);
    }
    for (int x = W; x < Wp; x++) out[x] = 0;
    for (int y = 0; y < H; y++) {
        int y0 = 0, y1 = H, code = 0;
        for (int d = 0; d < kOctDh; d++) {
            const int ym = y0 + ((y1 - y0 + 1) >> 1);
            const int bit = y < ym ? 0 : 1;
            if (bit) y0 = ym; else y1 = ym;
            code = code * 4 + 2 * bit;
        }
        out[Wp + y] = (uint16_t)code;
    }
    if (H & 1) out[Wp + H] = 0;
}

int oct_lut_entries_host(const LevelGeom& L) { return oct_lut_entries(L); }

int octtree_smem_bytes(const OrbGeom& g, int level, int key_smem_cap) {
    return oct_layout(g.lv[level].node_cap, key_smem_cap, oct_lut_entries(g.lv[level]), g.lv[level].n_ini).total;
}

// Warp-aggregated shared-memory counter increment: lanes that hit the same counter elect one leader, so the early
// quadtree passes (thousands of keys on a handful of counters) do not serialise on the shared-memory atomic unit.
__device__ __forceinline__ void count_add(int* counters, int idx) {
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, idx);
    if (idx >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counters[idx], __popc(peers));
}

__device__ __forceinline__ int quadrant_of(uint32_t xy, short4 b) {
    const int x = xy & 0xffff, y = xy >> 16;
    const int xm = b.x + ((b.z - b.x + 1) >> 1);  // UL.x + ceil((UR.x-UL.x)/2)   (:483)
    const int ym = b.y + ((b.w - b.y + 1) >> 1);
    return (x < xm ? 0 : 1) + (y < ym ? 0 : 2);   // n1,n2,n3,n4 (:511-524)
}
__device__ __forceinline__ short4 child_bounds(short4 b, int q) {
    const short xm = (short)(b.x + ((b.z - b.x + 1) >> 1));
    const short ym = (short)(b.y + ((b.w - b.y + 1) >> 1));
    short4 c;
    c.x = (q & 1) ? xm : b.x;
    c.z = (q & 1) ? b.z : xm;
    c.y = (q & 2) ? ym : b.y;
    c.w = (q & 2) ? b.w : ym;
    return c;
}

#ifdef CORB_OCT_TRACE
#define OCT_T(i) do { if (threadIdx.x == 0) oct_tr[i] = clock64(); } while (0)
#else
#define OCT_T(i) do { } while (0)
#endif
__global__ void __launch_bounds__(kOctThreads) k_octtree(OrbGeom g, OrbBuffers b0, int key_smem_cap, int level_begin, OrbBuffers b1) {
    const OrbBuffers b = blockIdx.z ? b1 : b0;
    extern __shared__ __align__(16) uint8_t smem[];
#ifdef CORB_OCT_TRACE
    long long oct_tr[24];
    for (int i = 0; i < 24; i++) oct_tr[i] = 0;
#endif
    TL_SCOPE(32 + blockIdx.x + level_begin);
    OCT_T(0);
    pdl_wait();  // the candidates come from the FAST launch right before this one in the stream
    const int l = blockIdx.x + level_begin;
    const LevelGeom L = g.lv[l];
    const int NC = L.node_cap, N = L.quota;
    const int tid = threadIdx.x, nt = blockDim.x;
    const OctLayout lay = oct_layout(NC, key_smem_cap, oct_lut_entries(L), L.n_ini);
    int* cnt_cur = reinterpret_cast<int*>(smem + lay.cntA);
    int* cnt_nxt = reinterpret_cast<int*>(smem + lay.cntB);
    short4* bnd_cur = reinterpret_cast<short4*>(smem + lay.bndA);
    short4* bnd_nxt = reinterpret_cast<short4*>(smem + lay.bndB);
    int* cc = reinterpret_cast<int*>(smem + lay.cc);
    int* procpos = reinterpret_cast<int*>(smem + lay.procpos);
    int* scan = reinterpret_cast<int*>(smem + lay.scan);
    int* crank = reinterpret_cast<int*>(smem + lay.crank);
    int* krank = reinterpret_cast<int*>(smem + lay.krank);
    uint8_t* fresh_cur = smem + lay.fresh;
    uint8_t* fresh_nxt = fresh_cur + ((NC + 3) & ~3);
    uint8_t* candf = smem + lay.candf;
    int* warp_tmp = reinterpret_cast<int*>(smem + lay.warp_tmp);
    volatile int* sh = reinterpret_cast<int*>(smem + lay.sh);
    long long* scan64 = reinterpret_cast<long long*>(smem + lay.scan64);
    long long* warp_tmp64 = reinterpret_cast<long long*>(smem + lay.warp_tmp64);

    // ---- the level's candidates: FAST appended them to the level's list (any order; the reference order travels as
    //      the order key in the low 24 bits of kro). The first batches are fetched before the count is known, so the
    //      count, the records and the path tables all arrive in one round of global latency.
    const int lut_n = oct_lut_entries(L);
    const bool fast_try = lay.fast && b.oct_fast;
    uint32_t* kxy = reinterpret_cast<uint32_t*>(smem + lay.keys_xy);
    uint16_t* knode = reinterpret_cast<uint16_t*>(smem + lay.keys_node);
    uint32_t* kro = reinterpret_cast<uint32_t*>(smem + lay.keys_ro);
    const uint32_t* gxy = b.cand_xy + L.cand_base;
    const uint32_t* gro = b.cand_ro + L.cand_base;
    const int cap_level = L.n_cols * L.n_rows * L.slot;
    int M;
    {
        uint32_t vx[4], vr[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = min(tid + u * nt, cap_level - 1);
            vx[u] = gxy[k];
            vr[u] = gro[k];
        }
        if (fast_try) {
            const uint32_t* glut = reinterpret_cast<const uint32_t*>(b.oct_lut + L.lut_off);
            uint32_t* slut = reinterpret_cast<uint32_t*>(smem + lay.f_lut);
            for (int i = tid; i < lut_n / 2; i += nt) slut[i] = glut[i];
        }
        M = b.level_cand[l];
        if (M <= key_smem_cap) {
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (tid + u * nt < M) { kxy[tid + u * nt] = vx[u]; kro[tid + u * nt] = vr[u]; }
            for (int k = tid + 4 * nt; k < M; k += nt) { kxy[k] = gxy[k]; kro[k] = gro[k]; }
        } else {  // too many for shared memory: the generic passes read the records in place
            kxy = const_cast<uint32_t*>(gxy);
            kro = const_cast<uint32_t*>(gro);
            knode = b.key_node + L.cand_base;
        }
    }
    __syncthreads();
    if (tid == 0) {
        b.level_cand[l] = 0;  // every thread has read it: ready for the next frame
        b.level_cand_out[l] = M;
    }
    OCT_T(2);
    // ---- closed-form path. The split geometry depends only on the root rectangle, so a key's path (root, q1, q2, ...)
    //      is a pure function of its coordinates and every node count is a range sum over one histogram of the keys'
    //      depth-kOctDh cells (exclusive scan P: count(d, c) = P[(c+1) << 2(Dh-d)] - P[c << 2(Dh-d)]). Full passes
    //      (:613-655) split every node with more than one key, hence after pass d the list holds one node per
    //      non-empty depth-d cell: s_d = #non-empty cells, nToExpand_d = #cells with more than one key, and the pass at
    //      which the loop stops or enters the near-quota phase follows from those two numbers per depth. The list after
    //      pass d is R_d ++ singles(R_{d-1}) ++ ... ++ singles(R_0), R_d = depth-d children of split nodes; push_front
    //      in creation order reverses the order each pass, which makes R_d sorted by the cell index with every other
    //      digit complemented (tau below). The near-quota rounds then run on nodes only (child counts are range sums),
    //      and keys find their final node through a dense (depth, cell) -> position map. Anything that needs a depth
    //      beyond kOctDh (sparse levels, tight clusters) falls through to the generic pass-by-pass code below.
    bool done = false;
    int s = 0;
    if (lay.fast && b.oct_fast && M > 0 && M <= key_smem_cap) {
        int* hist = reinterpret_cast<int*>(smem + lay.f_hist);
        int* fflag = reinterpret_cast<int*>(smem + lay.f_flag);
        int* fstat = reinterpret_cast<int*>(smem + lay.f_stat);
        uint16_t* kcode = reinterpret_cast<uint16_t*>(smem + lay.f_code);
        uint16_t* owner = reinterpret_cast<uint16_t*>(smem + lay.f_owner);
        int* fc_cur = reinterpret_cast<int*>(bnd_cur);  // node = depth << 24 | cell
        int* fc_nxt = reinterpret_cast<int*>(bnd_nxt);
        const int nini = L.n_ini;
        const int HC = nini << (2 * kOctDh);
        for (int i = tid; i < (HC + 4) / 4; i += nt) reinterpret_cast<int4*>(hist)[i] = make_int4(0, 0, 0, 0);
        if (tid < 16) fstat[tid] = 0;
        __syncthreads();
        {
            const uint16_t* xs = reinterpret_cast<const uint16_t*>(smem + lay.f_lut);
            const uint16_t* ys = xs + ((L.max_bx - kBorder + 1) & ~1);
            for (int k = tid; k < M; k += nt) {
                const uint32_t xy = kxy[k];
                const int code = xs[xy & 0xffff] | ys[xy >> 16];
                kcode[k] = (uint16_t)code;
                atomicAdd(&hist[code], 1);
            }
        }
        __syncthreads();
    OCT_T(3);
        block_excl_scan4(hist, HC, warp_tmp);
        if (tid == 0) hist[HC] = M;
        __syncthreads();
    OCT_T(4);
        auto hcount = [&](int d, int c) {
            const int sh2 = 2 * (kOctDh - d);
            return hist[(c + 1) << sh2] - hist[c << sh2];
        };
        // per depth: non-empty cells (low half) and cells holding more than one key (high half); depths 0..Dh-1 first,
        // the deepest tabulated depth only when the stop depth was not found among them
        auto depth_stats = [&](int d) {
            const int ncell = nini << (2 * d);
            int acc = 0;
            for (int c = tid; c < ncell; c += nt) {
                const int n = hcount(d, c);
                acc += (n > 0) + ((n > 1) << 16);
            }
            acc = __reduce_add_sync(0xffffffffu, acc);
            if ((tid & 31) == 0 && acc) {
                atomicAdd(&fstat[d], acc & 0xffff);
                atomicAdd(&fstat[8 + d], acc >> 16);
            }
        };
#pragma unroll
        for (int d = 0; d < kOctDh; d++) depth_stats(d);
        __syncthreads();
        int dstar = -1;
        bool final_phase = false;
        auto find_stop = [&](int dmax) {
            int sp = fstat[0];
            for (int d = 1; d <= dmax; d++) {
                const int sd = fstat[d];
                if (sd >= N || sd == sp) { dstar = d; return; }
                if (sd + 3 * fstat[8 + d] > N) { dstar = d; final_phase = true; return; }
                sp = sd;
            }
        };
        find_stop(kOctDh - 1);
        if (dstar < 0) {
            depth_stats(kOctDh);
            __syncthreads();
            find_stop(kOctDh);
        }
    OCT_T(5);
        if (dstar > 0 && fstat[dstar] <= NC) {
            // membership flags over the concatenated segments [depth dstar | singles of depth dstar-1 | ... | depth 0]
            const int etot = oct_depth_off(nini, dstar + 1);
            auto entry = [&](int e, int* d_out, int* c_out, int* n_out) -> bool {
                int d = dstar, base = 0;
                while (e >= base + (nini << (2 * d))) { base += nini << (2 * d); d--; }
                const int i = e - base;
                const int lowmask = (1 << (2 * d)) - 1;
                int r = i >> (2 * d);
                if (d & 1) r = nini - 1 - r;
                const int c = (r << (2 * d)) | ((i & lowmask) ^ (0x33333333 & lowmask));  // tau_d (an involution)
                const int n = hcount(d, c);
                const bool par = d == 0 || hcount(d - 1, c >> 2) > 1;
                *d_out = d; *c_out = c; *n_out = n;
                return par && (d == dstar ? n > 0 : n == 1);
            };
            for (int e = tid; e < etot; e += nt) {
                int d, c, n;
                fflag[e] = entry(e, &d, &c, &n) ? 1 : 0;
            }
            __syncthreads();
            s = block_excl_scan4(fflag, etot, warp_tmp);
    OCT_T(6);
            if (tid == 0) fflag[etot] = s;
            __syncthreads();
            for (int e = tid; e < etot; e += nt) {
                const int pos = fflag[e];
                if (fflag[e + 1] != pos) {
                    int d, c, n;
                    entry(e, &d, &c, &n);
                    fc_cur[pos] = d << 24 | c;
                    cnt_cur[pos] = n;
                    fresh_cur[pos] = d == dstar;
                }
            }
            __syncthreads();
    OCT_T(7);
            bool bail = false;
            int dmax = dstar;  // deepest node depth in the list
            while (final_phase) {  // near-quota rounds (:660-733) on nodes only
                const int prev = s;
                for (int i = tid; i < s; i += nt) {
                    const int f = fresh_cur[i] && cnt_cur[i] > 1;
                    scan[i] = f;
                    candf[i] = (uint8_t)f;
                }
                if (tid == 0) sh[3] = 0;
                __syncthreads();
                const int m = block_excl_scan4(scan, s, warp_tmp);
    OCT_T(8);
                for (int i = tid; i < s; i += nt) {
                    if (candf[i]) {
                        krank[scan[i]] = i;  // candidates in position order
                        const int fc = fc_cur[i], d = fc >> 24, c = fc & 0xffffff;
                        if (d + 1 > kOctDh) {
                            sh[3] = 1;
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; q++) cc[i * 4 + q] = hcount(d + 1, c * 4 + q);
                        }
                    }
                    crank[i] = -1;
                }
                __syncthreads();
                if (sh[3]) { bail = true; break; }
    OCT_T(9);
                for (int i = tid; i < m; i += nt) scan[i] = cnt_cur[krank[i]];  // candidate counts, contiguous
                if (tid == 0) { sh[0] = m; sh[1] = 0; }
                __syncthreads();
                // processing order: count descending, position ascending = a stable counting sort on the count.
                // Per-warp bins (match_any gives the rank among equal counts inside a warp), a prefix over the warps
                // per bin, and a scan over the bins from the largest count down. Counts >= 256 or more than one
                // candidate per thread (never seen on images) take the quadratic ranking instead.
                const int cmax = __syncthreads_or(tid < m && scan[tid] >= 256);
                if (!cmax && m <= nt) {
                    int* wh = fflag;              // [nw][256] per-warp bin counts -> exclusive prefix over warps
                    int* tot = fflag + (nt >> 5) * 256;  // [256] reversed bin totals -> exclusive scan
                    const int nwarp = nt >> 5, wid = tid >> 5, lane = tid & 31;
                    for (int i = tid; i < (nwarp + 1) * 64; i += nt) reinterpret_cast<int4*>(wh)[i] = make_int4(0, 0, 0, 0);
                    __syncthreads();
                    const int ci = tid < m ? scan[tid] : -1;
                    const unsigned peers = __match_any_sync(0xffffffffu, ci);
                    const int in_warp = __popc(peers & ((1u << lane) - 1u));
                    if (ci >= 0 && in_warp == 0) wh[wid * 256 + ci] = __popc(peers);
                    __syncthreads();
                    if (tid < 256) {
                        int acc = 0;
                        for (int w = 0; w < nwarp; w++) {
                            const int t = wh[w * 256 + tid];
                            wh[w * 256 + tid] = acc;
                            acc += t;
                        }
                        tot[255 - tid] = acc;
                    }
                    __syncthreads();
                    block_excl_scan4(tot, 256, warp_tmp);
                    if (ci >= 0) procpos[tot[255 - ci] + wh[wid * 256 + ci] + in_warp] = krank[tid];
                } else {
                    for (int i = tid; i < m; i += nt) {
                        const int ci = scan[i];
                        int rank = 0;
                        for (int j = 0; j < m; j++) {
                            const int cj = scan[j];
                            rank += (cj > ci) || (cj == ci && j < i);
                        }
                        procpos[rank] = krank[i];
                    }
                }
                __syncthreads();
    OCT_T(10);
                for (int r = tid; r < m; r += nt) {
                    const int* c4 = cc + procpos[r] * 4;
                    scan[r] = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0) - 1;
                }
                __syncthreads();
                block_excl_scan4(scan, m, warp_tmp);
                for (int r = tid; r < m; r += nt) {  // first prefix whose list size reaches N (:728-729)
                    const int* c4 = cc + procpos[r] * 4;
                    const int inc = scan[r] + (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0) - 1;
                    if (s + inc >= N && s + scan[r] < N) sh[0] = r + 1;
                }
                __syncthreads();
                const int mproc = sh[0];
    OCT_T(11);
                for (int r = tid; r < mproc; r += nt) {
                    const int nd = procpos[r];
                    crank[nd] = scan[r] + r;
                    if (r == mproc - 1) {
                        const int* c4 = cc + nd * 4;
                        sh[1] = scan[r] + r + (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
                    }
                }
                __syncthreads();
                const int T = sh[1];
                for (int i = tid; i < s; i += nt) scan[i] = crank[i] < 0;
                __syncthreads();
                const int K = block_excl_scan4(scan, s, warp_tmp);
                if (T + K > NC) { bail = true; break; }  // cannot happen (list size <= max(N + 3, 4 nIni))
    OCT_T(12);
                for (int i = tid; i < s; i += nt) {
                    if (crank[i] >= 0) {
                        const int fc = fc_cur[i], d = fc >> 24, c = fc & 0xffffff;
                        int idx = crank[i];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int n = cc[i * 4 + q];
                            if (n > 0) {
                                const int np = T - 1 - idx;
                                cnt_nxt[np] = n;
                                fc_nxt[np] = (d + 1) << 24 | (c * 4 + q);
                                fresh_nxt[np] = 1;
                                idx++;
                            }
                        }
                    } else {
                        const int np = T + scan[i];
                        cnt_nxt[np] = cnt_cur[i];
                        fc_nxt[np] = fc_cur[i];
                        fresh_nxt[np] = 0;
                    }
                }
                __syncthreads();
                { int* t = cnt_cur; cnt_cur = cnt_nxt; cnt_nxt = t; }
                { int* t = fc_cur; fc_cur = fc_nxt; fc_nxt = t; }
                { uint8_t* t = fresh_cur; fresh_cur = fresh_nxt; fresh_nxt = t; }
                s = T + K;
                dmax++;
    OCT_T(13);
                if (s >= N || s == prev) break;
            }
            if (!bail) {
                // every final node writes its list position over the depth-Dh cells it covers (aligned runs of
                // 4^(Dh - depth) entries), so a key finds its node with one lookup
                for (int i = tid; i < s; i += nt) {
                    const int fc = fc_cur[i], d = fc >> 24, c = fc & 0xffffff;
                    const int run = 1 << (2 * (kOctDh - d));
                    uint16_t* o = owner + (size_t)c * run;
                    if (run >= 8) {
                        const uint32_t w = (uint32_t)i | (uint32_t)i << 16;
                        const uint4 v = make_uint4(w, w, w, w);
                        for (int j = 0; j < run; j += 8) *reinterpret_cast<uint4*>(o + j) = v;
                    } else {
                        for (int j = 0; j < run; j++) o[j] = (uint16_t)i;
                    }
                }
                __syncthreads();
                for (int k = tid; k < M; k += nt) knode[k] = owner[kcode[k]];
                done = true;
    OCT_T(14);
            }
            // restore the canonical buffer roles for the generic path / epilogue (pointers only; contents are rebuilt)
            cnt_cur = reinterpret_cast<int*>(smem + lay.cntA);
            cnt_nxt = reinterpret_cast<int*>(smem + lay.cntB);
            fresh_cur = smem + lay.fresh;
            fresh_nxt = fresh_cur + ((NC + 3) & ~3);
        }
        __syncthreads();
    }
    if (!done) {
    OCT_T(15);
    // ---- initial nodes (:542-589)
    for (int i = tid; i < 4 * NC; i += nt) cc[i] = 0;
    __syncthreads();
    for (int k = tid; k < M; k += nt) {
        const int ni = (int)__fdiv_rn((float)(kxy[k] & 0xffff), L.h_x);
        knode[k] = (uint16_t)ni;
        count_add(cc, ni);
    }
    __syncthreads();
    for (int i = tid; i < L.n_ini; i += nt) scan[i] = cc[i] > 0;
    __syncthreads();
    s = block_excl_scan(scan, L.n_ini, warp_tmp);
    for (int i = tid; i < L.n_ini; i += nt) {
        if (cc[i] > 0) {
            const int p = scan[i];
            cnt_cur[p] = cc[i];
            short4 bb;
            bb.x = (short)(int)__fmul_rn(L.h_x, (float)i);
            bb.z = (short)(int)__fmul_rn(L.h_x, (float)(i + 1));
            bb.y = 0;
            bb.w = (short)(L.max_by - kBorder);
            bnd_cur[p] = bb;
            fresh_cur[p] = 0;
        }
    }
    for (int k = tid; k < M; k += nt) knode[k] = (uint16_t)scan[knode[k]];
    __syncthreads();

    bool finish = false, final_phase = false;
    int guard = 0;
    while (!finish) {
        if (++guard > 4096) {
            if (tid == 0) atomicExch(b.status, 102);
            break;
        }
        const int prev = s;
        if (!final_phase) {
            // ---- full pass (:613-655): every node holding more than one key splits, in list order, no early stop.
            //      One packed scan gives both the creation index of each parent's children and the rank of kept nodes.
            for (int i = tid; i < 4 * s; i += nt) cc[i] = 0;
            if (tid == 0) sh[2] = 0;
            __syncthreads();
            for (int k = tid; k < M; k += nt) {
                const int nd = knode[k];
                count_add(cc, cnt_cur[nd] > 1 ? nd * 4 + quadrant_of(kxy[k], bnd_cur[nd]) : -1);
            }
            __syncthreads();
            for (int i = tid; i < s; i += nt) {
                const int* c4 = cc + i * 4;
                const bool ex = cnt_cur[i] > 1;
                const int nch = ex ? (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0) : 0;
                scan64[i] = ((long long)nch << 32) | (ex ? 0 : 1);
            }
            __syncthreads();
            const long long tot = block_excl_scan64(scan64, s, warp_tmp64);
            const int T = (int)(tot >> 32), K = (int)(tot & 0xffffffffLL);
            if (T + K > NC) {
                if (tid == 0) atomicExch(b.status, 103);
                s = 0;
                break;
            }
            int n_expand = 0;
            for (int i = tid; i < s; i += nt) {
                const long long sc = scan64[i];
                if (cnt_cur[i] > 1) {
                    const short4 pb = bnd_cur[i];
                    int idx = (int)(sc >> 32);
                    crank[i] = idx;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int c = cc[i * 4 + q];
                        if (c > 0) {
                            const int np = T - 1 - idx;
                            cnt_nxt[np] = c;
                            bnd_nxt[np] = child_bounds(pb, q);
                            fresh_nxt[np] = 1;
                            n_expand += c > 1;
                            idx++;
                        }
                    }
                } else {
                    const int np = T + (int)(sc & 0xffffffffLL);
                    cnt_nxt[np] = cnt_cur[i];
                    bnd_nxt[np] = bnd_cur[i];
                    fresh_nxt[np] = 0;
                    krank[i] = np;
                }
            }
            if (n_expand) atomicAdd((int*)&sh[2], n_expand);
            __syncthreads();
            for (int k = tid; k < M; k += nt) {
                const int nd = knode[k];
                int np;
                if (cnt_cur[nd] > 1) {
                    const int q = quadrant_of(kxy[k], bnd_cur[nd]);
                    const int* c4 = cc + nd * 4;
                    int rank = 0;
                    if (q > 0) rank += c4[0] > 0;
                    if (q > 1) rank += c4[1] > 0;
                    if (q > 2) rank += c4[2] > 0;
                    np = T - 1 - (crank[nd] + rank);
                } else {
                    np = krank[nd];
                }
                knode[k] = (uint16_t)np;
            }
            __syncthreads();
            const int n_to_expand = sh[2];
            { int* t = cnt_cur; cnt_cur = cnt_nxt; cnt_nxt = t; }
            { short4* t = bnd_cur; bnd_cur = bnd_nxt; bnd_nxt = t; }
            { uint8_t* t = fresh_cur; fresh_cur = fresh_nxt; fresh_nxt = t; }
            s = T + K;
            __syncthreads();
            if (s >= N || s == prev) finish = true;
            else if (s + 3 * n_to_expand > N) final_phase = true;
            continue;
        }
        // ---- near-quota phase (:660-733): a step with a sorted processing order and an early stop (full passes `continue` above)
        int m;
        for (int i = tid; i < s; i += nt) {
            const int f = fresh_cur[i] && cnt_cur[i] > 1;
            scan[i] = f;
            candf[i] = (uint8_t)f;
        }
        __syncthreads();
        m = block_excl_scan(scan, s, warp_tmp);
        for (int i = tid; i < s; i += nt)
            if (candf[i]) krank[scan[i]] = i;  // candidates in position order (temporary)
        __syncthreads();
        for (int i = tid; i < m; i += nt) crank[i] = cnt_cur[krank[i]];  // candidate counts, contiguous (temporary)
        __syncthreads();
        for (int i = tid; i < m; i += nt) {
            const int ci = crank[i];
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < m; j++) {
                const int cj = crank[j];
                rank += (cj > ci) || (cj == ci && j < i);
            }
            procpos[rank] = krank[i];
        }
        __syncthreads();
        for (int i = tid; i < 4 * s; i += nt) cc[i] = 0;
        for (int i = tid; i < s; i += nt) crank[i] = -1;
        if (tid == 0) { sh[0] = m; sh[1] = 0; sh[2] = 0; }
        __syncthreads();
        // ---- keys per (node, quadrant)
        for (int k = tid; k < M; k += nt) {
            const int nd = knode[k];
            count_add(cc, candf[nd] ? nd * 4 + quadrant_of(kxy[k], bnd_cur[nd]) : -1);
        }
        __syncthreads();
        for (int r = tid; r < m; r += nt) {
            const int* c4 = cc + procpos[r] * 4;
            scan[r] = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0) - 1;
        }
        __syncthreads();
        block_excl_scan(scan, m, warp_tmp);
        if (final_phase) {  // first prefix whose list size reaches N (:728-729)
            for (int r = tid; r < m; r += nt) {
                const int* c4 = cc + procpos[r] * 4;
                const int inc = scan[r] + (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0) - 1;
                if (s + inc >= N && s + scan[r] < N) sh[0] = r + 1;
            }
            __syncthreads();
        }
        const int mproc = sh[0];
        for (int r = tid; r < mproc; r += nt) {
            const int nd = procpos[r];
            crank[nd] = scan[r] + r;
            if (r == mproc - 1) {
                const int* c4 = cc + nd * 4;
                sh[1] = scan[r] + r + (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
            }
        }
        __syncthreads();
        const int T = sh[1];
        for (int i = tid; i < s; i += nt) scan[i] = crank[i] < 0;
        __syncthreads();
        const int K = block_excl_scan(scan, s, warp_tmp);
        if (T + K > NC) {  // cannot happen (list size is bounded by max(N + 2, 4 nIni)); never write out of bounds
            if (tid == 0) atomicExch(b.status, 103);
            s = 0;
            break;
        }
        // ---- new list: children in reverse creation order, then the unprocessed nodes in their old order
        int n_expand = 0;
        for (int i = tid; i < s; i += nt) {
            if (crank[i] >= 0) {
                const short4 pb = bnd_cur[i];
                int idx = crank[i];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = cc[i * 4 + q];
                    if (c > 0) {
                        const int np = T - 1 - idx;
                        cnt_nxt[np] = c;
                        bnd_nxt[np] = child_bounds(pb, q);
                        fresh_nxt[np] = 1;
                        n_expand += c > 1;
                        idx++;
                    }
                }
            } else {
                const int np = T + scan[i];
                cnt_nxt[np] = cnt_cur[i];
                bnd_nxt[np] = bnd_cur[i];
                fresh_nxt[np] = 0;
                krank[i] = np;
            }
        }
        if (n_expand) atomicAdd((int*)&sh[2], n_expand);
        __syncthreads();
        for (int k = tid; k < M; k += nt) {
            const int nd = knode[k];
            int np;
            if (crank[nd] >= 0) {
                const int q = quadrant_of(kxy[k], bnd_cur[nd]);
                const int* c4 = cc + nd * 4;
                int rank = 0;
                if (q > 0) rank += c4[0] > 0;
                if (q > 1) rank += c4[1] > 0;
                if (q > 2) rank += c4[2] > 0;
                np = T - 1 - (crank[nd] + rank);
            } else {
                np = krank[nd];
            }
            knode[k] = (uint16_t)np;
        }
        __syncthreads();
        const int n_to_expand = sh[2];
        { int* t = cnt_cur; cnt_cur = cnt_nxt; cnt_nxt = t; }
        { short4* t = bnd_cur; bnd_cur = bnd_nxt; bnd_nxt = t; }
        { uint8_t* t = fresh_cur; fresh_cur = fresh_nxt; fresh_nxt = t; }
        s = T + K;
        __syncthreads();
        // ---- termination (:647-733)
        if (s >= N || s == prev) finish = true;
        else if (!final_phase && s + 3 * n_to_expand > N) final_phase = true;
    }
    }
    // ---- keep the max-response key of every node, first in candidate order on ties (:738-757)
    uint32_t* best = reinterpret_cast<uint32_t*>(cc);
    OCT_T(16);
    for (int i = tid; i < s; i += nt) best[i] = 0;
    __syncthreads();
    // kro = response << 24 | (0xffffff - order key): the maximum is the largest response, earliest candidate on ties
    for (int k = tid; k < M; k += nt) atomicMax(&best[knode[k]], kro[k]);
    __syncthreads();
    for (int k = tid; k < M; k += nt) {  // order keys are unique, so exactly one key per node matches
        const int nd = knode[k];
        const uint32_t v = kro[k];
        if (best[nd] == v) {
            const uint32_t xy = kxy[k];
            b.lvl_kp[L.kp_base + nd] = make_uint2(((xy & 0xffff) + kBorder) | ((xy >> 16) + kBorder) << 16, v >> 24);
        }
    }
    if (tid == 0) b.level_count[l] = s;
#ifdef CORB_OCT_TRACE
    OCT_T(17);
    if (threadIdx.x == 0 && blockIdx.x + level_begin == 0 && blockIdx.z == 0) {
        printf("oct L0 M=%d s=%d done=%d :", M, s, (int)done);
        long long last = oct_tr[0];
        for (int i = 1; i < 18; i++) if (oct_tr[i]) { printf(" t%d+%lld", i, oct_tr[i] - last); last = oct_tr[i]; }
        printf(" total %lld\n", oct_tr[17] - oct_tr[0]);
    }
#endif
}

cudaError_t prepare_octtree(const OrbGeom& g, int key_smem_cap, int* smem_bytes_out) {
    int bytes = 0;
    for (int l = 0; l < g.n_levels; l++) bytes = max(bytes, octtree_smem_bytes(g, l, key_smem_cap));
    *smem_bytes_out = bytes;
    return cudaFuncSetAttribute(k_octtree, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

void launch_octtree(const OrbGeom& g, const OrbBuffers& b, int level, int key_smem_cap, int smem_bytes, cudaStream_t s,
                    const OrbBuffers* b1) {
    const int nz = b1 ? 2 : 1;
    if (level < 0) launch_chain(k_octtree, dim3(g.n_levels, 1, nz), dim3(kOctThreads), (size_t)smem_bytes, s, g, b, key_smem_cap, 0, b1 ? *b1 : b);
    else launch_chain(k_octtree, dim3(1, 1, nz), dim3(kOctThreads), (size_t)octtree_smem_bytes(g, level, key_smem_cap), s, g, b, key_smem_cap,
                      level, b1 ? *b1 : b);
}

// ------------------------------------------------------------------------------------------------ K4 + K6
// fastAtan2 (OpenCV scalar atan_f32): float32, evaluated without FMA contraction.
__device__ __forceinline__ float fast_atan2_dev(float y, float x) {
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;  // folded at compile time in float, like the host compiler does
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// One warp per kept keypoint: IC_Angle on the un-blurred level (ORBextractor.cc:77-104), then the 256 steered BRIEF
// tests on the blurred level (:107-146), then the epilogue of operator() (:1092-1101, :837-847).
__global__ void __launch_bounds__(256) k_orient_desc(OrbGeom g, OrbBuffers b0, OrbBuffers b1) {
    const OrbBuffers b = blockIdx.z ? b1 : b0;
    // pattern transposed to [point-in-byte 0..15][lane 0..31] so the 32 lanes of a warp read 64 contiguous bytes
    __shared__ char2 pat[16 * 32];
    TL_SCOPE(49);
    for (int i = threadIdx.x; i < 512; i += 256)
        pat[(i & 15) * 32 + (i >> 4)] = make_char2(d_pattern[2 * i], d_pattern[2 * i + 1]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wslot = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wslot == 0 && lane == 0) {
        int tot = 0;
        for (int l = 0; l < g.n_levels; l++) tot += b.level_count[l];
        *b.count = tot;
    }
    if (wslot >= g.kp_cap) return;
    int l = 0, out_base = 0;
    while (l + 1 < g.n_levels && wslot >= g.lv[l + 1].kp_base) {
        out_base += b.level_count[l];
        l++;
    }
    const LevelGeom L = g.lv[l];
    const int pos = wslot - L.kp_base;
    if (pos >= b.level_count[l]) return;
    const uint2 kp = b.lvl_kp[wslot];
    const int X = kp.x & 0xffff, Y = kp.x >> 16;
    // ---- orientation: lane = column u in [-15,15]; pixel (u,v) is in the disc iff |u| <= umax[|v|]
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - kHalfPatch;
        const uint8_t* c = b.pyr + L.img_off + (size_t)Y * L.pitch + X + u;
        const int au = abs(u);
#pragma unroll
        for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
            if (au <= d_umax[v < 0 ? -v : v]) {
                const int val = c[v * L.pitch];
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_dev((float)m01, (float)m10);
    // ---- descriptor: lane = output byte; contract: cos/sin in double, rounded to float
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    double sd, cd;
    sincos((double)rad, &sd, &cd);
    const float ca = (float)cd, sa = (float)sd;
    const uint8_t* cb = b.blur + L.img_off + (size_t)Y * L.pitch + X;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int t[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const char2 pt = pat[(2 * k + j) * 32 + lane];
            const float px = (float)pt.x, py = (float)pt.y;
            const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, sa), __fmul_rn(py, ca)));
            const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, ca), __fmul_rn(py, sa)));
            t[j] = cb[iy * L.pitch + ix];
        }
        val |= (t[0] < t[1]) << k;
    }
    const int out = out_base + pos;
    b.desc[(size_t)out * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        corb_keypoint o;
        o.x = (float)X;
        o.y = (float)Y;
        if (l != 0) {
            o.x = __fmul_rn(o.x, L.scale);
            o.y = __fmul_rn(o.y, L.scale);
        }
        o.size = L.size;
        o.angle = angle;
        o.response = (float)kp.y;
        o.octave = l;
        o.class_id = -1;
        b.kps[out] = o;
    }
}

void launch_orient_desc(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1) {
    k_orient_desc<<<dim3((g.kp_cap + 7) / 8, 1, b1 ? 2 : 1), 256, 0, s>>>(g, b, b1 ? *b1 : b);
}

}  // namespace corb
