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

__device__ __align__(16) const int8_t d_pattern[1024] = {
#include "../../include/corb_brief_pattern.inc"
};
// umax of ORBextractor.cc:452-469 for HALF_PATCH_SIZE = 15 (the host recomputes it and checks equality)
__device__ const int d_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// ------------------------------------------------------------------------------------------------ K0 import
// Copies the caller's image (device memory, or page-locked host memory read over PCIe through its UVA mapping) into
// level 0 of the pitched pyramid. First node of the per-frame graph, so one cudaGraphLaunch is the only driver call
// an extraction needs; its (src, stride) arguments are patched per launch with cudaGraphExecKernelNodeSetParams.
__global__ void __launch_bounds__(256) k_import(const uint8_t* __restrict__ src, int stride, uint8_t* __restrict__ dst, int pitch,
                                                int w, int h) {
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

void launch_import(const OrbGeom& g, const OrbBuffers& b, const uint8_t* src, int stride, cudaStream_t s) {
    const LevelGeom& L0 = g.lv[0];
    dim3 grid((L0.w + 1023) / 1024, L0.h);
    k_import<<<grid, 256, 0, s>>>(src, stride, b.pyr + L0.img_off, L0.pitch, L0.w, L0.h);
}
const void* import_kernel_ptr() { return (const void*)k_import; }

// ------------------------------------------------------------------------------------------------ K1 resize
// cv::resize INTER_LINEAR u8 (ORBextractor.cc:1120). Coefficient tables are built on the host exactly as OpenCV
// builds them (double -> float -> short), so the kernel is pure integer arithmetic.
__global__ void __launch_bounds__(256) k_resize(LevelGeom src, LevelGeom dst, const uint8_t* __restrict__ pyr_src,
                                                uint8_t* __restrict__ pyr_dst, const int* __restrict__ xofs,
                                                const short2* __restrict__ alpha, const int* __restrict__ yofs,
                                                const short2* __restrict__ beta) {
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int dy = blockIdx.y * 8 + threadIdx.y;
    if (dy >= dst.h || dx0 >= dst.w) return;
    const int sy = yofs[dy];
    const int sy0 = min(max(sy, 0), src.h - 1), sy1 = min(max(sy + 1, 0), src.h - 1);
    const short2 bb = beta[dy];
    const uint8_t* s0 = pyr_src + (size_t)sy0 * src.pitch;
    const uint8_t* s1 = pyr_src + (size_t)sy1 * src.pitch;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int dx = dx0 + k;
        if (dx < dst.w) {
            const int sx = xofs[dx];
            const int sx1 = min(sx + 1, src.w - 1);
            const short2 a = alpha[dx];
            const int r0 = (int)s0[sx] * a.x + (int)s0[sx1] * a.y;
            const int r1 = (int)s1[sx] * a.x + (int)s1[sx1] * a.y;
            const int v = ((((int)bb.x * (r0 >> 4)) >> 16) + (((int)bb.y * (r1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)(v & 0xff) << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t*>(pyr_dst + (size_t)dy * dst.pitch + dx0) = packed;  // pitch % 128 == 0, dx0 % 4 == 0
}

void launch_resize(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s) {
    const LevelGeom& src = g.lv[level - 1];
    const LevelGeom& dst = g.lv[level];
    dim3 block(32, 8), grid((dst.w + 127) / 128, (dst.h + 7) / 8);
    k_resize<<<grid, block, 0, s>>>(src, dst, b.pyr + src.img_off, b.pyr + dst.img_off, b.xofs + dst.xtab_off,
                                    b.alpha + dst.xtab_off, b.yofs + dst.ytab_off, b.beta + dst.ytab_off);
}

// ------------------------------------------------------------------------------------------------ K2 FAST per cell
constexpr int kRoiPitch = 96;  // bytes per ROI row in shared memory = TMA box width: multiple of 16 and >= kCellRoiMax + 15,
                               // because the innermost TMA coordinate must be 16-byte aligned (measured: any other x traps
                               // with 'illegal instruction' on sm_100a), so the box starts at iniX & ~15
constexpr int kRoiPitchRaw = kRoiPitch;
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
    r[0] = c[3 * P]; r[4] = c[3]; r[8] = c[-3 * P]; r[12] = c[-3];
    const int hi0 = v + th, lo0 = v - th;
    if ((r[0] > hi0) + (r[4] > hi0) + (r[8] > hi0) + (r[12] > hi0) < 2 && (r[0] < lo0) + (r[4] < lo0) + (r[8] < lo0) + (r[12] < lo0) < 2)
        return 0;
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
__global__ void __launch_bounds__(256) k_fast_cells(OrbGeom g, const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                    const uint8_t* __restrict__ pyr, int* __restrict__ cell_count,
                                                    uint32_t* __restrict__ cand_xy, uint8_t* __restrict__ cand_r,
                                                    int* __restrict__ status, int cell_begin) {
    __shared__ __align__(128) uint8_t roi_raw[kCellRoiMax * kRoiPitchRaw + 16];
    __shared__ __align__(8) uint64_t tma_bar;
    const uint8_t* roi;
    __shared__ __align__(16) uint8_t sc[kScDim * kScPitch];
    __shared__ uint32_t row_ini[128], row_min[128];  // two ballot words per valid row (<= 60 rows)
    __shared__ int row_off[65];
    const int tid = threadIdx.x;
    const int cell = blockIdx.x + cell_begin;
    int l = 0;
    while (l + 1 < g.n_levels && cell >= g.lv[l + 1].cell_base) l++;
    const LevelGeom L = g.lv[l];
    const int c = cell - L.cell_base;
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
            tma_load_2d(roi_raw, &tmap, iniX & ~15, iniY, &tma_bar);
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
    for (int i = tid; i < kScDim * kScPitch / 4; i += 256) reinterpret_cast<uint32_t*>(sc)[i] = 0;
    if (use_tma) mbar_wait(&tma_bar, 0);
    __syncthreads();
    // warp w owns rows w, w + 8, ...; lane = column (two halves: x = lane, lane + 32) -> no integer divisions, and the
    // raster order needed for the output falls out of ballots (x order) and a scan over rows (y order)
    const int lane = tid & 31, warp = tid >> 5;
    const int th_lo = min(g.ini_th, g.min_th);
    for (int y = warp; y < vh; y += 8) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int x = lane + 32 * half;
            if (x < vw) sc[(y + 1) * kScPitch + (x + 1)] = (uint8_t)fast_score_dev(roi + (y + 3) * kRoiPitch + (x + 3), th_lo);
        }
    }
    __syncthreads();
    // strict 8-neighbour maxima of the (masked) score tile; per row two ballot words for each threshold
    bool any_ini_local = false;
    for (int y = warp; y < vh; y += 8) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int x = lane + 32 * half;
            bool f_ini = false, f_min = false;
            if (x < vw) {
                const uint8_t* q = sc + (y + 1) * kScPitch + (x + 1);
                const int s = q[0];
                if (s != 0 && s > q[-kScPitch - 1] && s > q[-kScPitch] && s > q[-kScPitch + 1] && s > q[-1] && s > q[1] &&
                    s > q[kScPitch - 1] && s > q[kScPitch] && s > q[kScPitch + 1]) {
                    f_ini = s >= g.ini_th;
                    f_min = s >= g.min_th;
                }
            }
            const uint32_t b_ini = __ballot_sync(0xffffffffu, f_ini), b_min = __ballot_sync(0xffffffffu, f_min);
            if (lane == 0) { row_ini[y * 2 + half] = b_ini; row_min[y * 2 + half] = b_min; }
            any_ini_local |= b_ini != 0;
        }
    }
    const int any_ini = __syncthreads_or(any_ini_local);
    const uint32_t* rowm = any_ini ? row_ini : row_min;
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
        if (lane == 31) row_off[64] = inc;
    }
    __syncthreads();
    const int total = row_off[64];
    if (total > L.slot) {  // impossible for strict 8-neighbour maxima; never truncate silently
        if (tid == 0) { atomicExch(status, 101); cell_count[cell] = 0; }
        return;
    }
    const int base = L.cand_base + c * L.slot;
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
                cand_r[base + pos] = sc[(y + 1) * kScPitch + (x + 1)];
            }
        }
    }
    if (tid == 0) cell_count[cell] = total;
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
    }
    return true;
}

// level < 0: every level (one launch each); else only that level (lets a level start as soon as it is resized)
void launch_fast_cells(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s) {
    // one launch per level: the level's tensor map travels as a __grid_constant__ parameter (the form TMA expects)
    const int l0 = level < 0 ? 0 : level, l1 = level < 0 ? g.n_levels : level + 1;
    for (int l = l0; l < l1; l++) {
        const int n = g.lv[l].n_cols * g.lv[l].n_rows;
        k_fast_cells<<<n, 256, 0, s>>>(g, b.tma_maps->m[l], b.use_tma, b.pyr, b.cell_count, b.cand_xy, b.cand_r, b.status,
                                       g.lv[l].cell_base);
    }
}

// ------------------------------------------------------------------------------------------------ K5 Gaussian 7x7
// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on u8 (ORBextractor.cc:1086): fixed-point taps
// {18,34,48,56,48,34,18}/256 on both axes, exact accumulation, one rounding (s + 2^15) >> 16.
constexpr int kBlurTW = 64, kBlurTH = 16;

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256) k_blur(OrbGeom g, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
    __shared__ uint8_t in[kBlurTH + 6][kBlurTW + 8];
    __shared__ uint16_t hb[kBlurTH + 6][kBlurTW];
    const int tid = threadIdx.x;
    int l = 0;
    while (l + 1 < g.n_levels && (int)blockIdx.x >= g.lv[l + 1].blur_tile_base) l++;
    const LevelGeom L = g.lv[l];
    const int t = blockIdx.x - L.blur_tile_base;
    const int ty = t / L.blur_tiles_x, tx = t - ty * L.blur_tiles_x;
    const int x0 = tx * kBlurTW, y0 = ty * kBlurTH;
    const uint8_t* src = pyr + L.img_off;
    for (int i = tid; i < (kBlurTH + 6) * (kBlurTW + 6); i += 256) {
        const int yy = i / (kBlurTW + 6), xx = i - yy * (kBlurTW + 6);
        const int gy = reflect101(min(max(y0 + yy - 3, -3), L.h + 2), L.h);
        const int gx = reflect101(min(max(x0 + xx - 3, -3), L.w + 2), L.w);
        in[yy][xx] = src[(size_t)gy * L.pitch + gx];
    }
    __syncthreads();
    for (int i = tid; i < (kBlurTH + 6) * kBlurTW; i += 256) {
        const int yy = i / kBlurTW, xx = i - yy * kBlurTW;
        const uint8_t* p = &in[yy][xx];
        hb[yy][xx] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    const int lx = (tid & 15) * 4, ly = tid >> 4;
    const int gx = x0 + lx, gy = y0 + ly;
    if (gy < L.h && gx < L.w) {
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = lx + k;
            const uint32_t acc = 18u * ((uint32_t)hb[ly][x] + hb[ly + 6][x]) + 34u * ((uint32_t)hb[ly + 1][x] + hb[ly + 5][x]) +
                                 48u * ((uint32_t)hb[ly + 2][x] + hb[ly + 4][x]) + 56u * (uint32_t)hb[ly + 3][x];
            packed |= ((acc + 32768u) >> 16) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(blur + L.img_off + (size_t)gy * L.pitch + gx) = packed;
    }
}

void launch_blur(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s) {
    k_blur<<<g.blur_tiles, 256, 0, s>>>(g, b.pyr, b.blur);
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
#define CORB_OCT_THREADS 512
#endif
constexpr int kOctThreads = CORB_OCT_THREADS;

struct OctLayout {
    int cntA, cntB, bndA, bndB, cc, procpos, scan, crank, krank, fresh, candf, warp_tmp, sh, scan64, warp_tmp64, cell_off, keys_xy, keys_node, keys_r, total;
};
__host__ __device__ inline OctLayout oct_layout(int NC, int key_cap, int n_cell) {
    OctLayout o;
    int p = 0;
    o.cntA = p; p += 4 * NC;
    o.cntB = p; p += 4 * NC;
    o.bndA = p; p += 8 * NC;
    o.bndB = p; p += 8 * NC;
    o.cc = p; p += 16 * NC;
    o.procpos = p; p += 4 * NC;
    o.scan = p; p += 4 * NC;
    o.crank = p; p += 4 * NC;
    o.krank = p; p += 4 * NC;
    o.fresh = p; p += 2 * ((NC + 3) & ~3);  // fresh A and B
    o.candf = p; p += (NC + 3) & ~3;
    o.warp_tmp = p; p += 4 * 36;
    o.sh = p; p += 4 * 8;
    p = (p + 15) & ~15;
    o.scan64 = p; p += 8 * NC;
    o.warp_tmp64 = p; p += 8 * 36;
    o.cell_off = p; p += 4 * (n_cell + 1);
    p = (p + 15) & ~15;
    o.keys_xy = p; p += 4 * key_cap;
    o.keys_node = p; p += 2 * ((key_cap + 1) & ~1);
    o.keys_r = p; p += (key_cap + 3) & ~3;
    o.total = (p + 15) & ~15;
    return o;
}

int octtree_smem_bytes(const OrbGeom& g, int level, int key_smem_cap) {
    return oct_layout(g.lv[level].node_cap, key_smem_cap, g.lv[level].n_cols * g.lv[level].n_rows).total;
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

__global__ void __launch_bounds__(kOctThreads) k_octtree(OrbGeom g, OrbBuffers b, int key_smem_cap, int level_begin) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int l = blockIdx.x + level_begin;
    const LevelGeom L = g.lv[l];
    const int NC = L.node_cap, N = L.quota;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n_cell = L.n_cols * L.n_rows;
    const OctLayout lay = oct_layout(NC, key_smem_cap, n_cell);
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

    // ---- gather the level's candidates in the order vToDistributeKeys is built: cell row, cell column, raster.
    //      Cell offsets are scanned in shared memory; every key then finds its cell by binary search, so all global
    //      loads of the gather are independent (a per-cell loop would serialise ~30 dependent L2 round trips per warp).
    int* cell_off = reinterpret_cast<int*>(smem + lay.cell_off);
    for (int i = tid; i < n_cell; i += nt) cell_off[i] = b.cell_count[L.cell_base + i];
    __syncthreads();
    const int M = block_excl_scan(cell_off, n_cell, warp_tmp);
    if (tid == 0) cell_off[n_cell] = M;
    uint32_t* kxy;
    uint16_t* knode;
    uint8_t* kr;
    if (M <= key_smem_cap) {
        kxy = reinterpret_cast<uint32_t*>(smem + lay.keys_xy);
        knode = reinterpret_cast<uint16_t*>(smem + lay.keys_node);
        kr = smem + lay.keys_r;
    } else {
        kxy = b.key_xy + L.cand_base;
        knode = b.key_node + L.cand_base;
        kr = b.key_r + L.cand_base;
    }
    __syncthreads();
    for (int k = tid; k < M; k += nt) {
        int lo = 0, hi = n_cell;  // last cell with cell_off[c] <= k
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cell_off[mid] <= k) lo = mid; else hi = mid;
        }
        const int src = L.cand_base + lo * L.slot + (k - cell_off[lo]);
        kxy[k] = b.cand_xy[src];
        kr[k] = b.cand_r[src];
    }
    if (tid == 0) b.level_cand[l] = M;
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
    int s = block_excl_scan(scan, L.n_ini, warp_tmp);
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
        // ---- near-quota phase (:660-733): generic step with a sorted processing order and an early stop
        int m;
        if (!final_phase) {
            for (int i = tid; i < s; i += nt) {
                const int f = cnt_cur[i] > 1;
                scan[i] = f;
                candf[i] = (uint8_t)f;
            }
            __syncthreads();
            m = block_excl_scan(scan, s, warp_tmp);
            for (int i = tid; i < s; i += nt)
                if (candf[i]) procpos[scan[i]] = i;
        } else {
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
        }
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
    // ---- keep the max-response key of every node, first in candidate order on ties (:738-757)
    uint32_t* best = reinterpret_cast<uint32_t*>(cc);
    for (int i = tid; i < s; i += nt) best[i] = 0;
    __syncthreads();
    for (int k = tid; k < M; k += nt) atomicMax(&best[knode[k]], (uint32_t)kr[k] << 24 | (0xffffffu - (uint32_t)k));
    __syncthreads();
    for (int i = tid; i < s; i += nt) {
        const int k = (int)(0xffffffu - (best[i] & 0xffffffu));
        const uint32_t xy = kxy[k];
        b.lvl_kp[L.kp_base + i] = make_uint2(((xy & 0xffff) + kBorder) | ((xy >> 16) + kBorder) << 16, kr[k]);
    }
    if (tid == 0) b.level_count[l] = s;
}

cudaError_t prepare_octtree(const OrbGeom& g, int key_smem_cap, int* smem_bytes_out) {
    int bytes = 0;
    for (int l = 0; l < g.n_levels; l++) bytes = max(bytes, octtree_smem_bytes(g, l, key_smem_cap));
    *smem_bytes_out = bytes;
    return cudaFuncSetAttribute(k_octtree, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

void launch_octtree(const OrbGeom& g, const OrbBuffers& b, int level, int key_smem_cap, int smem_bytes, cudaStream_t s) {
    if (level < 0) k_octtree<<<g.n_levels, kOctThreads, smem_bytes, s>>>(g, b, key_smem_cap, 0);
    else k_octtree<<<1, kOctThreads, octtree_smem_bytes(g, level, key_smem_cap), s>>>(g, b, key_smem_cap, level);
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
__global__ void __launch_bounds__(256) k_orient_desc(OrbGeom g, OrbBuffers b) {
    // pattern transposed to [point-in-byte 0..15][lane 0..31] so the 32 lanes of a warp read 64 contiguous bytes
    __shared__ char2 pat[16 * 32];
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

void launch_orient_desc(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s) {
    k_orient_desc<<<(g.kp_cap + 7) / 8, 256, 0, s>>>(g, b);
}

}  // namespace corb
