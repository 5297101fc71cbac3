// ORBmatcher::SearchByProjection on sm_100a:
//   K16 k_proj_candidates  one warp per query (last-frame feature / local map point): projection, Frame::GetFeaturesInArea
//                          over the 64 x 48 grid in the reference's enumeration order, level / window / right-coordinate
//                          filters, 256-bit Hamming distances, candidates sorted by (distance, enumeration order)
//   K17 k_proj_resolve     the reference's greedy loop (a feature that received a MapPoint with observations is skipped
//                          by every later query) is order dependent: one thread walks the queries in order, but all it
//                          does per query is to take the first not-yet-taken entry of the sorted list (and the next one
//                          for the ratio test), so the serial part is a few shared-memory reads per query.
//                          Rotation histogram + ComputeThreeMaxima run at the end.
// Reference: corbslam_client/src/ORBmatcher.cc:44-139, 1470-1614, 1746-1787; Frame.cc:331-384. Float arithmetic is
// evaluated with explicit round-to-nearest operations in the order the reference (built without FMA contraction)
// evaluates it; cv::Mat products follow cv::gemm as pinned by the oracle (see gemm3).
#include "common.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>

struct corb_matcher;

namespace corb {

int matcher_device(const corb_matcher* m);
cudaStream_t matcher_stream(const corb_matcher* m);
int matcher_proj_reserve(corb_matcher* m, size_t bytes, uint8_t** d, uint8_t** h);

constexpr int kGridCols = 64, kGridRows = 48;  // Frame.h:38-39
constexpr int kThHigh = 100, kHistoLength = 30;  // ORBmatcher.cc:37-39

struct ProjFrame {  // device view of corb_frame_view
    int n, n_levels;
    const float *x, *y, *angle, *u_right, *scale;
    const int *octave, *grid_off, *grid_idx;
    const uint4* desc;
    const uint8_t* taken;
    float min_x, min_y, max_x, max_y, gw_inv, gh_inv, fx, fy, cx, cy, mbf, mb;
    float Tcw[12];
};

struct ProjQueries {
    int n, variant;  // 0 = last frame (ORBmatcher.cc:1470), 1 = local map points (:44)
    const uint8_t *valid, *blocks;
    const float* xyz;       // variant 0: world position; variant 1: (mTrackProjX, mTrackProjY, mTrackProjXR)
    const uint4* desc;
    const int* level;       // last octave / mnTrackScaleLevel
    const float* aux;       // last angle / mTrackViewCos
    float th, nnratio;
    int forward, backward, check_ori;
};

struct ProjWork {
    uint2* cand;     // [n_query][K]: .x = dist << 20 | enumeration position, .y = feature index | octave << 24
    int* ncand;      // [n_query]
    int* overflow;   // max candidates seen when a list did not fit
    int* match;      // [frame.n]
    int* nmatches;
    int* ev_idx;     // [n_query] rotation-histogram events
    int* ev_bin;
    int K;
};

__device__ __forceinline__ int ham256q(const uint4& a0, const uint4& a1, const uint4* b) {
    const uint4 b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// cv::gemm on CV_32F, 3x3 (optionally transposed) times 3x1 (+ beta * c), as pinned against cv2.gemm by the oracle's golden
// vectors: untransposed = OpenCV's small-matrix case (float dot product, left to right, then alpha / beta in double);
// transposed = the generic kernel (double accumulation; products of two floats are exact in double). One rounding at the end.
__device__ __host__ inline void gemm3(const float* M, bool transpose, double alpha, const float* v, double beta, const float* c, float* out) {
    for (int r = 0; r < 3; r++) {
        double s;
        if (!transpose) {
#ifdef __CUDA_ARCH__
            const float t = __fadd_rn(__fadd_rn(__fmul_rn(M[r * 4], v[0]), __fmul_rn(M[r * 4 + 1], v[1])), __fmul_rn(M[r * 4 + 2], v[2]));
#else
            volatile float p0 = M[r * 4] * v[0], p1 = M[r * 4 + 1] * v[1], p2 = M[r * 4 + 2] * v[2];  // no FMA contraction
            volatile float p01 = p0 + p1;
            const float t = p01 + p2;
#endif
            s = (double)t * alpha;
        } else {
            s = 0;
            for (int k = 0; k < 3; k++) s = s + (double)M[k * 4 + r] * (double)v[k];
            s = s * alpha;
        }
        if (c) s = s + beta * (double)c[r];
        out[r] = (float)s;
    }
}

__global__ void __launch_bounds__(256) k_proj_candidates(ProjFrame F, ProjQueries Q, ProjWork W) {
    extern __shared__ uint2 s_list[];  // [8 warps][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * 8 + warp;
    if (q >= Q.n) return;
    uint2* mine = s_list + (size_t)warp * W.K;
    int nc = 0;
    bool active = Q.valid[q] != 0;
    float u = 0, v = 0, win = 0, ur_ref = 0;
    int minLevel = 0, maxLevel = -1;
    if (active) {
        const int lvl = Q.level[q];
        if (Q.variant == 0) {
            const float tcw[3] = {F.Tcw[3], F.Tcw[7], F.Tcw[11]};
            const float xw[3] = {Q.xyz[3 * q], Q.xyz[3 * q + 1], Q.xyz[3 * q + 2]};
            float xc[3];
            gemm3(F.Tcw, false, 1.0, xw, 1.0, tcw, xc);                       // Rcw*x3Dw+tcw (:1504)
            const float invzc = (float)(1.0 / (double)xc[2]);                 // :1508
            if (invzc < 0) active = false;
            u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, xc[0]), invzc), F.cx);    // :1513-1514
            v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, xc[1]), invzc), F.cy);
            if (u < F.min_x || u > F.max_x || v < F.min_y || v > F.max_y) active = false;
            win = __fmul_rn(Q.th, F.scale[lvl]);                              // :1524
            ur_ref = __fsub_rn(u, __fmul_rn(F.mbf, invzc));                   // :1561
            if (Q.forward) { minLevel = lvl; maxLevel = -1; }                 // :1528-1533
            else if (Q.backward) { minLevel = 0; maxLevel = lvl; }
            else { minLevel = lvl - 1; maxLevel = lvl + 1; }
        } else {
            float r = Q.aux[q] > 0.998f ? 2.5f : 4.0f;                        // RadiusByViewingCos (:133-139)
            if (Q.th != 1.0f) r = __fmul_rn(r, Q.th);
            win = __fmul_rn(r, F.scale[lvl]);
            u = Q.xyz[3 * q];
            v = Q.xyz[3 * q + 1];
            ur_ref = Q.xyz[3 * q + 2];
            minLevel = lvl - 1;
            maxLevel = lvl;
        }
    }
    if (active) {  // Frame::GetFeaturesInArea (Frame.cc:331-384)
        const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, F.min_x), win), F.gw_inv)));
        const int nMaxCellX = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, F.min_x), win), F.gw_inv)));
        const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, F.min_y), win), F.gh_inv)));
        const int nMaxCellY = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, F.min_y), win), F.gh_inv)));
        if (nMinCellX >= kGridCols || nMaxCellX < 0 || nMinCellY >= kGridRows || nMaxCellY < 0) active = false;
        if (active) {
            const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
            const uint4 d0 = Q.desc[2 * (size_t)q], d1 = Q.desc[2 * (size_t)q + 1];
            for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
                // the cells (ix, nMinCellY..nMaxCellY) are consecutive in the CSR: one run of entries
                const int e0 = F.grid_off[ix * kGridRows + nMinCellY], e1 = F.grid_off[ix * kGridRows + nMaxCellY + 1];
                for (int eb = e0; eb < e1; eb += 32) {
                    const int e = eb + lane;
                    bool ok = false;
                    int i = 0, oct = 0, dist = 0;
                    if (e < e1) {
                        i = F.grid_idx[e];
                        oct = F.octave[i];
                        ok = true;
                        if (bCheckLevels) {
                            if (oct < minLevel) ok = false;
                            if (maxLevel >= 0 && oct > maxLevel) ok = false;
                        }
                        const float distx = __fsub_rn(F.x[i], u), disty = __fsub_rn(F.y[i], v);
                        if (!(fabsf(distx) < win && fabsf(disty) < win)) ok = false;
                        const float ur = F.u_right[i];
                        if (ok && ur > 0) {
                            const float er = fabsf(__fsub_rn(ur_ref, ur));  // :87-92, :1559-1565 (the `taken` test comes later)
                            if (er > win) ok = false;
                        }
                        if (ok) dist = ham256q(d0, d1, F.desc + 2 * (size_t)i);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int pos = nc + __popc(bal & ((1u << lane) - 1u));
                        if (pos < W.K) mine[pos] = make_uint2((uint32_t)dist << 20 | (uint32_t)pos, (uint32_t)i | (uint32_t)oct << 24);
                    }
                    nc += __popc(bal);
                }
            }
        }
    }
    __syncwarp();
    if (nc > W.K) {
        if (lane == 0) atomicMax(W.overflow, nc);
        nc = W.K;
    }
    // sort by (distance, enumeration position): rank by counting, keys are unique
    uint2* out = W.cand + (size_t)q * W.K;
    for (int a = lane; a < nc; a += 32) {
        const uint2 me = mine[a];
        int rank = 0;
        for (int b = 0; b < nc; b++) rank += mine[b].x < me.x;
        out[rank] = me;
    }
    if (lane == 0) W.ncand[q] = nc;
}

// ComputeThreeMaxima (ORBmatcher.cc:1746-1787) on bin counts
__device__ inline void three_maxima_counts(const int* cnt, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
}

constexpr int kResolveChunk = 1024, kHeads = 4;

__global__ void __launch_bounds__(256) k_proj_resolve(ProjFrame F, ProjQueries Q, ProjWork W, int stage_angle) {
    // Everything the serial walk touches sits in shared memory: a global load on its path would cost one L2 round trip
    // (~700 cycles) per query, ten times the walk itself.
    extern __shared__ uint8_t s_dyn[];  // [F.n] taken flags, then (stage_angle) [F.n] frame angles
    uint8_t* s_taken = s_dyn;
    const float* s_angle = stage_angle ? reinterpret_cast<const float*>(s_dyn + ((F.n + 15) & ~15)) : F.angle;
    __shared__ uint2 heads[kResolveChunk][kHeads];
    __shared__ int hn[kResolveChunk];
    __shared__ float s_aux[kResolveChunk];
    __shared__ uint8_t s_blocks[kResolveChunk];
    __shared__ int bins[kHistoLength];
    __shared__ int s_nm, s_nev, s_keep[3];
    const int tid = threadIdx.x;
    for (int i = tid; i < F.n; i += 256) {
        s_taken[i] = F.taken ? F.taken[i] : 0;
        W.match[i] = -1;
        if (stage_angle) reinterpret_cast<float*>(s_dyn + ((F.n + 15) & ~15))[i] = F.angle[i];
    }
    if (tid < kHistoLength) bins[tid] = 0;
    if (tid == 0) { s_nm = 0; s_nev = 0; }
    const float factor = 1.0f / kHistoLength;
    for (int c0 = 0; c0 < Q.n; c0 += kResolveChunk) {
        const int cn = min(kResolveChunk, Q.n - c0);
        __syncthreads();
        for (int i = tid; i < cn * kHeads; i += 256) {
            const int ql = i / kHeads, h = i - ql * kHeads;
            const int nc = W.ncand[c0 + ql];
            if (h == 0) {
                hn[ql] = nc;
                s_aux[ql] = Q.aux[c0 + ql];
                s_blocks[ql] = Q.blocks ? Q.blocks[c0 + ql] : 1;
            }
            if (h < nc) heads[ql][h] = W.cand[(size_t)(c0 + ql) * W.K + h];
        }
        __syncthreads();
        if (tid < 32) {
            // One warp walks the chunk 32 queries at a time. Every lane proposes the first (and, for the ratio test, the
            // second) entry of its sorted list that is not taken yet; a lane is final when no earlier lane of the batch
            // takes one of the features it looked at, so the longest conflict-free prefix commits at once and the rest
            // proposes again. The lowest pending lane never conflicts, and conflicts are rare (a few per batch), so this
            // is the reference's query order at ~1/20 of the cost of a one-thread walk.
            const int lane = tid;
            int nm = s_nm, nev = s_nev;
            for (int base = 0; base < cn; base += 32) {
                const int ql = base + lane, q = c0 + ql;
                const int nc = ql < cn ? hn[ql] : 0;
                const uint2* full = W.cand + (size_t)q * W.K;
                unsigned pending = __ballot_sync(0xffffffffu, nc > 0);
                int hpos = 0;
                while (pending) {
                    const bool mine = pending >> lane & 1u;
                    int found = 0, f1 = -1, f2 = -1;
                    uint2 e1 = make_uint2(0, 0), e2 = make_uint2(0, 0);
                    if (mine) {
                        for (int h = hpos; h < nc && found < (Q.variant == 1 ? 2 : 1); h++) {
                            const uint2 e = h < kHeads ? heads[ql][h] : full[h];
                            if (s_taken[e.y & 0xffffff]) continue;
                            if (found == 0) { e1 = e; hpos = h; } else e2 = e;
                            found++;
                        }
                        if (found > 0) f1 = (int)(e1.y & 0xffffff);
                        if (found > 1) f2 = (int)(e2.y & 0xffffff);
                    }
                    bool acc = false;
                    if (found > 0) {
                        const int bestDist = (int)(e1.x >> 20);
                        acc = bestDist <= kThHigh;
                        if (acc && Q.variant == 1) {
                            const int bestLevel = (int)(e1.y >> 24);
                            const int bestDist2 = found > 1 ? (int)(e2.x >> 20) : 256, bestLevel2 = found > 1 ? (int)(e2.y >> 24) : -1;
                            if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(Q.nnratio, (float)bestDist2)) acc = false;  // :117-120
                        }
                    }
                    const unsigned accmask = __ballot_sync(0xffffffffu, mine && acc);
                    bool conflicted = false;
                    for (unsigned mm = accmask; mm; mm &= mm - 1) {  // every accepting lane against the later lanes' first / second entries
                        const int j = __ffs(mm) - 1;
                        const int pj = __shfl_sync(0xffffffffu, f1, j);
                        if (j < lane && (pj == f1 || pj == f2)) conflicted = true;
                    }
                    const unsigned confmask = __ballot_sync(0xffffffffu, mine && conflicted);
                    const int first_conf = confmask ? __ffs(confmask) - 1 : 32;
                    const bool commit = mine && lane < first_conf;
                    bool ev = false;
                    int bin = 0;
                    if (commit && acc) {
                        W.match[f1] = q;
                        if (s_blocks[ql]) s_taken[f1] = 1;
                        if (Q.variant == 0 && Q.check_ori) {
                            float rot = __fsub_rn(s_aux[ql], s_angle[f1]);  // :1585-1593
                            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                            bin = (int)roundf(__fmul_rn(rot, factor));
                            if (bin == kHistoLength) bin = 0;
                            ev = bin >= 0 && bin < kHistoLength;
                        }
                    }
                    const unsigned cm = __ballot_sync(0xffffffffu, commit && acc), em = __ballot_sync(0xffffffffu, ev);
                    if (ev) {
                        const int pos = nev + __popc(em & ((1u << lane) - 1u));
                        atomicAdd(&bins[bin], 1);
                        W.ev_idx[pos] = f1;
                        W.ev_bin[pos] = bin;
                    }
                    nm += __popc(cm);
                    nev += __popc(em);
                    pending &= first_conf >= 32 ? 0u : ~((1u << first_conf) - 1u);
                    __syncwarp();
                }
            }
            if (lane == 0) { s_nm = nm; s_nev = nev; }
        }
    }
    __syncthreads();
    if (Q.variant == 0 && Q.check_ori) {
        if (tid == 0) {
            int i1, i2, i3;
            three_maxima_counts(bins, kHistoLength, i1, i2, i3);
            s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3;
            int removed = 0;
            for (int b = 0; b < kHistoLength; b++)
                if (b != i1 && b != i2 && b != i3) removed += bins[b];
            s_nm -= removed;
        }
        __syncthreads();
        __threadfence_block();
        for (int e = tid; e < s_nev; e += 256) {
            const int b = W.ev_bin[e];
            if (b != s_keep[0] && b != s_keep[1] && b != s_keep[2]) W.match[W.ev_idx[e]] = -1;
        }
    }
    __syncthreads();
    if (tid == 0) *W.nmatches = s_nm;
}

static int run_projection(corb_matcher* m, const corb_frame_view* f, int variant, int nq, const uint8_t* valid, const uint8_t* blocks,
                          const float* xyz, const uint8_t* qdesc, const int32_t* level, const float* aux, const float* Tlw, float th,
                          float nnratio, int mono, int check_ori, int32_t* match, int32_t* nmatches) {
    CORB_CHECK(m && f && match && nmatches && nq >= 0, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(f->n >= 0 && f->n < (1 << 24) && f->n_levels >= 1 && f->n_levels <= 64, CORB_ERR_INVALID, "bad frame view");
    CORB_CHECK(f->n == 0 || (f->x && f->y && f->octave && f->angle && f->desc && f->u_right && f->grid_off && f->grid_idx && f->scale_factors),
               CORB_ERR_INVALID, "frame view has NULL arrays");
    CORB_CHECK(nq == 0 || (valid && xyz && qdesc && level && aux), CORB_ERR_INVALID, "query arrays are NULL");
    *nmatches = 0;
    for (int i = 0; i < f->n; i++) match[i] = -1;
    if (nq == 0 || f->n == 0) {
        int ndev = 0;
        CORB_CUDA(cudaGetDeviceCount(&ndev));  // still no CPU path: fail without a device
        return CORB_OK;
    }
    for (int q = 0; q < nq; q++)
        CORB_CHECK(!valid[q] || (level[q] >= 0 && level[q] < f->n_levels), CORB_ERR_INVALID, "query %d: level %d out of range", q, level[q]);
    CORB_CUDA(cudaSetDevice(matcher_device(m)));
    cudaStream_t st = matcher_stream(m);
    const int n = f->n, ncell = kGridCols * kGridRows;
    const int n_in_grid = f->grid_off[ncell];
    CORB_CHECK(n_in_grid >= 0 && n_in_grid <= n, CORB_ERR_INVALID, "grid CSR inconsistent");
    int forward = 0, backward = 0;
    if (variant == 0) {
        CORB_CHECK(Tlw, CORB_ERR_INVALID, "Tlw is NULL");
        float twc[3], tlc[3];
        const float tcw[3] = {f->Tcw[3], f->Tcw[7], f->Tcw[11]}, tlw[3] = {Tlw[3], Tlw[7], Tlw[11]};
        // twc = -Rcw.t()*tcw (:1483): cv::MatExpr's unary minus materialises the transpose first (matop.cpp: MatOp::subtract(Scalar,
        // expr) assigns the T node to a Mat), so the product is an UNtransposed gemm on the stored transpose with alpha = -1,
        // i.e. the small-matrix float path (checked against the reference's own expression in tests/test_ref_cpu.py).
        const float Rt[12] = {f->Tcw[0], f->Tcw[4], f->Tcw[8], 0.f, f->Tcw[1], f->Tcw[5], f->Tcw[9], 0.f, f->Tcw[2], f->Tcw[6], f->Tcw[10], 0.f};
        gemm3(Rt, false, -1.0, tcw, 0.0, nullptr, twc);
        gemm3(Tlw, false, 1.0, twc, 1.0, tlw, tlc);          // tlc = Rlw*twc+tlw (:1488)
        forward = tlc[2] > f->mb && !mono;
        backward = -tlc[2] > f->mb && !mono;
    }
    for (int K = 128;; K = 1024) {
#ifdef CORB_PROJ_TRACE
        const auto tp0 = std::chrono::steady_clock::now();
#endif
        // ---- pack everything into one pinned block -> one H2D copy
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = align_up_sz(off + bytes, 16); return o; };
        const size_t oX = take(4 * (size_t)n), oY = take(4 * (size_t)n), oOct = take(4 * (size_t)n), oAng = take(4 * (size_t)n),
                     oDesc = take(32 * (size_t)n), oUr = take(4 * (size_t)n), oTaken = take(n), oGoff = take(4 * (size_t)(ncell + 1)),
                     oGidx = take(4 * (size_t)std::max(1, n_in_grid)), oScale = take(4 * (size_t)f->n_levels);
        const size_t oQv = take(nq), oQb = take(nq), oQx = take(12 * (size_t)nq), oQd = take(32 * (size_t)nq), oQl = take(4 * (size_t)nq),
                     oQa = take(4 * (size_t)nq);
        const size_t in_bytes = off;
        const size_t oCand = take(8 * (size_t)nq * K), oNc = take(4 * (size_t)nq), oEvI = take(4 * (size_t)nq), oEvB = take(4 * (size_t)nq);
        const size_t oOut = take(4 * (size_t)n + 16);
        const size_t oMisc = oOut + 4 * (size_t)n;  // [nmatches, overflow]
        uint8_t *d, *h;
        int rc = matcher_proj_reserve(m, off, &d, &h);
        if (rc != CORB_OK) return rc;
        memcpy(h + oX, f->x, 4 * (size_t)n); memcpy(h + oY, f->y, 4 * (size_t)n); memcpy(h + oOct, f->octave, 4 * (size_t)n);
        memcpy(h + oAng, f->angle, 4 * (size_t)n); memcpy(h + oDesc, f->desc, 32 * (size_t)n); memcpy(h + oUr, f->u_right, 4 * (size_t)n);
        if (f->taken) memcpy(h + oTaken, f->taken, n); else memset(h + oTaken, 0, n);
        memcpy(h + oGoff, f->grid_off, 4 * (size_t)(ncell + 1));
        if (n_in_grid) memcpy(h + oGidx, f->grid_idx, 4 * (size_t)n_in_grid);
        memcpy(h + oScale, f->scale_factors, 4 * (size_t)f->n_levels);
        memcpy(h + oQv, valid, nq);
        if (blocks) memcpy(h + oQb, blocks, nq); else memset(h + oQb, 1, nq);
        memcpy(h + oQx, xyz, 12 * (size_t)nq); memcpy(h + oQd, qdesc, 32 * (size_t)nq); memcpy(h + oQl, level, 4 * (size_t)nq);
        memcpy(h + oQa, aux, 4 * (size_t)nq);
#ifdef CORB_PROJ_TRACE
        static cudaEvent_t e0, e1, e2, e3;
        static bool evi = false;
        if (!evi) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3); evi = true; }
        const auto tp1 = std::chrono::steady_clock::now();
        cudaEventRecord(e0, st);
#endif
        CORB_CUDA(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st));
        CORB_CUDA(cudaMemsetAsync(d + oMisc, 0, 16, st));
        ProjFrame F;
        F.n = n; F.n_levels = f->n_levels;
        F.x = (const float*)(d + oX); F.y = (const float*)(d + oY); F.octave = (const int*)(d + oOct); F.angle = (const float*)(d + oAng);
        F.desc = (const uint4*)(d + oDesc); F.u_right = (const float*)(d + oUr); F.taken = d + oTaken;
        F.grid_off = (const int*)(d + oGoff); F.grid_idx = (const int*)(d + oGidx); F.scale = (const float*)(d + oScale);
        F.min_x = f->min_x; F.min_y = f->min_y; F.max_x = f->max_x; F.max_y = f->max_y; F.gw_inv = f->grid_w_inv; F.gh_inv = f->grid_h_inv;
        F.fx = f->fx; F.fy = f->fy; F.cx = f->cx; F.cy = f->cy; F.mbf = f->mbf; F.mb = f->mb;
        memcpy(F.Tcw, f->Tcw, sizeof(F.Tcw));
        ProjQueries Q;
        Q.n = nq; Q.variant = variant;
        Q.valid = d + oQv; Q.blocks = d + oQb; Q.xyz = (const float*)(d + oQx); Q.desc = (const uint4*)(d + oQd);
        Q.level = (const int*)(d + oQl); Q.aux = (const float*)(d + oQa);
        Q.th = th; Q.nnratio = nnratio; Q.forward = forward; Q.backward = backward; Q.check_ori = check_ori;
        ProjWork W;
        W.cand = (uint2*)(d + oCand); W.ncand = (int*)(d + oNc); W.ev_idx = (int*)(d + oEvI); W.ev_bin = (int*)(d + oEvB);
        W.match = (int*)(d + oOut); W.nmatches = (int*)(d + oMisc); W.overflow = W.nmatches + 1; W.K = K;
        const size_t sm1 = (size_t)8 * K * sizeof(uint2);
        if (sm1 > 48 * 1024) CORB_SMEM_OPT_IN(k_proj_candidates);
#ifdef CORB_PROJ_TRACE
        cudaEventRecord(e1, st);
#endif
        k_proj_candidates<<<(nq + 7) / 8, 256, sm1, st>>>(F, Q, W);
#ifdef CORB_PROJ_TRACE
        cudaEventRecord(e2, st);
#endif
        CORB_CHECK(n <= 128 * 1024, CORB_ERR_UNSUPPORTED, "frame with %d features", n);
        const int stage_angle = n <= 24 * 1024;
        const size_t sm2 = (size_t)((n + 15) & ~15) + (stage_angle ? 4 * (size_t)n : 0) + 16;
        CORB_SMEM_OPT_IN(k_proj_resolve);
        k_proj_resolve<<<1, 256, sm2, st>>>(F, Q, W, stage_angle);
        CORB_CUDA(cudaGetLastError());
#ifdef CORB_PROJ_TRACE
        cudaEventRecord(e3, st);
#endif
        CORB_CUDA(cudaMemcpyAsync(h + oOut, d + oOut, 4 * (size_t)n + 16, cudaMemcpyDeviceToHost, st));
        CORB_CUDA(cudaStreamSynchronize(st));
#ifdef CORB_PROJ_TRACE
        {
            float a = 0, b2 = 0, c2 = 0;
            cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b2, e1, e2); cudaEventElapsedTime(&c2, e2, e3);
            const auto tp2 = std::chrono::steady_clock::now();
            static int cnt = 0;
            if (++cnt % 50 == 0)
                printf("proj trace: pack %.1f us | h2d %.1f us candidates %.1f us resolve %.1f us | launch..sync %.1f us\n",
                       std::chrono::duration<double, std::micro>(tp1 - tp0).count(), a * 1e3, b2 * 1e3, c2 * 1e3,
                       std::chrono::duration<double, std::micro>(tp2 - tp1).count());
        }
#endif
        const int* misc = (const int*)(h + oMisc);
        if (misc[1] > K) {  // a candidate list did not fit: once more with room for 1024 per query
            CORB_CHECK(K < 1024 && misc[1] <= 1024, CORB_ERR_CAPACITY, "%d candidates in one search window", misc[1]);
            continue;
        }
        memcpy(match, h + oOut, 4 * (size_t)n);
        *nmatches = misc[0];
        return CORB_OK;
    }
}

}  // namespace corb

using namespace corb;

extern "C" {

int corb_search_by_projection_last(corb_matcher* m, const corb_frame_view* cur, int32_t n_last, const uint8_t* last_valid,
                                   const uint8_t* last_blocks, const float* last_xyz, const uint8_t* last_mp_desc,
                                   const int32_t* last_octave, const float* last_angle, const float* Tlw, float th, int mono,
                                   int check_orientation, int32_t* match, int32_t* nmatches) {
    return run_projection(m, cur, 0, n_last, last_valid, last_blocks, last_xyz, last_mp_desc, last_octave, last_angle, Tlw, th, 0.f, mono,
                          check_orientation, match, nmatches);
}

int corb_search_by_projection_map(corb_matcher* m, const corb_frame_view* F, int32_t n_mp, const uint8_t* in_view, const uint8_t* blocks,
                                  const float* proj, const int32_t* level, const float* view_cos, const uint8_t* mp_desc, float th,
                                  float nnratio, int32_t* match, int32_t* nmatches) {
    return run_projection(m, F, 1, n_mp, in_view, blocks, proj, mp_desc, level, view_cos, nullptr, th, nnratio, 0, 0, match, nmatches);
}

}  // extern "C"
