// EPnP-RANSAC for a batch of relocalisation / map-fusion candidates:
//   corbslam_client/src/PnPsolver.cc:206-300  iterate       -> k_pnp_solve<0> + k_pnp_check + k_pnp_records + k_pnp_finalize
//   corbslam_client/src/PnPsolver.cc:302-346  Refine        -> k_pnp_solve<1> + k_pnp_check
//   corbslam_client/src/PnPsolver.cc:349-383  CheckInliers  -> k_pnp_check
//   corbslam_client/src/PnPsolver.cc:420-962  EPnP          -> pnp_core.cuh (one thread per hypothesis)
//
// The reference loop is sequential only in its bookkeeping: hypothesis `it` depends on nothing but its four draws, and
// Refine() works on mvbBestInliers, i.e. it is a function of WHICH hypothesis is the best so far. So every hypothesis is
// computed at once, the best-so-far records (strict running maxima of the inlier count) are found by one scan, Refine()
// is evaluated once per record, and the iteration at which the reference would return is the first one (from the
// caller's mnIterations on) with enough inliers whose current record refines successfully.
//
// Compiled with -fmad=false (csrc/Makefile): double arithmetic is then the same sequence of correctly rounded
// operations as the oracle's, and the parity tests compare poses bit for bit.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"
#include "pnp_core.cuh"

struct corb_matcher;
namespace corb {
int matcher_device(const corb_matcher* m);
cudaStream_t matcher_stream(const corb_matcher* m);
int matcher_pnp_reserve(corb_matcher* m, size_t bytes, uint8_t** d, uint8_t** h);
}  // namespace corb

using namespace corb;
using namespace corb::pnp;

namespace {

struct PnpProb {       // one solver object on the device
    int n, n_words;    // correspondences, mask words per hypothesis
    int pt_off;        // first correspondence in the packed point arrays
    int it_begin, it_end;
    int min_inliers;
    int hyp_off;       // first hypothesis slot (it_end slots per problem)
    int mask_off;      // first mask word (it_end * n_words per problem)
    int out_off;       // first inlier byte of the result
    float fx, fy, cx, cy;
};

struct PnpDev {
    const PnpProb* probs;
    const float* p3d;
    const float* p2d;
    const float* max_err;
    const int4* draws;   // indexed by hyp_off + it
    double* hyp_Rt;      // [H][12]
    int* hyp_cnt;        // [H]
    uint32_t* hyp_mask;  // per problem it_end x n_words
    int* rec_iter;       // [H] per problem: iteration of record r
    int* n_rec;          // [C]
    int* slot_of_iter;   // [H] record slot that is "best so far" at iteration it, -1 when the iteration has < min inliers
    double* ref_Rt;      // [H][12] by record slot
    int* ref_cnt;
    uint32_t* ref_mask;
    int* res;            // [C][4 + 16]: status, no_more, n_inliers, iterations, Tcw (float bits)
    uint8_t* res_inl;
};

constexpr int kSolveThreads = 32;  // lanes per warp of a solve block; a block has NB warps
constexpr int kTeamWs = WS_DOUBLES + 1;  // workspace stride of a lead-only team (odd: the four leads of a warp hit different banks)

// Shared memory of one warp of a solve block: its workspaces, then the shared Jacobi areas of its teams.
template <int MODE, int TEAM>
__host__ __device__ constexpr size_t warp_doubles() {
    return (MODE == 0 && TEAM > 1 ? (size_t)(kSolveThreads / TEAM) * kTeamWs : (size_t)WS_DOUBLES * kSolveThreads) +
           (TEAM > 1 ? (size_t)(kSolveThreads / TEAM) * COOP_DOUBLES : 0) + (MODE == 1 && TEAM == 32 ? (size_t)TBUF_DOUBLES : 0);
}
template <int MODE, int TEAM, int NB>
constexpr size_t solve_smem() { return (NB * warp_doubles<MODE, TEAM>() + (NB > 1 ? (kSolveThreads / TEAM) * 3 * 13 : 0)) * sizeof(double); }

// MODE 0: RANSAC iterations (minimal set of 4 from the draws, :228-246). MODE 1: best-so-far records (EPnP over the
// record's inlier mask = Refine(), :302-324).
// TEAM = lanes per item. 1: one thread per item (large batches: throughput). 8 / 32: a team per item (small batches:
// latency) that runs the 12 x 12 Jacobi sweeps as a wavefront on a shared copy; in MODE 0 the scalar phases run on the
// team's lane 0 (one workspace per team), in MODE 1 every lane replicates them on its own workspace and the ordered sums
// over the inliers are split across the lanes.
// NB = warps per block. 3: warp w evaluates only beta approximation w + 1 of its items (everything before it is
// repeated by the three warps side by side) and warp 0 applies compute_pose's selection rule to the three results.
template <int MODE, int TEAM, int NB>
__global__ void __launch_bounds__(kSolveThreads * NB) k_pnp_solve(PnpDev D) {
    extern __shared__ double smem[];
    constexpr int kItems = kSolveThreads / TEAM;  // items per block
    const PnpProb P = D.probs[blockIdx.y];
    const int n_items = MODE == 0 ? P.it_end : D.n_rec[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int item = blockIdx.x * kItems + lane / TEAM;
    if (TEAM == 1 ? item >= n_items : (int)blockIdx.x * kItems >= n_items) return;  // block-uniform for teams
    const bool valid = item < n_items;
    if (!valid) item = n_items - 1;  // a team without an item repeats the last one (its lanes take part in the barriers)
    PtSet s;
    s.p3d = D.p3d + 3 * (size_t)P.pt_off;
    s.p2d = D.p2d + 2 * (size_t)P.pt_off;
    s.n_words = P.n_words;
    double* out;
    if (MODE == 0) {
        const int4 dr = D.draws[P.hyp_off + item];
        const int r[4] = {dr.x, dr.y, dr.z, dr.w};
        resolve_draws(P.n, r, s.list);  // the vAvailableIndices bookkeeping of :228-242
        s.n = 4;
        s.mask = nullptr;
        out = D.hyp_Rt + 12 * (size_t)(P.hyp_off + item);
    } else {
        const int h = D.rec_iter[P.hyp_off + item];
        s.n = D.hyp_cnt[P.hyp_off + h];
        s.mask = D.hyp_mask + P.mask_off + (size_t)h * P.n_words;
        out = D.ref_Rt + 12 * (size_t)(P.hyp_off + item);
    }
    Epnp e;
    e.fu = (double)P.fx; e.fv = (double)P.fy; e.uc = (double)P.cx; e.vc = (double)P.cy;
    constexpr size_t kWarpDoubles = warp_doubles<MODE, TEAM>();
    double* wbase = smem + (size_t)warp * kWarpDoubles;
    constexpr bool kLeadOnly = MODE == 0 && TEAM > 1;
    Ws ws = kLeadOnly ? Ws{wbase + (size_t)(lane / TEAM) * kTeamWs, 1} : Ws{wbase + lane, kSolveThreads};
    if (TEAM > 1) {
        double* area = wbase + (kLeadOnly ? (size_t)kItems * kTeamWs : (size_t)WS_DOUBLES * kSolveThreads) + (size_t)(lane / TEAM) * COOP_DOUBLES;
        e.coop = Coop{area, area + 12 * COOP_ROW, (int*)(area + 12 * COOP_ROW + 12), lane % TEAM, TEAM,
                      MODE == 1 && TEAM == 32 ? area + COOP_DOUBLES : nullptr, kLeadOnly};
    }
    double Rt[12];
    const double err = e.template compute_pose_part<(MODE == 1 && TEAM == 32)>(s, ws, NB == 3 ? warp + 1 : 0, Rt);
    const bool lead = lane % TEAM == 0;
    if (NB == 1) {
        if (valid && lead) {
#pragma unroll
            for (int i = 0; i < 12; i++) out[i] = Rt[i];
        }
        return;
    }
    // three warps, one beta approximation each: N = 1; if (e2 < e1) N = 2; if (e3 < e[N]) N = 3  (:559-561)
    double* sel = smem + (size_t)NB * kWarpDoubles + (size_t)(lane / TEAM) * 3 * 13;
    if (lead) {
        sel[warp * 13] = err;
#pragma unroll
        for (int i = 0; i < 12; i++) sel[warp * 13 + 1 + i] = Rt[i];
    }
    __syncthreads();
    if (warp == 0 && lead && valid) {
        int N = 0;
        if (sel[13] < sel[0]) N = 1;
        if (sel[26] < sel[N * 13]) N = 2;
#pragma unroll
        for (int i = 0; i < 12; i++) out[i] = sel[N * 13 + 1 + i];
    }
}

// CheckInliers: one warp per pose (hypothesis or refined record), lane = correspondence, ballot = mask word.
constexpr int kCheckWarps = 8;
template <int MODE>
__global__ void __launch_bounds__(kCheckWarps * 32) k_pnp_check(PnpDev D) {
    const PnpProb P = D.probs[blockIdx.y];
    const int item = blockIdx.x * kCheckWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (MODE == 0 ? item >= P.it_end : item >= D.n_rec[blockIdx.y]) return;
    const double* src = (MODE == 0 ? D.hyp_Rt : D.ref_Rt) + 12 * (size_t)(P.hyp_off + item);
    double Rt[12];
#pragma unroll
    for (int i = 0; i < 12; i++) Rt[i] = src[i];
    uint32_t* mask = (MODE == 0 ? D.hyp_mask : D.ref_mask) + P.mask_off + (size_t)item * P.n_words;
    const float* p3d = D.p3d + 3 * (size_t)P.pt_off;
    const float* p2d = D.p2d + 2 * (size_t)P.pt_off;
    const float* me = D.max_err + P.pt_off;
    const double fu = (double)P.fx, fv = (double)P.fy, uc = (double)P.cx, vc = (double)P.cy;
    int cnt = 0;
    for (int w = 0; w < P.n_words; w++) {
        const int i = w * 32 + lane;
        bool in = false;
        if (i < P.n) in = is_inlier(Rt, fu, fv, uc, vc, p3d + 3 * i, p2d + 2 * i, me[i]);
        const unsigned b = __ballot_sync(0xffffffffu, in);
        if (lane == 0) mask[w] = b;
        cnt += __popc(b);
    }
    if (lane == 0) (MODE == 0 ? D.hyp_cnt : D.ref_cnt)[P.hyp_off + item] = cnt;
}

// The best-so-far bookkeeping of iterate() (:248-260) over all iterations from 0: records = strict running maxima of
// the inlier count among iterations with >= min_inliers. One warp per problem, counts staged in shared memory.
__global__ void __launch_bounds__(32) k_pnp_records(PnpDev D) {
    __shared__ int cnt[1024];
    const PnpProb P = D.probs[blockIdx.x];
    const int lane = threadIdx.x;
    int best = 0, slot = -1;  // mnBestInliers, index of the current record
    for (int base = 0; base < P.it_end; base += 1024) {
        const int m = min(1024, P.it_end - base);
        for (int i = lane; i < m; i += 32) cnt[i] = D.hyp_cnt[P.hyp_off + base + i];
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < m; i++) {
                const int c = cnt[i];
                int s = -1;
                if (c >= P.min_inliers) {
                    if (c > best) {
                        best = c;
                        slot++;
                        D.rec_iter[P.hyp_off + slot] = base + i;
                    }
                    s = slot;
                }
                cnt[i] = s;
            }
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) D.slot_of_iter[P.hyp_off + base + i] = cnt[i];
        __syncwarp();
    }
    if (lane == 0) D.n_rec[blockIdx.x] = slot + 1;
}

// Where the reference loop stops: the first iteration >= it_begin with >= min_inliers whose record refines to more than
// min_inliers (:262-273), else the best record at the end (:279-291). One warp per problem.
__global__ void __launch_bounds__(32) k_pnp_finalize(PnpDev D) {
    const PnpProb P = D.probs[blockIdx.x];
    const int lane = threadIdx.x;
    int found = -1;
    for (int base = P.it_begin; base < P.it_end && found < 0; base += 32) {
        const int i = base + lane;
        bool ok = false;
        if (i < P.it_end) {
            const int s = D.slot_of_iter[P.hyp_off + i];
            ok = s >= 0 && D.ref_cnt[P.hyp_off + s] > P.min_inliers;
        }
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        if (b) found = base + __ffs(b) - 1;
    }
    int* res = D.res + 20 * blockIdx.x;
    const int n_rec = D.n_rec[blockIdx.x];
    int status = 0, n_inl = 0, iterations = P.it_end, no_more = 1;
    const double* Rt = nullptr;
    const uint32_t* mask = nullptr;
    if (found >= 0) {
        const int s = D.slot_of_iter[P.hyp_off + found];
        status = 1; no_more = 0; iterations = found + 1;
        n_inl = D.ref_cnt[P.hyp_off + s];
        Rt = D.ref_Rt + 12 * (size_t)(P.hyp_off + s);
        mask = D.ref_mask + P.mask_off + (size_t)s * P.n_words;
    } else if (n_rec > 0) {  // mnBestInliers >= mRansacMinInliers
        const int h = D.rec_iter[P.hyp_off + n_rec - 1];
        status = 2;
        n_inl = D.hyp_cnt[P.hyp_off + h];
        Rt = D.hyp_Rt + 12 * (size_t)(P.hyp_off + h);
        mask = D.hyp_mask + P.mask_off + (size_t)h * P.n_words;
    }
    if (lane == 0) { res[0] = status; res[1] = no_more; res[2] = n_inl; res[3] = iterations; }
    if (lane < 16) {  // Rcw.convertTo(CV_32F), tcw.convertTo(CV_32F) into eye(4) (:254-260, :335-341)
        float v = (lane == 15) ? 1.f : 0.f;
        const int r = lane >> 2, c = lane & 3;
        if (Rt && r < 3) v = (float)(c < 3 ? Rt[3 * r + c] : Rt[9 + r]);
        res[4 + lane] = __float_as_int(v);
    }
    uint8_t* inl = D.res_inl + P.out_off;
    for (int i = lane; i < P.n; i += 32) inl[i] = mask ? (mask[i >> 5] >> (i & 31)) & 1u : 0;
}

// SetRansacParameters (:163-198), host scalar logic
void ransac_params(int N, double probability, int minInliers, int maxIterations, int minSet, float epsilon, int* out_min, int* out_its) {
    float mRansacEpsilon = epsilon;
    int nMinInliers = (int)(N * mRansacEpsilon);
    if (nMinInliers < minInliers) nMinInliers = minInliers;
    if (nMinInliers < minSet) nMinInliers = minSet;
    const int mRansacMinInliers = nMinInliers;
    int nIterations = 1;
    if (N > 0) {
        if (mRansacEpsilon < (float)mRansacMinInliers / N) mRansacEpsilon = (float)mRansacMinInliers / N;
        if (mRansacMinInliers == N)
            nIterations = 1;
        else {
            const double v = ceil(log(1 - probability) / log(1 - pow((double)mRansacEpsilon, 3)));
            // double -> int outside the int range (or NaN: epsilon > 1 when N < minInliers) is undefined in C++; x86's
            // cvttsd2si returns INT_MIN, which the clamp below turns into 1 - that is what the reference binary does
            nIterations = (v > -2147483649.0 && v < 2147483648.0) ? (int)v : INT32_MIN;
        }
    }
    *out_min = mRansacMinInliers;
    *out_its = std::max(1, std::min(nIterations, maxIterations));
}

}  // namespace

extern "C" {

int corb_pnp_ransac_params(int N, double probability, int min_inliers, int max_iterations, int min_set, float epsilon,
                           int* out_min_inliers, int* out_max_its) {
    CORB_CHECK(N >= 0 && out_min_inliers && out_max_its && max_iterations >= 1, CORB_ERR_INVALID, "bad argument");
    ransac_params(N, probability, min_inliers, max_iterations, min_set, epsilon, out_min_inliers, out_max_its);
    return CORB_OK;
}

int corb_pnp_iterate_batch(corb_matcher* m, int n_problems, const corb_pnp_problem* problems, corb_pnp_result* results,
                           uint8_t* const* inliers) {
    CORB_CHECK(m && n_problems >= 0 && (n_problems == 0 || (problems && results)), CORB_ERR_INVALID, "bad argument");
    int ndev = 0;
    CORB_CUDA(cudaGetDeviceCount(&ndev));  // no CPU path: fail without a device even for an empty batch
    CORB_CHECK(ndev > 0, CORB_ERR_CUDA, "no CUDA device");
    // ---- the problems that reach the loop (N >= mRansacMinInliers, :215-219); the others only set bNoMore
    std::vector<int> live;
    size_t n_pts = 0, n_hyp = 0, n_mask = 0;
    int max_it = 0, max_rec = 0;  // records are strict maxima of counts in [min_inliers, n]: at most n - min_inliers + 1
    for (int c = 0; c < n_problems; c++) {
        const corb_pnp_problem& p = problems[c];
        CORB_CHECK(p.n >= 0 && p.n < (1 << 24) && p.min_inliers >= 4 && p.max_its >= 1 && p.iterations_done >= 0 && p.n_iterations >= 0,
                   CORB_ERR_INVALID, "problem %d: bad sizes (n %d, min_inliers %d, max_its %d)", c, p.n, p.min_inliers, p.max_its);
        corb_pnp_result& r = results[c];
        memset(&r, 0, sizeof(r));
        r.iterations = p.iterations_done;
        if (inliers && inliers[c] && p.n) memset(inliers[c], 0, p.n);
        if (p.n < p.min_inliers) { r.no_more = 1; continue; }
        const int it_end = std::max(p.max_its, p.iterations_done + p.n_iterations);
        // (with nIterations == 0 and the iterations exhausted the loop body never runs, but the epilogue :279-291 still
        // reports the best record, which is a function of the hypotheses - they are computed all the same)
        CORB_CHECK(it_end <= (1 << 20), CORB_ERR_CAPACITY, "problem %d: %d iterations", c, it_end);
        CORB_CHECK(p.p2d && p.p3d && p.max_err && p.draws, CORB_ERR_INVALID, "problem %d: NULL arrays", c);
        for (int it = 0; it < it_end; it++)
            for (int k = 0; k < 4; k++) {
                const int d = p.draws[4 * it + k];
                CORB_CHECK(d >= 0 && d < p.n - k, CORB_ERR_INVALID, "problem %d: draw %d of iteration %d = %d is outside [0, %d)", c, k, it, d, p.n - k);
            }
        live.push_back(c);
        n_pts += p.n;
        n_hyp += it_end;
        n_mask += (size_t)it_end * ((p.n + 31) / 32);
        max_it = std::max(max_it, it_end);
        max_rec = std::max(max_rec, std::min(it_end, p.n - p.min_inliers + 1));
    }
    const int C = (int)live.size();
    if (C == 0) return CORB_OK;
    CORB_CUDA(cudaSetDevice(matcher_device(m)));
    cudaStream_t st = matcher_stream(m);
    // ---- one pinned block in (problems, points, draws), one out (results, inlier bytes)
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up_sz(off + bytes, 16); return o; };
    const size_t oProb = take(sizeof(PnpProb) * C), oP3 = take(12 * n_pts), oP2 = take(8 * n_pts), oMe = take(4 * n_pts), oDraw = take(16 * n_hyp);
    const size_t in_bytes = off;
    const size_t oRes = take(80 * (size_t)C), oInl = take(n_pts);
    const size_t out_end = off;
    const size_t oHypRt = take(96 * n_hyp), oRefRt = take(96 * n_hyp), oHypCnt = take(4 * n_hyp), oRefCnt = take(4 * n_hyp),
                 oRecIter = take(4 * n_hyp), oSlot = take(4 * n_hyp), oNrec = take(4 * (size_t)C), oHypMask = take(4 * n_mask),
                 oRefMask = take(4 * n_mask);
    uint8_t *d, *h;
    int rc = matcher_pnp_reserve(m, off, &d, &h);
    if (rc != CORB_OK) return rc;
    PnpProb* hp = (PnpProb*)(h + oProb);
    size_t pt = 0, hy = 0, mk = 0;
    for (int i = 0; i < C; i++) {
        const corb_pnp_problem& p = problems[live[i]];
        const int it_end = std::max(p.max_its, p.iterations_done + p.n_iterations);
        PnpProb& q = hp[i];
        q.n = p.n; q.n_words = (p.n + 31) / 32; q.pt_off = (int)pt; q.it_begin = p.iterations_done; q.it_end = it_end;
        q.min_inliers = p.min_inliers; q.hyp_off = (int)hy; q.mask_off = (int)mk; q.out_off = (int)pt;
        q.fx = p.fx; q.fy = p.fy; q.cx = p.cx; q.cy = p.cy;
        memcpy(h + oP3 + 12 * pt, p.p3d, 12 * (size_t)p.n);
        memcpy(h + oP2 + 8 * pt, p.p2d, 8 * (size_t)p.n);
        memcpy(h + oMe + 4 * pt, p.max_err, 4 * (size_t)p.n);
        memcpy(h + oDraw + 16 * hy, p.draws, 16 * (size_t)it_end);
        pt += p.n; hy += it_end; mk += (size_t)it_end * q.n_words;
    }
    CORB_CHECK(mk < (1u << 31) && hy < (1u << 27), CORB_ERR_CAPACITY, "batch too large (%zu mask words)", mk);
    CORB_CUDA(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st));
    PnpDev D;
    D.probs = (const PnpProb*)(d + oProb); D.p3d = (const float*)(d + oP3); D.p2d = (const float*)(d + oP2);
    D.max_err = (const float*)(d + oMe); D.draws = (const int4*)(d + oDraw);
    D.hyp_Rt = (double*)(d + oHypRt); D.hyp_cnt = (int*)(d + oHypCnt); D.hyp_mask = (uint32_t*)(d + oHypMask);
    D.rec_iter = (int*)(d + oRecIter); D.n_rec = (int*)(d + oNrec); D.slot_of_iter = (int*)(d + oSlot);
    D.ref_Rt = (double*)(d + oRefRt); D.ref_cnt = (int*)(d + oRefCnt); D.ref_mask = (uint32_t*)(d + oRefMask);
    D.res = (int*)(d + oRes); D.res_inl = d + oInl;
    static std::atomic<bool> attr_set[64];  // per device; setting the attribute twice from two threads is harmless
    const int dev = matcher_device(m);
    if (dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
        CORB_CUDA(cudaFuncSetAttribute(k_pnp_solve<0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem<0, 1, 1>()));
        CORB_CUDA(cudaFuncSetAttribute(k_pnp_solve<0, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem<0, 8, 3>()));
        CORB_CUDA(cudaFuncSetAttribute(k_pnp_solve<1, 32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem<1, 32, 1>()));
        CORB_CUDA(cudaFuncSetAttribute(k_pnp_solve<1, 32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem<1, 32, 3>()));
        if (dev < 64) attr_set[dev].store(true, std::memory_order_release);
    }
    // Small batches leave the GPU mostly idle and the call is bound by the latency of ONE hypothesis: a team of 8 lanes per
    // hypothesis and three warps per item (one beta approximation each) shorten it. Large batches keep one thread per
    // hypothesis and one warp per record. CORB_PNP_TEAM=0/1 forces either.
    const char* env = getenv("CORB_PNP_TEAM");
    const bool team = env ? atoi(env) != 0 : n_hyp <= 2400;
    const dim3 gc((max_it + kCheckWarps - 1) / kCheckWarps, C);
    if (team)
        k_pnp_solve<0, 8, 3><<<dim3((max_it + 3) / 4, C), 3 * kSolveThreads, solve_smem<0, 8, 3>(), st>>>(D);
    else
        k_pnp_solve<0, 1, 1><<<dim3((max_it + kSolveThreads - 1) / kSolveThreads, C), kSolveThreads, solve_smem<0, 1, 1>(), st>>>(D);
    k_pnp_check<0><<<gc, kCheckWarps * 32, 0, st>>>(D);
    k_pnp_records<<<C, 32, 0, st>>>(D);
    if (team)  // blocks beyond a problem's record count exit
        k_pnp_solve<1, 32, 3><<<dim3(max_rec, C), 3 * kSolveThreads, solve_smem<1, 32, 3>(), st>>>(D);
    else
        k_pnp_solve<1, 32, 1><<<dim3(max_rec, C), kSolveThreads, solve_smem<1, 32, 1>(), st>>>(D);
    k_pnp_check<1><<<dim3((max_rec + kCheckWarps - 1) / kCheckWarps, C), kCheckWarps * 32, 0, st>>>(D);
    k_pnp_finalize<<<C, 32, 0, st>>>(D);
    CORB_CUDA(cudaGetLastError());
    CORB_CUDA(cudaMemcpyAsync(h + oRes, d + oRes, out_end - oRes, cudaMemcpyDeviceToHost, st));
    CORB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < C; i++) {
        const int c = live[i];
        const int* r = (const int*)(h + oRes) + 20 * i;
        corb_pnp_result& o = results[c];
        o.status = r[0]; o.no_more = r[1]; o.n_inliers = r[2]; o.iterations = r[3];
        memcpy(o.Tcw, r + 4, 64);
        if (inliers && inliers[c]) memcpy(inliers[c], h + oInl + hp[i].out_off, problems[c].n);
    }
    return CORB_OK;
}

}  // extern "C"
