// Descriptor matching and DBoW2 slice for sm_100a:
//   K7  256-bit Hamming (uint4 x 2 loads + __popc) and SearchByBoW x3   ORBmatcher.cc:162-423,657-790,1792-1808
//   K8  vocabulary-tree descent                                          TemplatedVocabulary.h:1218-1259
//   K9  sparse L1 BoW score, one query against many candidates           ScoringObject.cpp:23-68
// All integer results are bit-exact against oracle/match_oracle.cpp; the fp64 score reproduces the reference's
// summation order (ascending common word id).
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "common.cuh"

namespace corb {

__device__ __forceinline__ int ham256(const uint4& a0, const uint4& a1, const uint4* b) {
    const uint4 b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ------------------------------------------------------------------------------------------------ Hamming pairs
__global__ void __launch_bounds__(256) k_hamming_pairs(const uint4* __restrict__ A, const uint4* __restrict__ B,
                                                       const int2* __restrict__ pairs, int n, int* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int2 p = pairs[i];
    const uint4 a0 = A[2 * p.x], a1 = A[2 * p.x + 1];
    out[i] = ham256(a0, a1, B + 2 * p.y);
}

// ------------------------------------------------------------------------------------------------ SearchByBoW
struct BowCall {  // everything is a device pointer
    const uint4 *descA, *descB;
    const uint32_t *nodesA, *idxA, *nodesB, *idxB;
    const int32_t *offA, *offB;
    const uint8_t *validA, *validB;
    const float *angA, *angB;
    int32_t* match;     // [nOut]
    int32_t* nmatches;  // [1]
    uint8_t* matchedB;  // [nB]   scratch
    uint8_t* bin;       // [nOut] scratch
    int* hist;          // [32]   scratch: 30 rotation bins, [31] = block ticket
    int nA, nB, nnA, nnB, nOut;
};

constexpr int kThLow = 50, kHistoLength = 30;  // ORBmatcher.cc:37-39
constexpr int kNodeRegs = 4;                   // B features of a vocabulary node a lane keeps in registers (nodes up to 128 features)

__global__ void __launch_bounds__(256) k_bow_init(const BowCall* __restrict__ calls) {
    const BowCall c = calls[blockIdx.y];
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < c.nOut) c.match[i] = -1;
    if (i < c.nB) c.matchedB[i] = 0;
    if (i < 32) c.hist[i] = 0;
}

// One warp per FeatureVector node of side A. A feature of B belongs to exactly one node, so the reference's greedy
// "skip B features that are already matched" only couples A features of the same node: nodes are independent, the
// A features of one node are visited sequentially in FeatureVector order, and the inner loop over B (best, first
// position on ties, multiset second-best) is a warp reduction.
__global__ void __launch_bounds__(128) k_bow_match(const BowCall* __restrict__ calls, int variant, float nnratio, int check_ori) {
    __shared__ int s_keep[3];
    __shared__ int s_cnt[4];
    __shared__ int s_last;
    __shared__ uint4 s_ad[4][32][2];  // per warp: descriptors of the next 32 A features of its node
    __shared__ int s_ai[4][32];       //           their indices (-1: not a live MapPoint)
    const BowCall c = calls[blockIdx.y];
    const bool kfkf = variant == 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ga = blockIdx.x * 4 + warp;
    if (ga < c.nnA) {
        const uint32_t node = c.nodesA[ga];
        int lo = 0, hi = c.nnB;  // lower_bound of node in nodesB
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c.nodesB[mid] < node) lo = mid + 1; else hi = mid;
        }
        if (lo < c.nnB && c.nodesB[lo] == node) {
            const int b_beg = c.offB[lo], nbg = c.offB[lo + 1] - b_beg;
            const float factor = 1.0f / kHistoLength;
            if (nbg <= 32 * kNodeRegs) {
                // The node's B features live in registers (feature j = lane + 32 s in slot s of its lane): descriptor, index and
                // an "available" bit that replaces the matchedB bytes - a B feature belongs to exactly one node, so only this warp
                // ever looks at it. Per A feature that leaves one descriptor load (fetched one feature ahead) and the reduction.
                uint4 q0[kNodeRegs], q1[kNodeRegs];
                int bb[kNodeRegs];
                bool ok[kNodeRegs];
#pragma unroll
                for (int s = 0; s < kNodeRegs; s++) {
                    const int j = lane + 32 * s;
                    ok[s] = false; bb[s] = -1;
                    q0[s] = q1[s] = make_uint4(0u, 0u, 0u, 0u);
                    if (j < nbg) {
                        const int b = (int)c.idxB[b_beg + j];
                        bb[s] = b;
                        ok[s] = !(kfkf && c.validB && !c.validB[b]);
                        q0[s] = c.descB[2 * b];
                        q1[s] = c.descB[2 * b + 1];
                    }
                }
                const int i_beg = c.offA[ga], i_end = c.offA[ga + 1];
                for (int i1 = i_beg; i1 < i_end; i1++) {
                    const int t = (i1 - i_beg) & 31;
                    if (t == 0) {  // the next 32 A features of the node: every lane fetches one into the warp's shared-memory stage
                        __syncwarp();
                        const int ii = i1 + lane;
                        int ax = -1;
                        if (ii < i_end) {
                            ax = (int)c.idxA[ii];
                            if (c.validA && !c.validA[ax]) ax = -1;
                        }
                        s_ai[warp][lane] = ax;
                        if (ax >= 0) { s_ad[warp][lane][0] = c.descA[2 * ax]; s_ad[warp][lane][1] = c.descA[2 * ax + 1]; }
                        __syncwarp();
                    }
                    const int a = s_ai[warp][t];
                    if (a < 0) continue;
                    const uint4 a0 = s_ad[warp][t][0], a1 = s_ad[warp][t][1];
                    int b1 = 256, p1 = INT_MAX, b2 = 256, bi = -1;
#pragma unroll
                    for (int s = 0; s < kNodeRegs; s++) {  // ascending j per lane, like the scan over the node's list
                        if (ok[s]) {
                            const int d = __popc(a0.x ^ q0[s].x) + __popc(a0.y ^ q0[s].y) + __popc(a0.z ^ q0[s].z) + __popc(a0.w ^ q0[s].w) +
                                          __popc(a1.x ^ q1[s].x) + __popc(a1.y ^ q1[s].y) + __popc(a1.z ^ q1[s].z) + __popc(a1.w ^ q1[s].w);
                            if (d < b1) { b2 = b1; b1 = d; p1 = lane + 32 * s; bi = bb[s]; }
                            else if (d < b2) b2 = d;
                        }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        const int ob1 = __shfl_xor_sync(0xffffffffu, b1, o), op1 = __shfl_xor_sync(0xffffffffu, p1, o);
                        const int ob2 = __shfl_xor_sync(0xffffffffu, b2, o), obi = __shfl_xor_sync(0xffffffffu, bi, o);
                        const bool other = ob1 < b1 || (ob1 == b1 && op1 < p1);
                        const int loser = other ? b1 : ob1;
                        if (other) { b1 = ob1; p1 = op1; bi = obi; }
                        b2 = min(min(b2, ob2), loser);
                    }
                    const bool near = kfkf ? b1 < kThLow : b1 <= kThLow;  // :231 / :733
                    if (near && (float)b1 < __fmul_rn(nnratio, (float)b2)) {
#pragma unroll
                        for (int s = 0; s < kNodeRegs; s++)
                            if (p1 == lane + 32 * s) ok[s] = false;
                        if (lane == 0) {
                            const int slot = kfkf ? a : bi;
                            c.match[slot] = kfkf ? bi : a;
                            if (check_ori) {
                                float rot = __fsub_rn(c.angA[a], c.angB[bi]);
                                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                                int bin = (int)roundf(__fmul_rn(rot, factor));
                                if (bin == kHistoLength) bin = 0;
                                c.bin[slot] = (uint8_t)bin;
                                if (bin >= 0 && bin < kHistoLength) atomicAdd(&c.hist[bin], 1);
                            }
                        }
                    }
                }
            } else
            for (int i1 = c.offA[ga]; i1 < c.offA[ga + 1]; i1++) {
                const int a = (int)c.idxA[i1];
                if (c.validA && !c.validA[a]) continue;
                const uint4 a0 = c.descA[2 * a], a1 = c.descA[2 * a + 1];
                int b1 = 256, p1 = INT_MAX, b2 = 256, bi = -1;
                for (int j = lane; j < nbg; j += 32) {
                    const int b = (int)c.idxB[b_beg + j];
                    if (c.matchedB[b]) continue;
                    if (kfkf && c.validB && !c.validB[b]) continue;
                    const int d = ham256(a0, a1, c.descB + 2 * b);
                    if (d < b1) { b2 = b1; b1 = d; p1 = j; bi = b; }
                    else if (d < b2) b2 = d;
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const int ob1 = __shfl_xor_sync(0xffffffffu, b1, o), op1 = __shfl_xor_sync(0xffffffffu, p1, o);
                    const int ob2 = __shfl_xor_sync(0xffffffffu, b2, o), obi = __shfl_xor_sync(0xffffffffu, bi, o);
                    const bool other = ob1 < b1 || (ob1 == b1 && op1 < p1);
                    const int loser = other ? b1 : ob1;
                    if (other) { b1 = ob1; p1 = op1; bi = obi; }
                    b2 = min(min(b2, ob2), loser);
                }
                const bool near = kfkf ? b1 < kThLow : b1 <= kThLow;  // :231 / :733
                if (near && (float)b1 < __fmul_rn(nnratio, (float)b2)) {
                    if (lane == 0) {
                        c.matchedB[bi] = 1;
                        const int slot = kfkf ? a : bi;
                        c.match[slot] = kfkf ? bi : a;
                        if (check_ori) {
                            float rot = __fsub_rn(c.angA[a], c.angB[bi]);
                            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                            int bin = (int)roundf(__fmul_rn(rot, factor));
                            if (bin == kHistoLength) bin = 0;
                            c.bin[slot] = (uint8_t)bin;
                            if (bin >= 0 && bin < kHistoLength) atomicAdd(&c.hist[bin], 1);
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    // ---- the last block of this call applies the rotation-consistency filter and counts (:270-288)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&c.hist[31], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp == 0) {
        // the 30 bins arrive in one round trip (a lane each) and go to lane 0 through shuffles; the scan itself is sequential
        const int mine = check_ori && lane < kHistoLength ? __ldcg(&c.hist[lane]) : 0;
        int ind1 = -1, ind2 = -1, ind3 = -1;
        int max1 = 0, max2 = 0, max3 = 0;
#pragma unroll
        for (int i = 0; i < kHistoLength; i++) {  // ComputeThreeMaxima (:1746-1787)
            const int s = __shfl_sync(0xffffffffu, mine, i);
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (check_ori) {
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
        } else {
            ind1 = ind2 = ind3 = -1;
        }
        if (lane == 0) { s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3; }
    }
    __syncthreads();
    int cnt = 0;
    const int k0 = s_keep[0], k1 = s_keep[1], k2 = s_keep[2];
    for (int s0 = threadIdx.x; s0 < c.nOut; s0 += 4 * blockDim.x) {  // four entries per thread in flight: the loads first
        int mt[4], bn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int s = s0 + u * blockDim.x;
            mt[u] = s < c.nOut ? __ldcg(&c.match[s]) : -1;
            bn[u] = check_ori && s < c.nOut ? (int)__ldcg(reinterpret_cast<const unsigned char*>(&c.bin[s])) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (mt[u] < 0) continue;
            if (check_ori && bn[u] != k0 && bn[u] != k1 && bn[u] != k2) { c.match[s0 + u * blockDim.x] = -1; continue; }
            cnt++;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) *c.nmatches = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
}

// ------------------------------------------------------------------------------------------------ vocabulary
struct VocDev {
    const uint4* desc;        // [N*2] node descriptors, device order (children of a node are contiguous)
    const int* child_begin;   // [N]
    const uint8_t* child_cnt; // [N]
    const uint32_t* orig_id;  // [N] DBoW2 node id
    const uint32_t* word_id;  // [N]
    const double* weight;     // [N]
};

// 16 lanes per feature: each lane takes one child (k <= 16 in one round), argmin (distance, child order) by
// half-warp shuffles; first minimum wins like the strict `d < best_d` of the reference (:1244).
__global__ void __launch_bounds__(256) k_voc_transform(VocDev v, const uint4* __restrict__ desc, int n, int nid_level,
                                                       uint32_t* __restrict__ word, double* __restrict__ weight,
                                                       uint32_t* __restrict__ nid) {
    const int g = (blockIdx.x * 256 + threadIdx.x) >> 4;
    if (g >= n) return;
    const int sub = threadIdx.x & 15;
    const unsigned mask = 0xffffu << (threadIdx.x & 16);
    const uint4 q0 = desc[2 * g], q1 = desc[2 * g + 1];
    int cur = 0, level = 0;
    long long nid_out = nid_level <= 0 ? (long long)v.orig_id[0] : -1;
    do {
        ++level;
        const int cb = v.child_begin[cur], cc = v.child_cnt[cur];
        int best = INT_MAX, bc = INT_MAX;
        for (int ch = sub; ch < cc; ch += 16) {
            const int d = ham256(q0, q1, v.desc + 2 * (cb + ch));
            if (d < best) { best = d; bc = ch; }
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) {
            const int ob = __shfl_xor_sync(mask, best, o), oc = __shfl_xor_sync(mask, bc, o);
            if (ob < best || (ob == best && oc < bc)) { best = ob; bc = oc; }
        }
        cur = cb + bc;
        if (level == nid_level) nid_out = v.orig_id[cur];
    } while (v.child_cnt[cur] > 0);
    if (sub == 0) {
        word[g] = v.word_id[cur];
        weight[g] = v.weight[cur];
        nid[g] = (uint32_t)(nid_out < 0 ? v.orig_id[cur] : nid_out);
    }
}

// One CTA (128 threads) per candidate BowVector. Both vectors are sorted by word, so a thread takes a contiguous run of the
// candidate's words, finds its start in the query with one binary search and walks both lists like the merge of
// L1Scoring::score; the term fabs(vi-wi)-fabs(vi)-fabs(wi) of a common word (strictly negative: the weights are positive)
// goes to the word's slot in shared memory, 0 elsewhere. One thread then adds the non-zero slots strictly in ascending word
// order - the reference's summation order, so the fp64 bits are DBoW2's.
constexpr int kScoreThreads = 128, kScoreCap = 4096;
__device__ __forceinline__ void bow_score_one(const uint32_t* __restrict__ qw, const double* __restrict__ qv, int nq,
                                              const uint32_t* __restrict__ w_, const double* __restrict__ v_, int n, double* out) {
    __shared__ double terms[kScoreCap];
    const int tid = threadIdx.x;
    double score = 0.0;
    for (int base = 0; base < n; base += kScoreCap) {  // one round unless a vector has more than kScoreCap words
        const int m = min(kScoreCap, n - base);
        const int per = (m + kScoreThreads - 1) / kScoreThreads;
        const int i0 = base + tid * per, i1 = min(base + m, i0 + per);
        if (i0 < i1) {
            const uint32_t wf = w_[i0];
            int lo = 0, hi = nq;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (qw[mid] < wf) lo = mid + 1; else hi = mid;
            }
            for (int i = i0; i < i1; i++) {
                const uint32_t w = w_[i];
                while (lo < nq && qw[lo] < w) lo++;
                double term = 0.0;
                if (lo < nq && qw[lo] == w) {
                    const double vi = qv[lo], wi = v_[i];
                    term = __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi));
                }
                terms[i - base] = term;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int k = 0;
            for (; k + 8 <= m; k += 8) {  // loads first: only the additions of common words sit on the dependent chain
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; u++) t[u] = terms[k + u];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (t[u] != 0.0) score = __dadd_rn(score, t[u]);
            }
            for (; k < m; k++)
                if (terms[k] != 0.0) score = __dadd_rn(score, terms[k]);
        }
        __syncthreads();
    }
    if (tid == 0) *out = -score / 2.0;
}

__global__ void __launch_bounds__(kScoreThreads) k_bow_score(const uint32_t* __restrict__ qw, const double* __restrict__ qv, int nq,
                                                             const uint32_t* __restrict__ cw, const double* __restrict__ cv,
                                                             const int* __restrict__ coff, int ncand, double* __restrict__ scores) {
    const int cand = blockIdx.x;
    const int beg = coff[cand], end = coff[cand + 1];
    bow_score_one(qw, qv, nq, cw + beg, cv + beg, end - beg, scores + cand);
}

// ------------------------------------------------------------------------------------------------ host side
struct Arena {  // growable device buffer + pinned host mirror
    uint8_t* d = nullptr;
    uint8_t* h = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return CORB_OK;
        size_t ncap = std::max(bytes, cap * 2);
        ncap = align_up_sz(ncap, 1 << 16);
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        d = nullptr; h = nullptr; cap = 0;
        CORB_CUDA(cudaMalloc((void**)&d, ncap));
        CORB_CUDA(cudaMallocHost((void**)&h, ncap));
        cap = ncap;
        return CORB_OK;
    }
    void release() {
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        d = nullptr; h = nullptr; cap = 0;
    }
};

struct Packer {  // lays out 16-byte aligned blocks in an Arena
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off = align_up_sz(off + bytes, 16);
        return o;
    }
};

}  // namespace corb

using namespace corb;

struct corb_matcher {
    int device;
    cudaStream_t stream = nullptr;
    Arena arena;
    Arena dev_calls;  // for the device-resident batch form (call table + scratch)
    Arena proj;       // projection matchers (proj_match.cu)
    Arena pnp;        // EPnP-RANSAC batches (pnp.cu)
};

namespace corb {
int matcher_device(const corb_matcher* m) { return m->device; }
cudaStream_t matcher_stream(const corb_matcher* m) { return m->stream; }
int matcher_proj_reserve(corb_matcher* m, size_t bytes, uint8_t** d, uint8_t** h) {
    int rc = m->proj.reserve(bytes);
    if (rc != CORB_OK) return rc;
    *d = m->proj.d;
    *h = m->proj.h;
    return CORB_OK;
}
int matcher_pnp_reserve(corb_matcher* m, size_t bytes, uint8_t** d, uint8_t** h) {
    int rc = m->pnp.reserve(bytes);
    if (rc != CORB_OK) return rc;
    *d = m->pnp.d;
    *h = m->pnp.h;
    return CORB_OK;
}
}  // namespace corb

struct corb_voc {
    int device;
    int k, L, scoring, weighting, n_nodes, n_words;
    VocDev dev;
    std::vector<void*> allocs;
    cudaStream_t stream = nullptr;
    Arena arena;
};

extern "C" {

int corb_matcher_create(int device, corb_matcher** out) {
    CORB_CHECK(out, CORB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    CORB_CUDA(cudaGetDeviceCount(&ndev));
    CORB_CHECK(device >= 0 && device < ndev, CORB_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    CORB_CUDA(cudaSetDevice(device));
    corb_matcher* m = new corb_matcher;
    m->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
        delete m;
        return CORB_ERR_CUDA;
    }
    *out = m;
    return CORB_OK;
}

void corb_matcher_destroy(corb_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    m->arena.release();
    m->dev_calls.release();
    m->proj.release();
    m->pnp.release();
    delete m;
}

int corb_matcher_sync(corb_matcher* m) {
    CORB_CHECK(m, CORB_ERR_INVALID, "matcher is NULL");
    CORB_CUDA(cudaSetDevice(m->device));
    CORB_CUDA(cudaStreamSynchronize(m->stream));
    return CORB_OK;
}
void* corb_matcher_stream(const corb_matcher* m) { return m ? (void*)m->stream : nullptr; }

int corb_hamming_pairs(corb_matcher* m, const uint8_t* A, int nA, const uint8_t* B, int nB, const int32_t* pairs, int n,
                       int32_t* out) {
    CORB_CHECK(m && A && B && pairs && out && nA >= 0 && nB >= 0 && n >= 0, CORB_ERR_INVALID, "bad argument");
    if (n == 0) return CORB_OK;
    for (int i = 0; i < n; i++)
        CORB_CHECK(pairs[2 * i] >= 0 && pairs[2 * i] < nA && pairs[2 * i + 1] >= 0 && pairs[2 * i + 1] < nB, CORB_ERR_INVALID,
                   "pair %d out of range", i);
    CORB_CUDA(cudaSetDevice(m->device));
    Packer p;
    const size_t oA = p.take((size_t)nA * 32), oB = p.take((size_t)nB * 32), oP = p.take((size_t)n * 8), oO = p.take((size_t)n * 4);
    int rc = m->arena.reserve(p.off);
    if (rc != CORB_OK) return rc;
    memcpy(m->arena.h + oA, A, (size_t)nA * 32);
    memcpy(m->arena.h + oB, B, (size_t)nB * 32);
    memcpy(m->arena.h + oP, pairs, (size_t)n * 8);
    CORB_CUDA(cudaMemcpyAsync(m->arena.d, m->arena.h, oO, cudaMemcpyHostToDevice, m->stream));
    k_hamming_pairs<<<(n + 255) / 256, 256, 0, m->stream>>>((const uint4*)(m->arena.d + oA), (const uint4*)(m->arena.d + oB),
                                                            (const int2*)(m->arena.d + oP), n, (int*)(m->arena.d + oO));
    CORB_CUDA(cudaGetLastError());
    CORB_CUDA(cudaMemcpyAsync(m->arena.h + oO, m->arena.d + oO, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    CORB_CUDA(cudaStreamSynchronize(m->stream));
    memcpy(out, m->arena.h + oO, (size_t)n * 4);
    return CORB_OK;
}

static int validate_side(const corb_bow_side& s, const char* name, int call, bool need_angles) {
    CORB_CHECK(s.n >= 0 && s.fv_n >= 0, CORB_ERR_INVALID, "call %d side %s: negative size", call, name);
    CORB_CHECK(s.n == 0 || s.desc, CORB_ERR_INVALID, "call %d side %s: desc is NULL", call, name);
    CORB_CHECK(s.fv_n == 0 || (s.fv_nodes && s.fv_off && s.fv_idx), CORB_ERR_INVALID, "call %d side %s: FeatureVector is NULL", call, name);
    CORB_CHECK(!need_angles || s.n == 0 || s.angles, CORB_ERR_INVALID, "call %d side %s: angles required with check_ori", call, name);
    return CORB_OK;
}

static int launch_bow(corb_matcher* m, const BowCall* d_calls, int ncalls, int max_nnA, int max_n, int variant, float nnratio,
                      int check_ori) {
    dim3 gi((std::max(max_n, 32) + 255) / 256, ncalls);
    k_bow_init<<<gi, 256, 0, m->stream>>>(d_calls);
    dim3 gm(std::max(1, (max_nnA + 3) / 4), ncalls);
    k_bow_match<<<gm, 128, 0, m->stream>>>(d_calls, variant, nnratio, check_ori);
    CORB_CUDA(cudaGetLastError());
    return CORB_OK;
}

int corb_bow_match_batch(corb_matcher* m, int variant, int ncalls, const corb_bow_side* A, const corb_bow_side* B, float nnratio,
                         int check_ori, int32_t* const* match, int32_t* nmatches) {
    CORB_CHECK(m && A && B && match && nmatches && ncalls >= 0, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(variant >= 0 && variant <= 2, CORB_ERR_INVALID, "unknown variant %d", variant);
    if (ncalls == 0) return CORB_OK;
    CORB_CUDA(cudaSetDevice(m->device));
    const bool kfkf = variant == 2;
    // ---- validate the host-side structure (indices the kernel will dereference)
    for (int i = 0; i < ncalls; i++) {
        int rc = validate_side(A[i], "A", i, check_ori);
        if (rc != CORB_OK) return rc;
        rc = validate_side(B[i], "B", i, check_ori);
        if (rc != CORB_OK) return rc;
        CORB_CHECK(match[i] || (kfkf ? A[i].n : B[i].n) == 0, CORB_ERR_INVALID, "call %d: match is NULL", i);
        for (int s = 0; s < 2; s++) {
            const corb_bow_side& sd = s ? B[i] : A[i];
            for (int g = 0; g < sd.fv_n; g++) {
                CORB_CHECK(sd.fv_off[g] <= sd.fv_off[g + 1] && sd.fv_off[g] >= 0, CORB_ERR_INVALID, "call %d: fv_off not monotone", i);
                CORB_CHECK(g == 0 || sd.fv_nodes[g - 1] < sd.fv_nodes[g], CORB_ERR_INVALID, "call %d: fv_nodes not ascending", i);
            }
            const int ne = sd.fv_n ? sd.fv_off[sd.fv_n] : 0;
            for (int e = 0; e < ne; e++)
                CORB_CHECK((int)sd.fv_idx[e] < sd.n, CORB_ERR_INVALID, "call %d: feature index %u out of range", i, sd.fv_idx[e]);
        }
    }
    // ---- pack: [inputs of all calls | call table | scratch | outputs]
    Packer p;
    struct Off { size_t descA, descB, nodesA, offA, idxA, nodesB, offB, idxB, validA, validB, angA, angB, matchedB, bin, hist, match; };
    std::vector<Off> off(ncalls);
    for (int i = 0; i < ncalls; i++) {
        const corb_bow_side &a = A[i], &b = B[i];
        const int neA = a.fv_n ? a.fv_off[a.fv_n] : 0, neB = b.fv_n ? b.fv_off[b.fv_n] : 0;
        Off& o = off[i];
        o.descA = p.take((size_t)a.n * 32); o.descB = p.take((size_t)b.n * 32);
        o.nodesA = p.take((size_t)a.fv_n * 4); o.offA = p.take((size_t)(a.fv_n + 1) * 4); o.idxA = p.take((size_t)neA * 4);
        o.nodesB = p.take((size_t)b.fv_n * 4); o.offB = p.take((size_t)(b.fv_n + 1) * 4); o.idxB = p.take((size_t)neB * 4);
        o.validA = a.valid ? p.take(a.n) : (size_t)-1;
        o.validB = (kfkf && b.valid) ? p.take(b.n) : (size_t)-1;
        o.angA = check_ori ? p.take((size_t)a.n * 4) : (size_t)-1;
        o.angB = check_ori ? p.take((size_t)b.n * 4) : (size_t)-1;
    }
    const size_t o_calls = p.take(sizeof(BowCall) * ncalls);
    const size_t in_bytes = p.off;
    int max_nnA = 0, max_n = 0;
    for (int i = 0; i < ncalls; i++) {
        const int nOut = kfkf ? A[i].n : B[i].n;
        off[i].matchedB = p.take(B[i].n); off[i].bin = p.take(nOut); off[i].hist = p.take(32 * 4);
        max_nnA = std::max(max_nnA, A[i].fv_n);
        max_n = std::max(max_n, std::max(nOut, B[i].n));
    }
    const size_t o_out = p.off;
    for (int i = 0; i < ncalls; i++) off[i].match = p.take((size_t)(kfkf ? A[i].n : B[i].n) * 4);
    const size_t o_nm = p.take((size_t)ncalls * 4);
    int rc = m->arena.reserve(p.off);
    if (rc != CORB_OK) return rc;
    uint8_t *h = m->arena.h, *d = m->arena.d;
    BowCall* calls = (BowCall*)(h + o_calls);
    for (int i = 0; i < ncalls; i++) {
        const corb_bow_side &a = A[i], &b = B[i];
        const Off& o = off[i];
        const int neA = a.fv_n ? a.fv_off[a.fv_n] : 0, neB = b.fv_n ? b.fv_off[b.fv_n] : 0;
        if (a.n) memcpy(h + o.descA, a.desc, (size_t)a.n * 32);
        if (b.n) memcpy(h + o.descB, b.desc, (size_t)b.n * 32);
        if (a.fv_n) { memcpy(h + o.nodesA, a.fv_nodes, (size_t)a.fv_n * 4); memcpy(h + o.offA, a.fv_off, (size_t)(a.fv_n + 1) * 4); memcpy(h + o.idxA, a.fv_idx, (size_t)neA * 4); }
        else memset(h + o.offA, 0, 4);
        if (b.fv_n) { memcpy(h + o.nodesB, b.fv_nodes, (size_t)b.fv_n * 4); memcpy(h + o.offB, b.fv_off, (size_t)(b.fv_n + 1) * 4); memcpy(h + o.idxB, b.fv_idx, (size_t)neB * 4); }
        else memset(h + o.offB, 0, 4);
        if (o.validA != (size_t)-1) memcpy(h + o.validA, a.valid, a.n);
        if (o.validB != (size_t)-1) memcpy(h + o.validB, b.valid, b.n);
        if (check_ori) { if (a.n) memcpy(h + o.angA, a.angles, (size_t)a.n * 4); if (b.n) memcpy(h + o.angB, b.angles, (size_t)b.n * 4); }
        BowCall& c = calls[i];
        c.descA = (const uint4*)(d + o.descA); c.descB = (const uint4*)(d + o.descB);
        c.nodesA = (const uint32_t*)(d + o.nodesA); c.offA = (const int32_t*)(d + o.offA); c.idxA = (const uint32_t*)(d + o.idxA);
        c.nodesB = (const uint32_t*)(d + o.nodesB); c.offB = (const int32_t*)(d + o.offB); c.idxB = (const uint32_t*)(d + o.idxB);
        c.validA = o.validA != (size_t)-1 ? d + o.validA : nullptr;
        c.validB = o.validB != (size_t)-1 ? d + o.validB : nullptr;
        c.angA = check_ori ? (const float*)(d + o.angA) : nullptr;
        c.angB = check_ori ? (const float*)(d + o.angB) : nullptr;
        c.match = (int32_t*)(d + o.match); c.nmatches = (int32_t*)(d + o_nm) + i;
        c.matchedB = d + o.matchedB; c.bin = d + o.bin; c.hist = (int*)(d + o.hist);
        c.nA = a.n; c.nB = b.n; c.nnA = a.fv_n; c.nnB = b.fv_n; c.nOut = kfkf ? a.n : b.n;
    }
    CORB_CUDA(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, m->stream));
    rc = launch_bow(m, (const BowCall*)(d + o_calls), ncalls, max_nnA, max_n, variant, nnratio, check_ori);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaMemcpyAsync(h + o_out, d + o_out, p.off - o_out, cudaMemcpyDeviceToHost, m->stream));
    CORB_CUDA(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < ncalls; i++) {
        const int nOut = kfkf ? A[i].n : B[i].n;
        if (nOut) memcpy(match[i], h + off[i].match, (size_t)nOut * 4);
        nmatches[i] = ((const int32_t*)(h + o_nm))[i];
    }
    return CORB_OK;
}

int corb_bow_match(corb_matcher* m, int variant, const corb_bow_side* A, const corb_bow_side* B, float nnratio, int check_ori,
                   int32_t* match, int32_t* nmatches) {
    CORB_CHECK(A && B && nmatches, CORB_ERR_INVALID, "bad argument");
    int32_t* mp = match;
    return corb_bow_match_batch(m, variant, 1, A, B, nnratio, check_ori, &mp, nmatches);
}

int corb_bow_match_batch_device(corb_matcher* m, int variant, int ncalls, const corb_bow_side* A, const corb_bow_side* B,
                                float nnratio, int check_ori, int32_t* const* d_match, int32_t* d_nmatches) {
    CORB_CHECK(m && A && B && d_match && d_nmatches && ncalls >= 1, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(variant >= 0 && variant <= 2, CORB_ERR_INVALID, "unknown variant %d", variant);
    CORB_CUDA(cudaSetDevice(m->device));
    const bool kfkf = variant == 2;
    Packer p;
    const size_t o_calls = p.take(sizeof(BowCall) * ncalls);
    const size_t in_bytes = p.off;
    std::vector<size_t> oM(ncalls), oB(ncalls), oH(ncalls);
    int max_nnA = 0, max_n = 0;
    for (int i = 0; i < ncalls; i++) {
        const int nOut = kfkf ? A[i].n : B[i].n;
        CORB_CHECK(((uintptr_t)A[i].desc & 15) == 0 && ((uintptr_t)B[i].desc & 15) == 0, CORB_ERR_INVALID,
                   "call %d: device descriptors must be 16-byte aligned", i);
        CORB_CHECK(!check_ori || (A[i].angles && B[i].angles), CORB_ERR_INVALID, "call %d: angles required", i);
        oM[i] = p.take(B[i].n); oB[i] = p.take(nOut); oH[i] = p.take(32 * 4);
        max_nnA = std::max(max_nnA, A[i].fv_n);
        max_n = std::max(max_n, std::max(nOut, B[i].n));
    }
    int rc = m->dev_calls.reserve(p.off);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaStreamSynchronize(m->stream));  // the pinned call table may still be in flight from the previous batch
    BowCall* calls = (BowCall*)(m->dev_calls.h + o_calls);
    uint8_t* d = m->dev_calls.d;
    for (int i = 0; i < ncalls; i++) {
        const corb_bow_side &a = A[i], &b = B[i];
        BowCall& c = calls[i];
        c.descA = (const uint4*)a.desc; c.descB = (const uint4*)b.desc;
        c.nodesA = a.fv_nodes; c.offA = a.fv_off; c.idxA = a.fv_idx;
        c.nodesB = b.fv_nodes; c.offB = b.fv_off; c.idxB = b.fv_idx;
        c.validA = a.valid; c.validB = kfkf ? b.valid : nullptr;
        c.angA = a.angles; c.angB = b.angles;
        c.match = d_match[i]; c.nmatches = d_nmatches + i;
        c.matchedB = d + oM[i]; c.bin = d + oB[i]; c.hist = (int*)(d + oH[i]);
        c.nA = a.n; c.nB = b.n; c.nnA = a.fv_n; c.nnB = b.fv_n; c.nOut = kfkf ? a.n : b.n;
    }
    CORB_CUDA(cudaMemcpyAsync(d, m->dev_calls.h, in_bytes, cudaMemcpyHostToDevice, m->stream));
    return launch_bow(m, (const BowCall*)(d + o_calls), ncalls, max_nnA, max_n, variant, nnratio, check_ori);
}

// ------------------------------------------------------------------------------------------------ vocabulary
static int voc_build(int k, int L, int scoring, int weighting, int n, const int32_t* parent, const uint8_t* is_leaf,
                     const uint8_t* desc, const double* weight, int device, corb_voc** out) {
    CORB_CHECK(out, CORB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    CORB_CHECK(k >= 1 && k <= 255 && L >= 1 && n >= 1 && parent && is_leaf && desc && weight, CORB_ERR_INVALID, "bad vocabulary arguments");
    CORB_CHECK(scoring == 0, CORB_ERR_UNSUPPORTED, "only L1_NORM scoring (0) is supported, got %d", scoring);
    CORB_CHECK(weighting == 0 || weighting == 1, CORB_ERR_UNSUPPORTED, "only TF_IDF (0) / TF (1) weighting is supported, got %d", weighting);
    const int N = n + 1;
    std::vector<std::vector<int>> children(N);
    for (int i = 1; i < N; i++) {
        CORB_CHECK(parent[i - 1] >= 0 && parent[i - 1] < i, CORB_ERR_INVALID, "node %d: parent %d is not an earlier node", i, parent[i - 1]);
        children[parent[i - 1]].push_back(i);
    }
    CORB_CHECK(!children[0].empty(), CORB_ERR_INVALID, "root has no children");
    // device order: breadth first, children of a node contiguous and in file order
    std::vector<int> order;  // device index -> DBoW2 id
    order.reserve(N);
    order.push_back(0);
    std::vector<int> cbeg(N, 0);
    std::vector<uint8_t> ccnt(N, 0);
    for (size_t head = 0; head < order.size(); head++) {
        const int id = order[head];
        CORB_CHECK(children[id].size() <= 255, CORB_ERR_UNSUPPORTED, "node %d has more than 255 children", id);
        cbeg[head] = (int)order.size();
        ccnt[head] = (uint8_t)children[id].size();
        for (int c : children[id]) order.push_back(c);
    }
    CORB_CHECK((int)order.size() == N, CORB_ERR_INVALID, "vocabulary tree is not connected");
    std::vector<uint8_t> ddesc((size_t)N * 32, 0);
    std::vector<uint32_t> orig(N), word(N, 0);
    std::vector<double> wgt(N, 0.0);
    std::vector<int> word_of(N, 0);
    int n_words = 0;
    for (int i = 1; i < N; i++)
        if (is_leaf[i - 1]) word_of[i] = n_words++;  // word ids in file order (:1407-1413)
    for (int di = 0; di < N; di++) {
        const int id = order[di];
        orig[di] = (uint32_t)id;
        if (id > 0) {
            memcpy(&ddesc[(size_t)di * 32], desc + (size_t)(id - 1) * 32, 32);
            wgt[di] = weight[id - 1];
            word[di] = (uint32_t)word_of[id];
        }
    }
    int ndev = 0;
    CORB_CUDA(cudaGetDeviceCount(&ndev));
    CORB_CHECK(device >= 0 && device < ndev, CORB_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    CORB_CUDA(cudaSetDevice(device));
    corb_voc* v = new corb_voc;
    v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->n_nodes = N; v->n_words = n_words;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> cudaError_t {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes + 256);
        if (e != cudaSuccess) return e;
        v->allocs.push_back(q);
        *dst = q;
        return cudaMemcpy(q, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = up(ddesc.data(), ddesc.size(), (const void**)&v->dev.desc);
    if (e == cudaSuccess) e = up(cbeg.data(), (size_t)N * 4, (const void**)&v->dev.child_begin);
    if (e == cudaSuccess) e = up(ccnt.data(), (size_t)N, (const void**)&v->dev.child_cnt);
    if (e == cudaSuccess) e = up(orig.data(), (size_t)N * 4, (const void**)&v->dev.orig_id);
    if (e == cudaSuccess) e = up(word.data(), (size_t)N * 4, (const void**)&v->dev.word_id);
    if (e == cudaSuccess) e = up(wgt.data(), (size_t)N * 8, (const void**)&v->dev.weight);
    if (e != cudaSuccess) {
        set_error("vocabulary upload failed: %s", cudaGetErrorString(e));
        corb_voc_destroy(v);
        return CORB_ERR_CUDA;
    }
    *out = v;
    return CORB_OK;
}

int corb_voc_create(int k, int L, int scoring, int weighting, int n, const int32_t* parent, const uint8_t* is_leaf,
                    const uint8_t* desc, const double* weight, int device, corb_voc** out) {
    return voc_build(k, L, scoring, weighting, n, parent, is_leaf, desc, weight, device, out);
}

int corb_voc_load_text(const char* path, int device, corb_voc** out) {
    CORB_CHECK(path && out, CORB_ERR_INVALID, "bad argument");
    FILE* f = fopen(path, "r");
    CORB_CHECK(f, CORB_ERR_IO, "cannot open %s", path);
    std::vector<char> line(1 << 16);
    int k = 0, L = 0, n1 = 0, n2 = 0;
    if (!fgets(line.data(), (int)line.size(), f) || sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 ||
        L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {  // same sanity test as the reference (:1361)
        fclose(f);
        set_error("%s is not a DBoW2 text vocabulary", path);
        return CORB_ERR_IO;
    }
    std::vector<int32_t> parent;
    std::vector<uint8_t> leaf, desc;
    std::vector<double> weight;
    while (fgets(line.data(), (int)line.size(), f)) {
        char *p = line.data(), *e;
        const long pid = strtol(p, &e, 10);
        if (e == p) continue;  // blank line: ignored (the reference would append a node with an uninitialised descriptor)
        p = e;
        const long isleaf = strtol(p, &e, 10);
        p = e;
        for (int i = 0; i < 32; i++) {
            const long b = strtol(p, &e, 10);
            p = e;
            desc.push_back((uint8_t)b);
        }
        const double w = strtod(p, &e);
        parent.push_back((int32_t)pid);
        leaf.push_back(isleaf > 0);
        weight.push_back(w);
    }
    fclose(f);
    CORB_CHECK(!parent.empty(), CORB_ERR_IO, "%s holds no nodes", path);
    return voc_build(k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), device, out);
}

void corb_voc_destroy(corb_voc* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    for (void* p : v->allocs) cudaFree(p);
    v->arena.release();
    delete v;
}

int corb_voc_info(const corb_voc* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words) {
    CORB_CHECK(v, CORB_ERR_INVALID, "vocabulary is NULL");
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (scoring) *scoring = v->scoring;
    if (weighting) *weighting = v->weighting;
    if (n_nodes) *n_nodes = v->n_nodes;
    if (n_words) *n_words = v->n_words;
    return CORB_OK;
}

int corb_voc_transform_features(corb_voc* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                                uint32_t* node_id) {
    CORB_CHECK(v && n >= 0 && (n == 0 || (desc && word_id && weight && node_id)), CORB_ERR_INVALID, "bad argument");
    if (n == 0) return CORB_OK;
    CORB_CUDA(cudaSetDevice(v->device));
    Packer p;
    const size_t oD = p.take((size_t)n * 32), oW = p.take((size_t)n * 4), oN = p.take((size_t)n * 4), oWt = p.take((size_t)n * 8);
    int rc = v->arena.reserve(p.off);
    if (rc != CORB_OK) return rc;
    uint8_t *h = v->arena.h, *d = v->arena.d;
    memcpy(h + oD, desc, (size_t)n * 32);
    CORB_CUDA(cudaMemcpyAsync(d + oD, h + oD, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    k_voc_transform<<<(n * 16 + 255) / 256, 256, 0, v->stream>>>(v->dev, (const uint4*)(d + oD), n, v->L - levelsup,
                                                                 (uint32_t*)(d + oW), (double*)(d + oWt), (uint32_t*)(d + oN));
    CORB_CUDA(cudaGetLastError());
    CORB_CUDA(cudaMemcpyAsync(h + oW, d + oW, p.off - oW, cudaMemcpyDeviceToHost, v->stream));
    CORB_CUDA(cudaStreamSynchronize(v->stream));
    memcpy(word_id, h + oW, (size_t)n * 4);
    memcpy(node_id, h + oN, (size_t)n * 4);
    memcpy(weight, h + oWt, (size_t)n * 8);
    return CORB_OK;
}

int corb_voc_transform(corb_voc* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_words, double* bow_vals, int* n_bow,
                       uint32_t* fv_nodes, int32_t* fv_off, uint32_t* fv_idx, int* n_fv) {
    CORB_CHECK(v && n_bow && n_fv && fv_off, CORB_ERR_INVALID, "bad argument");
    std::vector<uint32_t> word(n), node(n);
    std::vector<double> wt(n);
    int rc = corb_voc_transform_features(v, desc, n, levelsup, word.data(), wt.data(), node.data());
    if (rc != CORB_OK) return rc;
    // BowVector::addWeight / FeatureVector::addFeature in feature order, then L1 normalisation in word order
    // (TemplatedVocabulary.h:1147-1192, BowVector.cpp:34-84): host-side so every fp64 sum has the reference's order.
    std::map<uint32_t, double> bow;
    std::map<uint32_t, std::vector<uint32_t>> fv;
    for (int i = 0; i < n; i++) {
        if (wt[i] > 0) {
            const double w = wt[i];  // the node weight: idf for TF_IDF, 1 for TF (set when the vocabulary was trained)
            auto it = bow.lower_bound(word[i]);
            if (it != bow.end() && it->first == word[i]) it->second += w;
            else bow.insert(it, {word[i], w});
            fv[node[i]].push_back((uint32_t)i);
        }
    }
    double norm = 0.0;
    for (auto& kv : bow) norm += fabs(kv.second);
    if (norm > 0.0)
        for (auto& kv : bow) kv.second /= norm;
    int mcount = 0;
    for (auto& kv : bow) {
        bow_words[mcount] = kv.first;
        bow_vals[mcount] = kv.second;
        mcount++;
    }
    int g = 0, o = 0;
    for (auto& kv : fv) {
        fv_nodes[g] = kv.first;
        fv_off[g] = o;
        for (uint32_t i : kv.second) fv_idx[o++] = i;
        g++;
    }
    fv_off[g] = o;
    *n_bow = mcount;
    *n_fv = g;
    return CORB_OK;
}

int corb_bow_score_batch(corb_voc* v, const uint32_t* q_words, const double* q_vals, int nq, int ncand,
                         const uint32_t* const* c_words, const double* const* c_vals, const int32_t* c_n, double* scores) {
    CORB_CHECK(v && nq >= 0 && ncand >= 0 && (nq == 0 || (q_words && q_vals)), CORB_ERR_INVALID, "bad argument");
    if (ncand == 0) return CORB_OK;
    CORB_CHECK(c_words && c_vals && c_n && scores, CORB_ERR_INVALID, "bad argument");
    CORB_CUDA(cudaSetDevice(v->device));
    size_t total = 0;
    for (int i = 0; i < ncand; i++) {
        CORB_CHECK(c_n[i] >= 0 && (c_n[i] == 0 || (c_words[i] && c_vals[i])), CORB_ERR_INVALID, "candidate %d is malformed", i);
        total += c_n[i];
    }
    Packer p;
    const size_t oQW = p.take((size_t)nq * 4), oQV = p.take((size_t)nq * 8), oCW = p.take(total * 4), oCV = p.take(total * 8),
                 oOff = p.take((size_t)(ncand + 1) * 4), oS = p.take((size_t)ncand * 8);
    int rc = v->arena.reserve(p.off);
    if (rc != CORB_OK) return rc;
    uint8_t *h = v->arena.h, *d = v->arena.d;
    if (nq) { memcpy(h + oQW, q_words, (size_t)nq * 4); memcpy(h + oQV, q_vals, (size_t)nq * 8); }
    int* offs = (int*)(h + oOff);
    size_t run = 0;
    for (int i = 0; i < ncand; i++) {
        offs[i] = (int)run;
        if (c_n[i]) {
            memcpy(h + oCW + run * 4, c_words[i], (size_t)c_n[i] * 4);
            memcpy(h + oCV + run * 8, c_vals[i], (size_t)c_n[i] * 8);
        }
        run += c_n[i];
    }
    offs[ncand] = (int)run;
    CORB_CUDA(cudaMemcpyAsync(d, h, oS, cudaMemcpyHostToDevice, v->stream));
    k_bow_score<<<ncand, kScoreThreads, 0, v->stream>>>((const uint32_t*)(d + oQW), (const double*)(d + oQV), nq,
                                                                 (const uint32_t*)(d + oCW), (const double*)(d + oCV),
                                                                 (const int*)(d + oOff), ncand, (double*)(d + oS));
    CORB_CUDA(cudaGetLastError());
    CORB_CUDA(cudaMemcpyAsync(h + oS, d + oS, (size_t)ncand * 8, cudaMemcpyDeviceToHost, v->stream));
    CORB_CUDA(cudaStreamSynchronize(v->stream));
    memcpy(scores, h + oS, (size_t)ncand * 8);
    return CORB_OK;
}


}  // extern "C"

// ------------------------------------------------------------------------------------------------ device-resident BoW record
// A frame's / keyframe's matching record kept in HBM: descriptors, keypoint angles, BowVector and FeatureVector, built from
// the extractor's device-resident results without a trip through the host (Frame::ComputeBoW, Frame.cc:399-406, is
// transform() on the descriptors the extractor just produced). The std::map bookkeeping of BowVector::addWeight /
// FeatureVector::addFeature / normalize (BowVector.cpp:34-84, FeatureVector.cpp:31-45) becomes, in one CTA: a bitonic sort
// of (word, feature) keys, per-word sums in feature order, the L1 norm summed in ascending word order by one thread (the
// order the reference's loop has: the fp64 bits are the same), and a second sort of (node, feature) keys for the CSR.
namespace corb {

constexpr int kBowCap = 4096;  // features per record (shared-memory sort)

struct BowStoreDev {
    uint8_t* desc;      // [cap * 32]
    float* angles;      // [cap]
    uint32_t* word;     // [cap] per feature
    double* weight;     // [cap]
    uint32_t* node;     // [cap]
    uint32_t* bow_words; double* bow_vals;              // [cap]
    uint32_t* fv_nodes; int32_t* fv_off; uint32_t* fv_idx;  // [cap], [cap + 1], [cap]
    int32_t* counts;    // [4]: n_bow, n_fv, n_fv_idx, n
};

__global__ void __launch_bounds__(256) k_bow_gather(const corb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, int n,
                                                    BowStoreDev st) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) st.angles[i] = kps ? kps[i].angle : 0.f;
    const uint4* s4 = reinterpret_cast<const uint4*>(desc);
    uint4* d4 = reinterpret_cast<uint4*>(st.desc);
    if (i < 2 * n && desc != st.desc) d4[i] = s4[i];
    if (i + 256 * gridDim.x < 2 * n && desc != st.desc) d4[i + 256 * gridDim.x] = s4[i + 256 * gridDim.x];
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* a, int N, int tid, int nt) {
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < N / 2; t += nt) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // the lower index of the pair
                const int p = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long x = a[i], y = a[p];
                if ((x > y) == up) { a[i] = y; a[p] = x; }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(1024) k_bow_build(BowStoreDev st, int n) {
    extern __shared__ unsigned long long keys[];  // [N] | double vals[N] | int flags[N + 1]
    int N = 32;
    while (N < n) N <<= 1;
    double* vals = reinterpret_cast<double*>(keys + N);
    int* flag = reinterpret_cast<int*>(vals + N);
    __shared__ int s_total[1];
    __shared__ double s_norm;
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long kNone = ~0ull;
    {   // CTA 0: BowVector from (word, feature); CTA 1: FeatureVector from (node, feature) - independent, side by side
        const int pass = blockIdx.x;
        for (int i = tid; i < N; i += nt)
            keys[i] = (i < n && st.weight[i] > 0.0) ? ((unsigned long long)(pass ? st.node[i] : st.word[i]) << 32 | (unsigned)i) : kNone;
        __syncthreads();
        bitonic_sort_u64(keys, N, tid, nt);
        // heads of runs with the same upper word; valid keys come first
        for (int i = tid; i < N; i += nt) flag[i] = keys[i] != kNone && (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32));
        __syncthreads();
        // exclusive scan of the flags (N <= 4096: each thread scans a run of N / nt, then a block scan of the run totals)
        {
            const int per = (N + nt - 1) / nt;
            const int b0 = tid * per;
            int sum = 0;
            for (int i = b0; i < min(b0 + per, N); i++) sum += flag[i];
            __shared__ int wsum[32];
            int inc = sum;
            const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) wsum[wid] = inc;
            __syncthreads();
            if (wid == 0) {
                int w = lane < (nt >> 5) ? wsum[lane] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, w, o);
                    if (lane >= o) w += u;
                }
                wsum[lane] = w;
            }
            __syncthreads();
            int run = (wid ? wsum[wid - 1] : 0) + inc - sum;
            for (int i = b0; i < min(b0 + per, N); i++) {
                const int f = flag[i];
                flag[i] = f ? run : -1;  // index of the run this head starts, -1 for non-heads
                run += f;
            }
            if (tid == nt - 1) s_total[0] = run;
            __syncthreads();
        }
        if (pass == 0) {
            for (int i = tid; i < N; i += nt) {
                const int u = flag[i];
                if (u < 0) continue;
                const unsigned w = (unsigned)(keys[i] >> 32);
                double sum = 0.0;  // addWeight in feature order (keys are sorted by feature index inside a word)
                for (int j = i; j < N && keys[j] != kNone && (unsigned)(keys[j] >> 32) == w; j++) sum += st.weight[(unsigned)keys[j]];
                st.bow_words[u] = w;
                vals[u] = sum;
            }
            __syncthreads();
            const int nb = s_total[0];
            if (tid == 0) st.counts[0] = nb;
            if (tid == 0) {  // BowVector::normalize(L1): norm += fabs(value) in ascending word order
                double norm = 0.0;
                for (int u = 0; u < nb; u++) norm += fabs(vals[u]);
                s_norm = norm;
            }
            __syncthreads();
            const double norm = s_norm;
            for (int u = tid; u < nb; u += nt) st.bow_vals[u] = norm > 0.0 ? vals[u] / norm : vals[u];
            __syncthreads();
        } else {
            for (int i = tid; i < N; i += nt) {
                if (keys[i] == kNone) continue;
                st.fv_idx[i] = (unsigned)keys[i];
                const int u = flag[i];
                if (u >= 0) { st.fv_nodes[u] = (unsigned)(keys[i] >> 32); st.fv_off[u] = i; }
            }
            __syncthreads();
            if (tid == 0) {
                int cnt = 0;  // valid keys come first after the sort: binary search of the first kNone
                int lo = 0, hi = N;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] != kNone) lo = mid + 1; else hi = mid; }
                cnt = lo;
                st.fv_off[s_total[0]] = cnt;
                st.counts[1] = s_total[0]; st.counts[2] = cnt; st.counts[3] = n;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScoreThreads) k_bow_score_ptr(const uint32_t* __restrict__ qw, const double* __restrict__ qv, int nq,
                                                                 const uint32_t* const* __restrict__ cw, const double* const* __restrict__ cv,
                                                                 const int* __restrict__ cn, int ncand, double* __restrict__ scores) {
    const int cand = blockIdx.x;
    bow_score_one(qw, qv, nq, cw[cand], cv[cand], cn[cand], scores + cand);
}

}  // namespace corb

struct corb_bow_store {
    int device, cap;
    uint8_t* base = nullptr;
    BowStoreDev dev;
    int32_t* h_counts = nullptr;  // pinned [4]
    int n = 0, n_bow = 0, n_fv = 0, n_fv_idx = 0;
    cudaStream_t last_stream = nullptr;
    bool counts_pending = false;  // the fill is in flight: the counts are read (and the fill waited for) by the first consumer
};

extern "C" {

int corb_bow_store_create(int device, int capacity, corb_bow_store** out) {
    CORB_CHECK(out && capacity >= 1 && capacity <= kBowCap, CORB_ERR_INVALID, "capacity must be 1..%d features", kBowCap);
    CORB_CUDA(cudaSetDevice(device));
    corb_bow_store* s = new corb_bow_store;
    s->device = device;
    s->cap = capacity;
    Packer p;
    const size_t c = capacity;
    const size_t oD = p.take(c * 32), oA = p.take(c * 4), oW = p.take(c * 4), oWt = p.take(c * 8), oN = p.take(c * 4), oBW = p.take(c * 4),
                 oBV = p.take(c * 8), oFN = p.take(c * 4), oFO = p.take((c + 1) * 4), oFI = p.take(c * 4), oC = p.take(16);
    if (cudaMalloc((void**)&s->base, p.off) != cudaSuccess || cudaMallocHost((void**)&s->h_counts, 16) != cudaSuccess) {
        if (s->base) cudaFree(s->base);
        delete s;
        CORB_CHECK(false, CORB_ERR_CUDA, "allocating a BoW record of %d features failed", capacity);
    }
    uint8_t* b = s->base;
    s->dev.desc = b + oD; s->dev.angles = (float*)(b + oA); s->dev.word = (uint32_t*)(b + oW); s->dev.weight = (double*)(b + oWt);
    s->dev.node = (uint32_t*)(b + oN); s->dev.bow_words = (uint32_t*)(b + oBW); s->dev.bow_vals = (double*)(b + oBV);
    s->dev.fv_nodes = (uint32_t*)(b + oFN); s->dev.fv_off = (int32_t*)(b + oFO); s->dev.fv_idx = (uint32_t*)(b + oFI);
    s->dev.counts = (int32_t*)(b + oC);
    *out = s;
    return CORB_OK;
}

void corb_bow_store_destroy(corb_bow_store* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->base) cudaFree(s->base);
    if (s->h_counts) cudaFreeHost(s->h_counts);
    delete s;
}

int corb_bow_store_fill(corb_bow_store* s, corb_voc* v, const corb_keypoint* d_kps, const uint8_t* d_desc, int n, int levelsup,
                        void* stream) {
    CORB_CHECK(s && v && n >= 0 && n <= s->cap && (n == 0 || d_desc), CORB_ERR_INVALID, "bad argument (n = %d, capacity %d)", n,
               s ? s->cap : 0);
    CORB_CHECK(v->device == s->device, CORB_ERR_INVALID, "vocabulary and record live on different devices");
    CORB_CHECK(((uintptr_t)d_desc & 15) == 0, CORB_ERR_INVALID, "device descriptors must be 16-byte aligned");
    CORB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : v->stream;
    s->last_stream = st;
    s->n = n;
    if (n == 0) {
        s->counts_pending = false;
        s->n_bow = s->n_fv = s->n_fv_idx = 0;
        CORB_CUDA(cudaMemsetAsync(s->dev.counts, 0, 16, st));
        CORB_CUDA(cudaMemsetAsync(s->dev.fv_off, 0, 4, st));
        return CORB_OK;
    }
    k_bow_gather<<<(n + 255) / 256, 256, 0, st>>>(d_kps, d_desc, n, s->dev);
    k_voc_transform<<<(n * 16 + 255) / 256, 256, 0, st>>>(v->dev, (const uint4*)s->dev.desc, n, v->L - levelsup, s->dev.word, s->dev.weight,
                                                           s->dev.node);
    int N = 32;
    while (N < n) N <<= 1;
    const size_t smem = (size_t)N * 16 + (size_t)(N + 1) * 4;
    static std::atomic<unsigned long long> opt{0};
    if (smem > 48 * 1024 && !(opt.load() & (1ull << (s->device & 63)))) {
        CORB_CUDA(cudaFuncSetAttribute(k_bow_build, cudaFuncAttributeMaxDynamicSharedMemorySize, kBowCap * 16 + (kBowCap + 1) * 4));
        opt.fetch_or(1ull << (s->device & 63));
    }
    k_bow_build<<<2, 1024, smem, st>>>(s->dev, n);
    CORB_CUDA(cudaGetLastError());
    CORB_CUDA(cudaMemcpyAsync(s->h_counts, s->dev.counts, 16, cudaMemcpyDeviceToHost, st));
    s->counts_pending = true;  // no wait here: the record is built behind the caller; consumers settle it first
    return CORB_OK;
}

// waits for a record's fill and takes its counts over; every consumer calls it before it reads the record or launches on it
static int bow_store_settle(const corb_bow_store* cs) {
    corb_bow_store* s = const_cast<corb_bow_store*>(cs);
    if (!s->counts_pending) return CORB_OK;
    CORB_CUDA(cudaSetDevice(s->device));
    CORB_CUDA(cudaStreamSynchronize(s->last_stream));
    s->n_bow = s->h_counts[0]; s->n_fv = s->h_counts[1]; s->n_fv_idx = s->h_counts[2];
    s->counts_pending = false;
    return CORB_OK;
}

int corb_bow_store_features(const corb_bow_store* s) { return s ? s->n : 0; }

int corb_bow_store_sync(corb_bow_store* s) {
    CORB_CHECK(s, CORB_ERR_INVALID, "bad argument");
    return bow_store_settle(s);
}

int corb_bow_store_side(const corb_bow_store* s, const uint8_t* d_valid, corb_bow_side* side, int* n_bow) {
    CORB_CHECK(s && side, CORB_ERR_INVALID, "bad argument");
    {
        const int rc = bow_store_settle(s);
        if (rc != CORB_OK) return rc;
    }
    side->desc = s->dev.desc; side->n = s->n;
    side->fv_nodes = s->dev.fv_nodes; side->fv_off = s->dev.fv_off; side->fv_idx = s->dev.fv_idx; side->fv_n = s->n_fv;
    side->valid = d_valid; side->angles = s->dev.angles;
    if (n_bow) *n_bow = s->n_bow;
    return CORB_OK;
}

int corb_bow_store_download(const corb_bow_store* s, uint32_t* bow_words, double* bow_vals, int* n_bow, uint32_t* fv_nodes,
                            int32_t* fv_off, uint32_t* fv_idx, int* n_fv) {
    CORB_CHECK(s && n_bow && n_fv, CORB_ERR_INVALID, "bad argument");
    {
        const int rc = bow_store_settle(s);
        if (rc != CORB_OK) return rc;
    }
    CORB_CUDA(cudaSetDevice(s->device));
    *n_bow = s->n_bow; *n_fv = s->n_fv;
    if (bow_words && s->n_bow) CORB_CUDA(cudaMemcpy(bow_words, s->dev.bow_words, (size_t)s->n_bow * 4, cudaMemcpyDeviceToHost));
    if (bow_vals && s->n_bow) CORB_CUDA(cudaMemcpy(bow_vals, s->dev.bow_vals, (size_t)s->n_bow * 8, cudaMemcpyDeviceToHost));
    if (fv_nodes && s->n_fv) CORB_CUDA(cudaMemcpy(fv_nodes, s->dev.fv_nodes, (size_t)s->n_fv * 4, cudaMemcpyDeviceToHost));
    if (fv_off) CORB_CUDA(cudaMemcpy(fv_off, s->dev.fv_off, ((size_t)s->n_fv + 1) * 4, cudaMemcpyDeviceToHost));
    if (fv_idx && s->n_fv_idx) CORB_CUDA(cudaMemcpy(fv_idx, s->dev.fv_idx, (size_t)s->n_fv_idx * 4, cudaMemcpyDeviceToHost));
    return CORB_OK;
}

int corb_bow_score_stores(corb_voc* v, const corb_bow_store* query, int ncand, const corb_bow_store* const* cands, double* scores) {
    CORB_CHECK(v && query && ncand >= 0 && (ncand == 0 || (cands && scores)), CORB_ERR_INVALID, "bad argument");
    if (ncand == 0) return CORB_OK;
    CORB_CUDA(cudaSetDevice(v->device));
    Packer p;
    const size_t oW = p.take((size_t)ncand * 8), oV = p.take((size_t)ncand * 8), oN = p.take((size_t)ncand * 4), oS = p.take((size_t)ncand * 8);
    int rc = v->arena.reserve(p.off);
    if (rc != CORB_OK) return rc;
    uint8_t *h = v->arena.h, *d = v->arena.d;
    if ((rc = bow_store_settle(query)) != CORB_OK) return rc;
    for (int i = 0; i < ncand; i++) {
        CORB_CHECK(cands[i] && cands[i]->device == v->device, CORB_ERR_INVALID, "candidate %d is NULL or on another device", i);
        if ((rc = bow_store_settle(cands[i])) != CORB_OK) return rc;
        ((const uint32_t**)(h + oW))[i] = cands[i]->dev.bow_words;
        ((const double**)(h + oV))[i] = cands[i]->dev.bow_vals;
        ((int*)(h + oN))[i] = cands[i]->n_bow;
    }
    CORB_CUDA(cudaMemcpyAsync(d, h, oS, cudaMemcpyHostToDevice, v->stream));
    k_bow_score_ptr<<<ncand, kScoreThreads, 0, v->stream>>>(query->dev.bow_words, query->dev.bow_vals, query->n_bow,
                                                                     (const uint32_t* const*)(d + oW), (const double* const*)(d + oV),
                                                                     (const int*)(d + oN), ncand, (double*)(d + oS));
    CORB_CUDA(cudaGetLastError());
    CORB_CUDA(cudaMemcpyAsync(h + oS, d + oS, (size_t)ncand * 8, cudaMemcpyDeviceToHost, v->stream));
    CORB_CUDA(cudaStreamSynchronize(v->stream));
    memcpy(scores, h + oS, (size_t)ncand * 8);
    return CORB_OK;
}


/* SearchByBoW x3 between device-resident records: only the MapPoint liveness bytes go up and the match arrays come down. */
int corb_bow_match_stores(corb_matcher* m, int variant, int ncalls, const corb_bow_store* const* A, const uint8_t* const* validA,
                          const corb_bow_store* const* B, const uint8_t* const* validB, float nnratio, int check_ori,
                          int32_t* const* match, int32_t* nmatches) {
    CORB_CHECK(m && A && B && match && nmatches && ncalls >= 1, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(variant >= 0 && variant <= 2, CORB_ERR_INVALID, "unknown variant %d", variant);
    CORB_CUDA(cudaSetDevice(m->device));
    const bool kfkf = variant == 2;
    Packer p;
    std::vector<size_t> oVA(ncalls), oVB(ncalls), oM(ncalls);
    for (int i = 0; i < ncalls; i++) {
        CORB_CHECK(A[i] && B[i] && A[i]->device == m->device && B[i]->device == m->device, CORB_ERR_INVALID, "call %d: records missing or on another device", i);
        oVA[i] = p.take(validA && validA[i] ? A[i]->n : 0);
        oVB[i] = p.take(kfkf && validB && validB[i] ? B[i]->n : 0);
    }
    const size_t in_bytes = p.off;
    for (int i = 0; i < ncalls; i++) oM[i] = p.take((size_t)(kfkf ? A[i]->n : B[i]->n) * 4);
    const size_t oNm = p.take((size_t)ncalls * 4), oPtr = p.take((size_t)ncalls * 8);
    int rc = m->arena.reserve(p.off);
    if (rc != CORB_OK) return rc;
    uint8_t *h = m->arena.h, *d = m->arena.d;
    std::vector<corb_bow_side> sa(ncalls), sb(ncalls);
    std::vector<int32_t*> dm(ncalls);
    for (int i = 0; i < ncalls; i++) {
        const uint8_t *va = nullptr, *vb = nullptr;
        if (validA && validA[i]) { memcpy(h + oVA[i], validA[i], A[i]->n); va = d + oVA[i]; }
        if (kfkf && validB && validB[i]) { memcpy(h + oVB[i], validB[i], B[i]->n); vb = d + oVB[i]; }
        corb_bow_store_side(A[i], va, &sa[i], nullptr);
        corb_bow_store_side(B[i], vb, &sb[i], nullptr);
        dm[i] = (int32_t*)(d + oM[i]);
    }
    if (in_bytes) CORB_CUDA(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, m->stream));
    rc = corb_bow_match_batch_device(m, variant, ncalls, sa.data(), sb.data(), nnratio, check_ori, dm.data(), (int32_t*)(d + oNm));
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaMemcpyAsync(h + oM[0], d + oM[0], oPtr - oM[0], cudaMemcpyDeviceToHost, m->stream));
    CORB_CUDA(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < ncalls; i++) {
        memcpy(match[i], h + oM[i], (size_t)(kfkf ? A[i]->n : B[i]->n) * 4);
        nmatches[i] = ((int32_t*)(h + oNm))[i];
    }
    return CORB_OK;
}

}  // extern "C"
