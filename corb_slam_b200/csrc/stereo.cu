// Frame::ComputeStereoMatches (corbslam_client/src/Frame.cc:470-644) on the device-resident results of the left and
// right extractor handles: row-band Hamming search, 11x11 SAD refinement over +-5 px on the keypoint's pyramid level,
// parabola fit, median-based outlier rejection. With this on the GPU the pyramids never cross PCIe;
// only mvuRight / mvDepth (2 x N floats) go back to the host. Bit-exact against oracle_stereo_matches.
#include <limits.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "orb_kernels.cuh"

namespace corb {

constexpr int kThHigh = 100, kThLowS = 50;

constexpr int kBandShift = 3;  // 8 image rows per band

// one thread per right keypoint: row range / octave / x record + membership in the bands its rows touch
__global__ void __launch_bounds__(256) k_stereo_bands(OrbGeom g, StereoArgs a) {
    const int jr = blockIdx.x * 256 + threadIdx.x;
    if (jr >= *a.nr) return;
    const corb_keypoint kpR = a.kr[jr];
    const float r = __fmul_rn(2.0f, a.scale[kpR.octave]);
    const int lo = (int)floorf(__fsub_rn(kpR.y, r)), hi = (int)ceilf(__fadd_rn(kpR.y, r));
    a.rinfo[jr] = make_int4(lo, hi, kpR.octave, __float_as_int(kpR.x));
    const int b0 = max(lo, 0) >> kBandShift, b1 = min(min(hi, a.n_rows - 1) >> kBandShift, a.n_bands - 1);
    for (int b = b0; b <= b1; b++) {
        const int pos = atomicAdd(&a.band_cnt[b], 1);
        if (pos < a.band_cap) a.band_list[(size_t)b * a.band_cap + pos] = jr;
    }
}

// one warp per left keypoint
__global__ void __launch_bounds__(256) k_stereo_match(OrbGeom g, StereoArgs a) {
    const int iL = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nl = *a.nl, nr = *a.nr;
    corb_keypoint kpL;
    kpL.x = kpL.y = 0.f; kpL.octave = 0;
    if (iL < nl) kpL = a.kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    const float minZ = a.mb, minD = 0.0f, maxD = __fdiv_rn(a.mbf, minZ);
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
    float out_u = -1.0f, out_d = -1.0f;
    int out_sad = -1;
    int bestDist = kThHigh, bestIdx = INT_MAX;
    // The left keypoint looks only at the right keypoints of its 8-row band (a twentieth of them on a 375-row image);
    // the exact row-table membership of the reference (:487-497) is two compares on the keypoint's record. The reference
    // scans candidates in ascending right index and keeps the first minimum: here the order inside a band is arbitrary,
    // so ties go to the lower index explicitly.
    const bool searching = iL < nl && !(maxU < 0) && row >= 0 && row < a.n_rows;
    uint4 l0 = make_uint4(0, 0, 0, 0), l1 = l0;
    if (searching) {
        l0 = a.dl[2 * iL]; l1 = a.dl[2 * iL + 1];
        const int band = min(row >> kBandShift, a.n_bands - 1);
        const int cnt = min(a.band_cnt[band], a.band_cap);
        const int* list = a.band_list + (size_t)band * a.band_cap;
        for (int j = lane; j < cnt; j += 32) {
            const int iR = list[j];
            const int4 t = a.rinfo[iR];
            if (row < t.x || row > t.y) continue;                          // vRowIndices[(int)vL] membership (:487-497)
            if (t.z < levelL - 1 || t.z > levelL + 1) continue;
            const float uR = __int_as_float(t.w);
            if (!(uR >= minU && uR <= maxU)) continue;
            const uint4 r0 = a.dr[2 * iR], r1 = a.dr[2 * iR + 1];
            const int d = __popc(l0.x ^ r0.x) + __popc(l0.y ^ r0.y) + __popc(l0.z ^ r0.z) + __popc(l0.w ^ r0.w) + __popc(l1.x ^ r1.x) +
                          __popc(l1.y ^ r1.y) + __popc(l1.z ^ r1.z) + __popc(l1.w ^ r1.w);
            if (d < bestDist || (d == bestDist && iR < bestIdx)) { bestDist = d; bestIdx = iR; }
        }
    }
    (void)nr;
    if (iL >= nl) {
        if (lane == 0 && iL < g.kp_cap) { a.u_right[iL] = -1.0f; a.depth[iL] = -1.0f; a.best_dist[iL] = -1; }
        return;
    }
    if (searching) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int od = __shfl_xor_sync(0xffffffffu, bestDist, o), oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
            if (od < bestDist || (od == bestDist && oi < bestIdx)) { bestDist = od; bestIdx = oi; }
        }
        if (bestDist < (kThHigh + kThLowS) / 2) {
            // ---- sub-pixel refinement on level kpL.octave of both pyramids (:551-626)
            const float uR0 = a.kr[bestIdx].x;
            const float sf = a.inv_scale[levelL];
            const float scaleduL = roundf(__fmul_rn(uL, sf)), scaledvL = roundf(__fmul_rn(vL, sf));
            const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
            const LevelGeom L = g.lv[levelL];
            const float iniu = scaleduR0, endu = __fadd_rn(scaleduR0, 11.0f);   // scaleduR0 + L - w, scaleduR0 + L + w + 1
            if (!(iniu < 0 || endu >= (float)L.w)) {
                const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
                const uint8_t* pl = a.pyr_l + L.img_off + (size_t)cv * L.pitch + cu;
                const uint8_t* pr = a.pyr_r + L.img_off + (size_t)cv * L.pitch + cr;
                const int cL = pl[0];
                int av[4], off[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int p = lane + 32 * t;
                    const int dy = p / 11 - 5, dx = p - (p / 11) * 11 - 5;
                    off[t] = p < 121 ? dy * L.pitch + dx : INT_MIN;
                    av[t] = p < 121 ? (int)pl[off[t]] - cL : 0;
                }
                int sadBest = INT_MAX, bestinc = 0;
                float d_prev = 0.f, d_best = 0.f, d_next = 0.f, d_last = 0.f;
                bool want_next = false;
#pragma unroll 1
                for (int inc = -5; inc <= 5; inc++) {
                    const int cR = pr[inc];
                    int acc = 0;
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (off[t] != INT_MIN) acc += abs(av[t] - ((int)pr[off[t] + inc] - cR));
#pragma unroll
                    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    const float dist = (float)acc;   // exact: the SAD of integer-valued floats stays below 2^24
                    if (want_next) { d_next = dist; want_next = false; }
                    if (dist < (float)sadBest) {
                        sadBest = (int)dist; bestinc = inc;
                        d_prev = d_last; d_best = dist; want_next = true;
                    }
                    d_last = dist;
                }
                if (bestinc != -5 && bestinc != 5) {
                    const float deltaR = __fdiv_rn(__fsub_rn(d_prev, d_next),
                                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d_prev, d_next), __fmul_rn(2.0f, d_best))));
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = __fmul_rn(a.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= minD && disparity < maxD) {
                            if (disparity <= 0) {
                                disparity = (float)0.01;
                                bestuR = (float)((double)uL - 0.01);
                            }
                            out_d = __fdiv_rn(a.mbf, disparity);
                            out_u = bestuR;
                            out_sad = sadBest;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) { a.u_right[iL] = out_u; a.depth[iL] = out_d; a.best_dist[iL] = out_sad; }
}

// median of the accepted SADs (element size/2 of the sorted list) and rejection of matches >= 1.5 * 1.4 * median (:630-643).
// The SAD of an 11x11 u8 patch is < 2^16, so the k-th smallest value is found by a two-pass radix select on shared-
// memory histograms (high byte, then low byte) instead of sorting.
__global__ void __launch_bounds__(1024) k_stereo_outliers(OrbGeom g, StereoArgs a) {
    __shared__ int hist[256];
    __shared__ int s_sel[3];  // [0] selected bin, [1] remaining rank, [2] count
    const int tid = threadIdx.x;
    const int nl = *a.nl;
    for (int i = tid; i < a.n_bands; i += 1024) a.band_cnt[i] = 0;  // k_stereo_match is done with the bands: ready for the next frame
    int hi_bin = 0, k = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < nl; i += 1024) {
            const int d = a.best_dist[i];
            if (d < 0) continue;
            if (pass == 0) atomicAdd(&hist[(d >> 8) & 255], 1);
            else if (((d >> 8) & 255) == hi_bin) atomicAdd(&hist[d & 255], 1);
        }
        __syncthreads();
        if (tid == 0) {
            if (pass == 0) {
                int n = 0;
                for (int b = 0; b < 256; b++) n += hist[b];
                s_sel[2] = n;
                k = n / 2;
            }
            int cum = 0, sel = 255;
            for (int b = 0; b < 256; b++) {
                if (k < cum + hist[b]) { sel = b; break; }
                cum += hist[b];
            }
            s_sel[0] = sel;
            s_sel[1] = k - cum;
        }
        __syncthreads();
        if (s_sel[2] == 0) return;  // no match at all (the reference would read an empty vector here)
        if (pass == 0) hi_bin = s_sel[0];
        k = s_sel[1];
        __syncthreads();
    }
    const float median = (float)((hi_bin << 8) | s_sel[0]);
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = tid; i < nl; i += 1024) {
        const int di = a.best_dist[i];
        if (di >= 0 && !((float)di < thDist)) { a.u_right[i] = -1.0f; a.depth[i] = -1.0f; }
    }
}

void launch_stereo(const OrbGeom& g, const StereoArgs& a, cudaStream_t s) {
    k_stereo_bands<<<(g.kp_cap + 255) / 256, 256, 0, s>>>(g, a);
    k_stereo_match<<<(g.kp_cap * 32 + 255) / 256, 256, 0, s>>>(g, a);
    k_stereo_outliers<<<1, 1024, 0, s>>>(g, a);
}

}  // namespace corb
