/* CPU ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Plain C++ restatement of the reference's projection matchers:
 *   corbslam_client/src/ORBmatcher.cc:44-131   SearchByProjection(Frame&, const vector<MapPoint*>&, th)   [local map]
 *   corbslam_client/src/ORBmatcher.cc:133-139  RadiusByViewingCos
 *   corbslam_client/src/ORBmatcher.cc:1470-1614 SearchByProjection(Frame& Current, const Frame& Last, th, bMono)
 *   corbslam_client/src/Frame.cc:331-384       Frame::GetFeaturesInArea (64 x 48 grid, Frame.h:38-39)
 *   corbslam_client/src/ORBmatcher.cc:1746-1808 ComputeThreeMaxima, DescriptorDistance
 *
 * PARITY STATUS: PINNED against the reference's own ORBmatcher.cc / Frame.cc (GetFeaturesInArea on the real mGrid)
 * compiled unmodified into oracle/_ref/libref.so: match arrays equal on every test scene (tests/test_ref_cpu.py).
 * Also pinned by source constants (TH_HIGH 100, HISTO_LENGTH 30, FRAME_GRID 64 x 48, radii 2.5 / 4.0) and, for the one
 * piece of arithmetic delegated to OpenCV (cv::Mat products `Rcw*x3Dw+tcw`, `-Rcw.t()*tcw`, `Rlw*twc+tlw` =
 * cv::gemm on CV_32F: float dot product in the untransposed small-matrix case, double accumulation with a transposed
 * operand, alpha/beta applied in double; which gemm call a cv::MatExpr becomes follows matop.cpp), by golden vectors
 * generated with cv2.gemm 4.13.0 (tests/golden/opencv_primitives.npz, tools/gen_golden.py).
 * Built with -ffp-contract=off: `fx*xc*invzc+cx` is two float multiplications and one float addition.
 *
 * The MapPoint / Frame pointer graph is flattened by the caller exactly like the C-ABI shim does (include/corb_b200.h,
 * corb_frame_view): `taken[i]` = the frame feature holds a MapPoint with Observations() > 0, `blocks[q]` = the query's
 * MapPoint has Observations() > 0 (so the feature it is assigned to is skipped by later queries).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

const int TH_HIGH = 100, HISTO_LENGTH = 30; /* ORBmatcher.cc:37-39 */
const int GRID_COLS = 64, GRID_ROWS = 48;   /* Frame.h:38-39 */

inline int hamming256(const uint8_t* a, const uint8_t* b) { /* ORBmatcher.cc:1792-1808 */
    const int32_t* pa = (const int32_t*)a;
    const int32_t* pb = (const int32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

} // namespace

extern "C" {

struct oracle_frame_view { /* same layout as corb_frame_view (include/corb_b200.h) */
    int32_t n;
    const float* x; const float* y; const int32_t* octave; const float* angle;
    const uint8_t* desc;
    const float* u_right;
    const uint8_t* taken;
    const int32_t* grid_off;
    const int32_t* grid_idx;
    float min_x, min_y, max_x, max_y, grid_w_inv, grid_h_inv;
    const float* scale_factors; int32_t n_levels;
    float fx, fy, cx, cy, mbf, mb;
    float Tcw[12];
};

/* cv::gemm model for a 3x3 (optionally transposed) times 3x1 product on CV_32F, pinned against cv2.gemm 4.13.0:
 *   untransposed (flags == 0): OpenCV's small-matrix special case - the dot product is evaluated in float, left to
 *     right, then d = (float)(s * alpha + c * beta) with alpha, beta double;
 *   transposed operand: the generic kernel GEMMSingleMul<float,double> - double accumulation, one rounding.
 * M row-major 3x4 [R | t]. */
void oracle_gemm3(const float* M, int transpose, double alpha, const float* v, double beta, const float* c, float* out) {
    for (int r = 0; r < 3; r++) {
        double s;
        if (!transpose) {
            const float p0 = M[r * 4] * v[0], p1 = M[r * 4 + 1] * v[1], p2 = M[r * 4 + 2] * v[2];
            const float t = (p0 + p1) + p2;
            s = (double)t * alpha;
        } else {
            s = 0;
            for (int k = 0; k < 3; k++) s += (double)M[k * 4 + r] * (double)v[k];
            s *= alpha;
        }
        if (c) s += beta * (double)c[r];
        out[r] = (float)s;
    }
}

/* Frame::GetFeaturesInArea (Frame.cc:331-384); appends indices in the reference's order */
static void features_in_area(const oracle_frame_view* F, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) {
    out.clear();
    const int nMinCellX = std::max(0, (int)floorf((x - F->min_x - r) * F->grid_w_inv));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)ceilf((x - F->min_x + r) * F->grid_w_inv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)floorf((y - F->min_y - r) * F->grid_h_inv));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)ceilf((y - F->min_y + r) * F->grid_h_inv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * GRID_ROWS + iy;
            for (int e = F->grid_off[c]; e < F->grid_off[c + 1]; e++) {
                const int i = F->grid_idx[e];
                if (bCheckLevels) {
                    if (F->octave[i] < minLevel) continue;
                    if (maxLevel >= 0 && F->octave[i] > maxLevel) continue;
                }
                const float distx = F->x[i] - x, disty = F->y[i] - y;
                if (fabsf(distx) < r && fabsf(disty) < r) out.push_back(i);
            }
        }
}

static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) { /* :1746-1787 */
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

/* SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), ORBmatcher.cc:1470-1614.
 * match[i2] (size cur->n, preset by the caller to -1) = index of the last-frame feature whose MapPoint the current
 * feature i2 received. Returns nmatches. */
int oracle_search_by_projection_last(const oracle_frame_view* cur, int n_last, const uint8_t* last_valid, const uint8_t* last_blocks,
                                     const float* last_xyz, const uint8_t* last_mp_desc, const int32_t* last_octave,
                                     const float* last_angle, const float* Tlw, float th, int mono, int check_ori, int32_t* match) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<uint8_t> taken(cur->n, 0);
    if (cur->taken) memcpy(taken.data(), cur->taken, cur->n);
    float twc[3], tlc[3];
    const float tcw[3] = {cur->Tcw[3], cur->Tcw[7], cur->Tcw[11]}, tlw[3] = {Tlw[3], Tlw[7], Tlw[11]};
    /* twc = -Rcw.t()*tcw (:1483). cv::MatExpr: unary minus on a transpose node materialises the transpose (matop.cpp,
     * MatOp::subtract(Scalar, expr)), the product then is gemm(Rt, tcw, alpha = -1, flags = 0) = the small-matrix float path. */
    const float Rt[12] = {cur->Tcw[0], cur->Tcw[4], cur->Tcw[8], 0.f, cur->Tcw[1], cur->Tcw[5], cur->Tcw[9], 0.f,
                          cur->Tcw[2], cur->Tcw[6], cur->Tcw[10], 0.f};
    oracle_gemm3(Rt, 0, -1.0, tcw, 0.0, nullptr, twc);
    oracle_gemm3(Tlw, 0, 1.0, twc, 1.0, tlw, tlc);             /* tlc = Rlw*twc+tlw    (:1488) */
    const bool bForward = tlc[2] > cur->mb && !mono, bBackward = -tlc[2] > cur->mb && !mono;
    std::vector<int> vIndices2;
    for (int i = 0; i < n_last; i++) {
        if (!last_valid[i]) continue;
        float x3Dc[3];
        oracle_gemm3(cur->Tcw, 0, 1.0, last_xyz + 3 * i, 1.0, tcw, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / (double)x3Dc[2]);
        if (invzc < 0) continue;
        const float u = cur->fx * xc * invzc + cur->cx;
        const float v = cur->fy * yc * invzc + cur->cy;
        if (u < cur->min_x || u > cur->max_x) continue;
        if (v < cur->min_y || v > cur->max_y) continue;
        const int nLastOctave = last_octave[i];
        const float radius = th * cur->scale_factors[nLastOctave];
        if (bForward) features_in_area(cur, u, v, radius, nLastOctave, -1, vIndices2);
        else if (bBackward) features_in_area(cur, u, v, radius, 0, nLastOctave, vIndices2);
        else features_in_area(cur, u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = last_mp_desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (taken[i2]) continue;
            if (cur->u_right[i2] > 0) {
                const float ur = u - cur->mbf * invzc;
                const float er = fabsf(ur - cur->u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = hamming256(dMP, cur->desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            match[bestIdx2] = i;
            if (!last_blocks || last_blocks[i]) taken[bestIdx2] = 1;
            nmatches++;
            if (check_ori) {
                float rot = last_angle[i] - cur->angle[bestIdx2];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { match[idx] = -1; nmatches--; }
    }
    return nmatches;
}

/* SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th), ORBmatcher.cc:44-131.
 * proj = (mTrackProjX, mTrackProjY, mTrackProjXR) per map point. match[idx] (size cur->n, preset to -1) = index of the
 * map point assigned to frame feature idx. Returns nmatches. */
int oracle_search_by_projection_map(const oracle_frame_view* F, int n_mp, const uint8_t* in_view, const uint8_t* blocks, const float* proj,
                                    const int32_t* level, const float* view_cos, const uint8_t* mp_desc, float th, float nnratio,
                                    int32_t* match) {
    int nmatches = 0;
    const bool bFactor = th != 1.0f;
    std::vector<uint8_t> taken(F->n, 0);
    if (F->taken) memcpy(taken.data(), F->taken, F->n);
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < n_mp; iMP++) {
        if (!in_view[iMP]) continue;
        const int nPredictedLevel = level[iMP];
        float r = view_cos[iMP] > 0.998f ? 2.5f : 4.0f; /* RadiusByViewingCos :133-139 */
        if (bFactor) r *= th;
        const float win = r * F->scale_factors[nPredictedLevel];
        features_in_area(F, proj[3 * iMP], proj[3 * iMP + 1], win, nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* MPdescriptor = mp_desc + 32 * (size_t)iMP;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (taken[idx]) continue;
            if (F->u_right[idx] > 0) {
                const float er = fabsf(proj[3 * iMP + 2] - F->u_right[idx]);
                if (er > win) continue;
            }
            const int dist = hamming256(MPdescriptor, F->desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F->octave[idx]; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F->octave[idx]; bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            match[bestIdx] = iMP;
            if (!blocks || blocks[iMP]) taken[bestIdx] = 1;
            nmatches++;
        }
    }
    return nmatches;
}

} // extern "C"
