// EPnP (PnPsolver.cc:420-962) for one thread: everything one RANSAC hypothesis or one Refine() needs, as functions that
// run unchanged on the device (pnp.cu, compiled with -fmad=false) and - for the CPU debug harness in tests/ only - on
// the host. Only + - * / sqrt on doubles, every sum in the reference's order, so the result of a thread is a pure
// function of its inputs and of nothing else (no atomics, no shuffles, no contraction).
//
// OpenCV's cvSVD / cvSolve(CV_SVD) / cvInvert(CV_SVD) / cvMulTransposed (un-vendored dependency, call sites
// PnPsolver.cc:446-447,468,518-519,626,718,749,782) are the one-sided Jacobi SVD of modules/core/src/lapack.cpp
// (JacobiSVDImpl_, SVBkSbImpl_) with hypot(p, beta) written as sqrt(p*p + beta*beta) - see DESIGN.md §4.4.
//
// The correspondences of a set are never copied: a set is either 4 indices (a minimal set) or a bit mask over the
// problem's points (the inliers Refine() works on), walked in ascending order = the order of the reference's vIndices.
// alphas and pcs are recomputed per pass from the 3x3 inverse / the control points instead of being stored per point.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#ifndef CORB_HD
#define CORB_HD __host__ __device__
#endif

namespace corb {
namespace pnp {

// Per-thread workspace (doubles), element i at p[i * stride]: stride 32 in shared memory makes a warp's accesses to
// the same element conflict-free.
constexpr int WS_A = 0;     // 144: M^T M, then the rows of Ut
constexpr int WS_W = 144;   // 12 singular values
constexpr int WS_S = 156;   // small SVDs: At (<= 30) | Vt (<= 25) | W (<= 5)
constexpr int WS_SV = WS_S + 30;
constexpr int WS_SW = WS_SV + 25;
constexpr int WS_DOUBLES = 216;

struct Ws {
    double* p;
    int stride;
    CORB_HD double& operator[](int i) const { return p[(size_t)i * stride]; }
    CORB_HD Ws at(int off) const { return Ws{p + (size_t)off * stride, stride}; }
};

struct PtSet {
    const float* p3d;  // the problem's MapPoint positions (x, y, z)
    const float* p2d;  // the problem's keypoints (u, v)
    int n;             // correspondences in the set
    int list[4];       // minimal set (mask == nullptr)
    const uint32_t* mask;
    int n_words;
};

struct PtIter {
    int k, wi;
    uint32_t w;
};

CORB_HD inline int ctz32(uint32_t w) {
#ifdef __CUDA_ARCH__
    return __ffs((int)w) - 1;
#else
    return __builtin_ctz(w);
#endif
}
CORB_HD inline void it_begin(PtIter& it) { it.k = 0; it.wi = -1; it.w = 0; }
CORB_HD inline bool it_next(const PtSet& s, PtIter& it, int& idx) {
    if (!s.mask) {
        if (it.k >= s.n) return false;
        idx = it.k == 0 ? s.list[0] : it.k == 1 ? s.list[1] : it.k == 2 ? s.list[2] : s.list[3];
        it.k++;
        return true;
    }
    while (it.w == 0) {
        if (++it.wi >= s.n_words) return false;
        it.w = s.mask[it.wi];
    }
    idx = it.wi * 32 + ctz32(it.w);
    it.w &= it.w - 1;
    return true;
}
CORB_HD inline void load_pw(const PtSet& s, int idx, double pw[3]) {  // add_correspondence: floats widened, :407-418
    pw[0] = (double)s.p3d[3 * idx]; pw[1] = (double)s.p3d[3 * idx + 1]; pw[2] = (double)s.p3d[3 * idx + 2];
}

// acc[a] += term_a(point) for every point of the set, in set order. WARP (device only, mask sets only, all 32 lanes of
// the warp hold identical state and call together): lane l evaluates the addends of bit l of each mask word, then the
// additions are performed in ascending bit order on every lane (shuffle broadcast), so each lane ends with the same
// accumulators a single thread would have produced - same addends, same order, same bits.
// The minimal set of one RANSAC iteration (PnPsolver.cc:228-242): vAvailableIndices = mvAllIndices (the identity list of
// n entries); four times: idx = avail[randi]; avail[randi] = avail.back(); avail.pop_back(). At most four positions of the
// identity list are ever modified, so the list is never materialised: (mp, mv) record the modified positions, the latest
// modification of a position wins.
CORB_HD inline void resolve_draws(int n, const int r[4], int list[4]) {
    int mp[4], mv[4], sz = n;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        int idx = r[k];
        const int last = sz - 1;
        int lastval = last;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int j = 0; j < k; j++) {  // ascending j: the latest modification wins
            if (mp[j] == r[k]) idx = mv[j];
            if (mp[j] == last) lastval = mv[j];
        }
        list[k] = idx;
        mp[k] = r[k];
        mv[k] = lastval;
        sz--;
    }
}

constexpr int TBUF_ROW = 33;                 // transpose buffer: [accumulator][lane], padded so that column reads spread over the banks
constexpr int TBUF_DOUBLES = 9 * TBUF_ROW;  // up to 9 accumulators per pass

template <bool WARP, int NA, class F>
CORB_HD inline void ordered_sum(const PtSet& s, double* acc, F&& term, double* tbuf = nullptr) {
#ifdef __CUDA_ARCH__
    if (WARP && s.mask && tbuf) {
        // addends of a mask word go through shared memory: lane a < NA then owns accumulator a and adds its column in
        // ascending bit order (one load + one add per addend instead of a shuffle broadcast of every addend to every lane);
        // the finished accumulators are broadcast once
        const int lane = threadIdx.x & 31;
        double own = lane < NA ? acc[lane < NA ? lane : 0] : 0.0;
        for (int w = 0; w < s.n_words; w++) {
            const uint32_t bits = s.mask[w];
            if (!bits) continue;
            if ((bits >> lane) & 1u) {
                double t[NA];
                term(w * 32 + lane, t);
#pragma unroll
                for (int a = 0; a < NA; a++) tbuf[a * TBUF_ROW + lane] = t[a];
            }
            __syncwarp();
            if (lane < NA)
                for (uint32_t b = bits; b; b &= b - 1) own += tbuf[lane * TBUF_ROW + __ffs((int)b) - 1];
            __syncwarp();
        }
#pragma unroll
        for (int a = 0; a < NA; a++) acc[a] = __shfl_sync(0xffffffffu, own, a);
        return;
    }
    if (WARP && s.mask) {
        const int lane = threadIdx.x & 31;
        for (int w = 0; w < s.n_words; w++) {
            const uint32_t bits = s.mask[w];
            if (!bits) continue;
            double t[NA];
#pragma unroll
            for (int a = 0; a < NA; a++) t[a] = 0;
            if ((bits >> lane) & 1u) term(w * 32 + lane, t);
            for (uint32_t b = bits; b; b &= b - 1) {
                const int k = __ffs((int)b) - 1;
#pragma unroll
                for (int a = 0; a < NA; a++) acc[a] += __shfl_sync(0xffffffffu, t[a], k);
            }
        }
        return;
    }
#endif
    PtIter it;
    int idx;
    double t[NA];
    for (it_begin(it); it_next(s, it, idx);) {
        term(idx, t);
        for (int a = 0; a < NA; a++) acc[a] += t[a];
    }
}

// JacobiSVDImpl_<double>: At = A^T (n rows of length m, m >= n). On return row i of At = left singular vector i, W
// descending, row i of Vt = right singular vector i (only when with_v).
// The sizes are template parameters so that the loops over a row unroll: a thread then has all loads of a row pair in
// flight at once instead of one dependent load-multiply-add per trip (3x on the whole EPnP).
template <int m, int n, bool with_v>
CORB_HD inline void jacobi_svd(Ws At, Ws W, Ws Vt) {
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    const int max_iter = m > 30 ? m : 30;
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { const double t = At[i * m + k]; sd += t * t; }
        W[i] = sd;
        if (with_v) {
            for (int k = 0; k < n; k++) Vt[i * n + k] = 0;
            Vt[i * n + i] = 1;
        }
    }
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                const Ws Ai = At.at(i * m), Aj = At.at(j * m);
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = sqrt(p * p + beta * beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
#pragma unroll
                for (int k = 0; k < m; k++) {
                    const double x = Ai[k], y = Aj[k];
                    const double t0 = c * x + s * y;
                    const double t1 = -s * x + c * y;
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                if (with_v) {
                    const Ws Vi = Vt.at(i * n), Vj = Vt.at(j * n);
#pragma unroll
                    for (int k = 0; k < n; k++) {
                        const double x = Vi[k], y = Vj[k];
                        const double t0 = c * x + s * y;
                        const double t1 = -s * x + c * y;
                        Vi[k] = t0; Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { const double t = At[i * m + k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {
        int j = i;
        for (int k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double tmp = W[i]; W[i] = W[j]; W[j] = tmp;
            for (int k = 0; k < m; k++) { tmp = At[i * m + k]; At[i * m + k] = At[j * m + k]; At[j * m + k] = tmp; }
            if (with_v)
                for (int k = 0; k < n; k++) { tmp = Vt[i * n + k]; Vt[i * n + k] = Vt[j * n + k]; Vt[j * n + k] = tmp; }
        }
    }
    for (int i = 0; i < n; i++) {
        const double sd = W[i];
        const double s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < m; k++) At[i * m + k] *= s;
    }
}

// Team-cooperative form of jacobi_svd<12, 12, false> (device only). The pair visits of a sweep, (0,1), (0,2), ... (10,11),
// touch only rows i and j and W[i], W[j]; visits on disjoint rows do not interact at all, so any schedule that keeps the
// lexicographic order between visits that SHARE a row produces the sequential result bit for bit. Visit (i, j) runs at
// step i + j: two visits of one step have different i and different j (and i = j' would need j < i), so they are
// disjoint, and for visits sharing a row the step order equals the lexicographic order. 21 steps of up to 6 concurrent
// visits replace 66 sequential ones. S is the team's matrix in shared memory (row stride 13: the rows six lanes touch in
// one step fall into different banks), SW its 12 squared norms / singular values, perm 12 ints. All lanes of the WARP
// must call (full-mask barriers); `tl` is the lane's index inside its team of `team` lanes (8 or 32), teams converge
// independently. On return rows 8..11 of S are the left singular vectors of the four smallest singular values (all
// EPnP uses, compute_L_6x10 / compute_ccs), the other rows are left unnormalised.
struct Coop {
    double* S;   // nullptr: single-thread Jacobi
    double* SW;
    int* perm;
    int tl, team;
    double* tbuf;    // per-warp transpose buffer of the ordered sums (TBUF_DOUBLES), or nullptr
    bool lead_only;  // the team shares ONE workspace and only its lane 0 runs the scalar phases (the others only take part in
                     // the wavefront); false: every lane replicates the scalar phases on its own workspace
};
constexpr int COOP_ROW = 13;
constexpr int COOP_DOUBLES = 12 * COOP_ROW + 12 + 6;  // S | SW | perm (12 ints)

#ifdef __CUDACC__
__device__ inline void jacobi12_wavefront(const Coop& co) {
    const double eps = DBL_EPSILON * 10;
    double* S = co.S;
    double* W = co.SW;
    const int tl = co.tl;
    const unsigned lane = threadIdx.x & 31;
    const unsigned team_mask = co.team == 32 ? 0xffffffffu : (((1u << co.team) - 1u) << (lane & ~(unsigned)(co.team - 1)));
    for (int r = tl; r < 12; r += co.team) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) { const double t = S[r * COOP_ROW + k]; sd += t * t; }
        W[r] = sd;
    }
    __syncwarp();
    bool done = false;
    for (int iter = 0; iter < 30; iter++) {  // max_iter = max(m, 30)
        bool changed = false;
        for (int t = 1; t <= 21; t++) {
            const int i = (t > 11 ? t - 11 : 0) + tl, j = t - i;
            if (!done && i < j) {
                double* Ai = S + i * COOP_ROW;
                double* Aj = S + j * COOP_ROW;
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < 12; k++) p += Ai[k] * Aj[k];
                if (!(fabs(p) <= eps * sqrt(a * b))) {
                    p *= 2;
                    const double beta = a - b, gamma = sqrt(p * p + beta * beta);
                    double c, sn;
                    if (beta < 0) {
                        const double delta = (gamma - beta) * 0.5;
                        sn = sqrt(delta / gamma);
                        c = p / (gamma * sn * 2);
                    } else {
                        c = sqrt((gamma + beta) / (gamma * 2));
                        sn = p / (gamma * c * 2);
                    }
                    a = b = 0;
#pragma unroll
                    for (int k = 0; k < 12; k++) {
                        const double x = Ai[k], y = Aj[k];
                        const double t0 = c * x + sn * y;
                        const double t1 = -sn * x + c * y;
                        Ai[k] = t0; Aj[k] = t1;
                        a += t0 * t0; b += t1 * t1;
                    }
                    W[i] = a; W[j] = b;
                    changed = true;
                }
            }
            __syncwarp();
        }
        if (!(__ballot_sync(0xffffffffu, changed) & team_mask)) done = true;  // if (!changed) break;
        if (__all_sync(0xffffffffu, done)) break;
    }
    for (int r = tl; r < 12; r += co.team) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) { const double t = S[r * COOP_ROW + k]; sd += t * t; }
        W[r] = sqrt(sd);
    }
    __syncwarp();
    if (tl == 0) {  // the descending selection sort (first maximum wins, swap), on a permutation instead of on the rows
        int* perm = co.perm;
        for (int i = 0; i < 12; i++) perm[i] = i;
        for (int i = 0; i < 11; i++) {
            int j = i;
            for (int k = i + 1; k < 12; k++)
                if (W[j] < W[k]) j = k;
            if (i != j) {
                const double tw = W[i]; W[i] = W[j]; W[j] = tw;
                const int tp = perm[i]; perm[i] = perm[j]; perm[j] = tp;
            }
        }
    }
    __syncwarp();
}
#endif

// SVD of a small row-major m x n matrix held in registers / local memory, through the S area of the workspace.
template <int m, int n, bool with_v>
CORB_HD inline void svd_small(const double* A, Ws ws) {
    const Ws At = ws.at(WS_S);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < m; k++) At[i * m + k] = A[k * n + i];
    jacobi_svd<m, n, with_v>(At, ws.at(WS_SW), ws.at(WS_SV));
}

// cvSolve(CV_SVD), one right-hand side (SVBkSbImpl_): x = sum_i v_i ((u_i . b) (1 / w_i)) over w_i > 2 eps sum(w)
template <int m, int n>
CORB_HD inline void svd_solve(const double* A, const double* b, double* x, Ws ws) {
    svd_small<m, n, true>(A, ws);
    const Ws Ut = ws.at(WS_S), Vt = ws.at(WS_SV), W = ws.at(WS_SW);
    double threshold = 0;
    for (int i = 0; i < n; i++) { x[i] = 0; threshold += W[i]; }
    threshold *= DBL_EPSILON * 2;
    for (int i = 0; i < n; i++) {
        double wi = W[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < m; j++) s += Ut[i * m + j] * b[j];
        s *= wi;
        for (int j = 0; j < n; j++) x[j] = x[j] + s * Vt[i * n + j];
    }
}

// cvInvert(CV_SVD) of a 3 x 3 matrix
CORB_HD inline void svd_invert3(const double* A, double* inv, Ws ws) {
    svd_small<3, 3, true>(A, ws);
    const Ws Ut = ws.at(WS_S), Vt = ws.at(WS_SV), W = ws.at(WS_SW);
    double threshold = 0;
    for (int i = 0; i < 3; i++) threshold += W[i];
    threshold *= DBL_EPSILON * 2;
    for (int i = 0; i < 9; i++) inv[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = W[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        const double b0 = Ut[i * 3] * wi, b1 = Ut[i * 3 + 1] * wi, b2 = Ut[i * 3 + 2] * wi;
        for (int j = 0; j < 3; j++) {
            const double v = Vt[i * 3 + j];
            inv[j * 3] = inv[j * 3] + v * b0;
            inv[j * 3 + 1] = inv[j * 3 + 1] + v * b1;
            inv[j * 3 + 2] = inv[j * 3 + 2] + v * b2;
        }
    }
}

CORB_HD inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
CORB_HD inline double dist2(const double* p1, const double* p2) {
    return (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) + (p1[2] - p2[2]) * (p1[2] - p2[2]);
}

struct Epnp {
    double fu, fv, uc, vc;
    double cws[4][3], ccs[4][3], ci[9];
    Coop coop = {nullptr, nullptr, nullptr, 0, 32, nullptr, false};

    // SVD of the symmetric 12 x 12 in ws[WS_A..]: afterwards rows 8..11 hold the left singular vectors EPnP uses. With a
    // team (coop.S != nullptr; every lane of the team holds the same matrix in its own workspace) the sweeps run as a
    // wavefront on the team's shared copy and each lane takes rows 8..11 back, scaled by 1 / w like JacobiSVDImpl_ does.
    CORB_HD void svd12(Ws ws) {
        const Ws A = ws.at(WS_A);
#ifdef __CUDA_ARCH__
        if (coop.S) {
            for (int r = coop.tl; r < 12; r += coop.team)
                for (int k = 0; k < 12; k++) coop.S[r * COOP_ROW + k] = A[r * 12 + k];
            __syncwarp();
            jacobi12_wavefront(coop);
            if (!coop.lead_only || coop.tl == 0)
                for (int r = 8; r < 12; r++) {
                    const double sd = coop.SW[r];
                    const double sc = sd > DBL_MIN ? 1 / sd : 0.;
                    const double* row = coop.S + coop.perm[r] * COOP_ROW;
                    for (int k = 0; k < 12; k++) A[r * 12 + k] = row[k] * sc;
                }
            __syncwarp();
            return;
        }
#endif
        jacobi_svd<12, 12, false>(A, ws.at(WS_W), A);
    }

    CORB_HD void alphas_of(const double pw[3], double a[4]) const {  // compute_barycentric_coordinates :471-481
        for (int j = 0; j < 3; j++)
            a[1 + j] = ci[3 * j] * (pw[0] - cws[0][0]) + ci[3 * j + 1] * (pw[1] - cws[0][1]) + ci[3 * j + 2] * (pw[2] - cws[0][2]);
        a[0] = 1.0f - a[1] - a[2] - a[3];
    }
    CORB_HD void pc_of(const double a[4], double pc[3]) const {  // compute_pcs :516-525
        for (int j = 0; j < 3; j++) pc[j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }

    template <bool WARP>
    CORB_HD void choose_control_points(const PtSet& s, Ws ws) {  // :420-455
        const int n = s.n;
        double c0[3] = {0, 0, 0};
        ordered_sum<WARP, 3>(s, c0, [&](int idx, double* t) { load_pw(s, idx, t); }, coop.tbuf);
        for (int j = 0; j < 3; j++) cws[0][j] = c0[j] / n;
        double m[6] = {0, 0, 0, 0, 0, 0};  // upper triangle of PW0^T PW0: 00 01 02 11 12 22
        ordered_sum<WARP, 6>(s, m, [&](int idx, double* t) {
            double pw[3];
            load_pw(s, idx, pw);
            const double d0 = pw[0] - cws[0][0], d1 = pw[1] - cws[0][1], d2 = pw[2] - cws[0][2];
            t[0] = d0 * d0; t[1] = d0 * d1; t[2] = d0 * d2; t[3] = d1 * d1; t[4] = d1 * d2; t[5] = d2 * d2;
        }, coop.tbuf);
        const double ptp[9] = {m[0], m[1], m[2], m[1], m[3], m[4], m[2], m[4], m[5]};
        svd_small<3, 3, false>(ptp, ws);
        const Ws uct = ws.at(WS_S), dc = ws.at(WS_SW);
        for (int i = 1; i < 4; i++) {
            const double k = sqrt(dc[i - 1] / n);
            for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + k * uct[3 * (i - 1) + j];
        }
    }

    CORB_HD void compute_barycentric(Ws ws) {  // :457-470
        double cc[9];
        for (int i = 0; i < 3; i++)
            for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
        svd_invert3(cc, ci, ws);
    }

    // the two rows fill_M (:484-500) writes for one correspondence, column c: M1 = a[c/3] * {fu, 0, uc-u}[c%3],
    // M2 = a[c/3] * {0, fv, vc-v}[c%3]
    CORB_HD void m_rows(const PtSet& s, int idx, double* M1, double* M2) const {
        double pw[3], a[4];
        load_pw(s, idx, pw);
        alphas_of(pw, a);
        const double u = (double)s.p2d[2 * idx], v = (double)s.p2d[2 * idx + 1];
        for (int i = 0; i < 4; i++) {
            M1[3 * i] = a[i] * fu; M1[3 * i + 1] = 0.0; M1[3 * i + 2] = a[i] * (uc - u);
            M2[3 * i] = 0.0; M2[3 * i + 1] = a[i] * fv; M2[3 * i + 2] = a[i] * (vc - v);
        }
    }

    // M^T M accumulated row by row (cvMulTransposed: every element summed over the rows in order), then its SVD: rows
    // of ws[WS_A..] = Ut. WARP: lane l owns the upper-triangle entries l, l+32, l+64 and adds every correspondence's two
    // products to them in set order; the entries are then broadcast into every lane's copy of the workspace.
    template <bool WARP>
    CORB_HD void mtm_build(const PtSet& s, Ws ws) {
        const Ws A = ws.at(WS_A);
#ifdef __CUDA_ARCH__
        if (WARP && s.mask) {
            const int lane = threadIdx.x & 31;
            int ei[3], ej[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {  // entry e = lane + 32 q of the row-major upper triangle -> (i, j)
                int e = lane + 32 * q, i = 0;
                if (e >= 78) e = 77;
                while (e >= 12 - i) { e -= 12 - i; i++; }
                ei[q] = i; ej[q] = i + e;
            }
            double acc[3] = {0, 0, 0};
            for (int w = 0; w < s.n_words; w++) {
                for (uint32_t b = s.mask[w]; b; b &= b - 1) {
                    double M1[12], M2[12];
                    m_rows(s, w * 32 + __ffs((int)b) - 1, M1, M2);
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        double x1 = 0, y1 = 0, x2 = 0, y2 = 0;
#pragma unroll
                        for (int c = 0; c < 12; c++) {  // register selects instead of dynamic indexing
                            if (c == ei[q]) { x1 = M1[c]; x2 = M2[c]; }
                            if (c == ej[q]) { y1 = M1[c]; y2 = M2[c]; }
                        }
                        acc[q] += x1 * y1;
                        acc[q] += x2 * y2;
                    }
                }
            }
            for (int e = 0, i = 0, j = 0; e < 78; e++) {
                const double v = __shfl_sync(0xffffffffu, e < 32 ? acc[0] : e < 64 ? acc[1] : acc[2], e & 31);
                A[i * 12 + j] = v;
                A[j * 12 + i] = v;
                if (++j == 12) { i++; j = i; }
            }
            return;
        }
#endif
        for (int i = 0; i < 144; i++) A[i] = 0;
        PtIter it;
        int idx;
        double M1[12], M2[12];
        for (it_begin(it); it_next(s, it, idx);) {
            m_rows(s, idx, M1, M2);
            for (int i = 0; i < 12; i++)
                for (int j = i; j < 12; j++) {
                    double acc = A[i * 12 + j];
                    acc += M1[i] * M1[j];
                    acc += M2[i] * M2[j];
                    A[i * 12 + j] = acc;
                }
        }
        for (int i = 0; i < 12; i++)
            for (int j = i + 1; j < 12; j++) A[j * 12 + i] = A[i * 12 + j];
    }

    CORB_HD void compute_L_6x10(Ws ws, double* L) const {  // :787-829
        const Ws ut = ws.at(WS_A);
        int a = 0, b = 1;
        for (int i = 0; i < 6; i++) {
            double dv[4][3];
            for (int k = 0; k < 4; k++) {
                const Ws v = ut.at(12 * (11 - k));
                dv[k][0] = v[3 * a] - v[3 * b];
                dv[k][1] = v[3 * a + 1] - v[3 * b + 1];
                dv[k][2] = v[3 * a + 2] - v[3 * b + 2];
            }
            double* row = L + 10 * i;
            row[0] = dot3(dv[0], dv[0]);
            row[1] = 2.0f * dot3(dv[0], dv[1]);
            row[2] = dot3(dv[1], dv[1]);
            row[3] = 2.0f * dot3(dv[0], dv[2]);
            row[4] = 2.0f * dot3(dv[1], dv[2]);
            row[5] = dot3(dv[2], dv[2]);
            row[6] = 2.0f * dot3(dv[0], dv[3]);
            row[7] = 2.0f * dot3(dv[1], dv[3]);
            row[8] = 2.0f * dot3(dv[2], dv[3]);
            row[9] = dot3(dv[3], dv[3]);
            b++;
            if (b > 3) { a++; b = a + 1; }
        }
    }

    CORB_HD void compute_ccs(const double* betas, Ws ws) {  // :502-514
        const Ws ut = ws.at(WS_A);
        for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0f;
        for (int i = 0; i < 4; i++) {
            const Ws v = ut.at(12 * (11 - i));
            for (int j = 0; j < 4; j++)
                for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
        }
    }

    // compute_R_and_t :676-687 = compute_ccs, compute_pcs, solve_for_sign, estimate_R_and_t, reprojection_error
    template <bool WARP>
    CORB_HD double compute_R_and_t(const PtSet& s, const double* betas, Ws ws, double R[3][3], double t[3]) {
        const int n = s.n;
        compute_ccs(betas, ws);
        // solve_for_sign (:660-674): the sign of the first point's depth; negating ccs negates every pc exactly
        {
            PtIter it;
            int idx;
            it_begin(it);
            if (it_next(s, it, idx)) {
                double pw[3], a[4], pc[3];
                load_pw(s, idx, pw);
                alphas_of(pw, a);
                pc_of(a, pc);
                if (pc[2] < 0.0)
                    for (int i = 0; i < 4; i++)
                        for (int j = 0; j < 3; j++) ccs[i][j] = -ccs[i][j];
            }
        }
        // estimate_R_and_t :587-651
        double c6[6] = {0, 0, 0, 0, 0, 0};  // pc0 | pw0
        ordered_sum<WARP, 6>(s, c6, [&](int idx, double* tt) {
            double a[4];
            load_pw(s, idx, tt + 3);
            alphas_of(tt + 3, a);
            pc_of(a, tt);
        }, coop.tbuf);
        double pc0[3], pw0[3];
        for (int j = 0; j < 3; j++) { pc0[j] = c6[j] / n; pw0[j] = c6[3 + j] / n; }
        double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        ordered_sum<WARP, 9>(s, abt, [&](int idx, double* tt) {
            double pw[3], a[4], pc[3];
            load_pw(s, idx, pw);
            alphas_of(pw, a);
            pc_of(a, pc);
            for (int j = 0; j < 3; j++) {
                tt[3 * j] = (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
                tt[3 * j + 1] = (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
                tt[3 * j + 2] = (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
            }
        }, coop.tbuf);
        svd_small<3, 3, true>(abt, ws);
        const Ws Ut = ws.at(WS_S), Vt = ws.at(WS_SV);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[i][j] = Ut[i] * Vt[j] + Ut[3 + i] * Vt[3 + j] + Ut[6 + i] * Vt[6 + j];
        const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                           R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
        if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
        t[0] = pc0[0] - dot3(R[0], pw0);
        t[1] = pc0[1] - dot3(R[1], pw0);
        t[2] = pc0[2] - dot3(R[2], pw0);
        // reprojection_error :568-585
        double sum2 = 0.0;
        ordered_sum<WARP, 1>(s, &sum2, [&](int idx, double* tt) {
            double pw[3];
            load_pw(s, idx, pw);
            const double Xc = dot3(R[0], pw) + t[0];
            const double Yc = dot3(R[1], pw) + t[1];
            const double inv_Zc = 1.0 / (dot3(R[2], pw) + t[2]);
            const double ue = uc + fu * Xc * inv_Zc;
            const double ve = vc + fv * Yc * inv_Zc;
            const double u = (double)s.p2d[2 * idx], v = (double)s.p2d[2 * idx + 1];
            tt[0] = sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
        }, coop.tbuf);
        return sum2 / n;
    }

    template <int which>
    CORB_HD static void find_betas(const double* L, const double* rho, double* betas, Ws ws) {  // :692-785
        double l[30], b[5];
        if (which == 1) {
            for (int i = 0; i < 6; i++) { l[4 * i] = L[10 * i]; l[4 * i + 1] = L[10 * i + 1]; l[4 * i + 2] = L[10 * i + 3]; l[4 * i + 3] = L[10 * i + 6]; }
            svd_solve<6, 4>(l, rho, b, ws);
            if (b[0] < 0) {
                betas[0] = sqrt(-b[0]); betas[1] = -b[1] / betas[0]; betas[2] = -b[2] / betas[0]; betas[3] = -b[3] / betas[0];
            } else {
                betas[0] = sqrt(b[0]); betas[1] = b[1] / betas[0]; betas[2] = b[2] / betas[0]; betas[3] = b[3] / betas[0];
            }
            return;
        }
        constexpr int nc = which == 2 ? 3 : 5;
        for (int i = 0; i < 6; i++)
            for (int k = 0; k < nc; k++) l[nc * i + k] = L[10 * i + k];
        svd_solve<6, nc>(l, rho, b, ws);
        if (b[0] < 0) {
            betas[0] = sqrt(-b[0]);
            betas[1] = (b[2] < 0) ? sqrt(-b[2]) : 0.0;
        } else {
            betas[0] = sqrt(b[0]);
            betas[1] = (b[2] > 0) ? sqrt(b[2]) : 0.0;
        }
        if (b[1] < 0) betas[0] = -betas[0];
        betas[2] = which == 2 ? 0.0 : b[3] / betas[0];
        betas[3] = 0.0;
    }

    // Householder least squares 6 x 4 (:888-977), including the reference's column-maximum scan that stops one row
    // short. A zero column returns with x unchanged.
    CORB_HD static void qr_solve(double* A, double* b, double* X) {
        const int nr = 6, nc = 4;
        double A1[4], A2[4];
        for (int k = 0; k < nc; k++) {
            double eta = fabs(A[k * nc + k]);
            for (int i = k + 1; i < nr; i++) {
                const double elt = fabs(A[(i - 1) * nc + k]);
                if (eta < elt) eta = elt;
            }
            if (eta == 0) return;
            double sum = 0.0;
            const double inv_eta = 1. / eta;
            for (int i = k; i < nr; i++) {
                A[i * nc + k] *= inv_eta;
                sum += A[i * nc + k] * A[i * nc + k];
            }
            double sigma = sqrt(sum);
            if (A[k * nc + k] < 0) sigma = -sigma;
            A[k * nc + k] += sigma;
            A1[k] = sigma * A[k * nc + k];
            A2[k] = -eta * sigma;
            for (int j = k + 1; j < nc; j++) {
                double sm = 0;
                for (int i = k; i < nr; i++) sm += A[i * nc + k] * A[i * nc + j];
                const double tau = sm / A1[k];
                for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
            }
        }
        for (int j = 0; j < nc; j++) {
            double tau = 0;
            for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
            tau /= A1[j];
            for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
        }
        X[nc - 1] = b[nc - 1] / A2[nc - 1];
        for (int i = nc - 2; i >= 0; i--) {
            double sum = 0;
            for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
            X[i] = (b[i] - sum) / A2[i];
        }
    }

    CORB_HD static void gauss_newton(const double* L, const double* rho, double betas[4]) {  // :841-886
        double A[24], b[6], x[4] = {0, 0, 0, 0};
        for (int k = 0; k < 5; k++) {
            for (int i = 0; i < 6; i++) {
                const double* rowL = L + i * 10;
                double* rowA = A + i * 4;
                rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
                rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
                rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
                rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
                b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                                 rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                                 rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                                 rowL[9] * betas[3] * betas[3]);
            }
            qr_solve(A, b, x);
            for (int i = 0; i < 4; i++) betas[i] += x[i];
        }
    }

    // one of the three beta approximations of compute_pose (:545-557): betas, Gauss-Newton, R | t, reprojection error
    template <bool WARP, int which>
    CORB_HD double branch_raw(const PtSet& s, Ws ws, const double* L, const double* rho, double* Rt) {
        double betas[4], R[3][3], t[3];
        find_betas<which>(L, rho, betas, ws);
        gauss_newton(L, rho, betas);
        const double err = compute_R_and_t<WARP>(s, betas, ws, R, t);
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[i][j];
            Rt[9 + i] = t[i];
        }
        return err;
    }
    // ... and the selection N = 1; if (e2 < e1) N = 2; if (e3 < e[N]) N = 3 (:559-561): strict improvements only (NaN never wins)
    template <bool WARP, int which>
    CORB_HD void branch(const PtSet& s, Ws ws, const double* L, const double* rho, double& best_err, double* Rt) {
        double cand[12];
        const double err = branch_raw<WARP, which>(s, ws, L, rho, cand);
        if (which == 1 || err < best_err) {
            best_err = err;
            for (int i = 0; i < 12; i++) Rt[i] = cand[i];
        }
    }

    // compute_pose :527-574. Rt = R (row-major 9) followed by t (3).
    template <bool WARP = false>
    CORB_HD double compute_pose(const PtSet& s, Ws ws, double* Rt) { return compute_pose_part<WARP>(s, ws, 0, Rt); }

    // only_branch = 0: the whole compute_pose. 1..3 (warp-uniform): everything up to the SVD, then that beta approximation
    // alone; the caller applies the selection rule to the three (error, pose) results. In a lead-only team the scalar
    // phases run on the team's lane 0 and the other lanes only join the wavefront (their Rt and result are not written).
    template <bool WARP>
    CORB_HD double compute_pose_part(const PtSet& s, Ws ws, int only_branch, double* Rt) {
        const bool lead = !coop.lead_only || coop.tl == 0;
        if (lead) {
            choose_control_points<WARP>(s, ws);
            compute_barycentric(ws);
            mtm_build<WARP>(s, ws);
        }
#ifdef __CUDA_ARCH__
        if (coop.lead_only) __syncwarp();
#endif
        svd12(ws);
        double best_err = 0;
        if (lead) {
            double L[60], rho[6];
            compute_L_6x10(ws, L);
            rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);  // compute_rho :831-839
            rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
            if (only_branch == 0) {
                branch<WARP, 1>(s, ws, L, rho, best_err, Rt);
                branch<WARP, 2>(s, ws, L, rho, best_err, Rt);
                branch<WARP, 3>(s, ws, L, rho, best_err, Rt);
            } else if (only_branch == 1) {
                best_err = branch_raw<WARP, 1>(s, ws, L, rho, Rt);
            } else if (only_branch == 2) {
                best_err = branch_raw<WARP, 2>(s, ws, L, rho, Rt);
            } else {
                best_err = branch_raw<WARP, 3>(s, ws, L, rho, Rt);
            }
        }
        return best_err;
    }
};

// CheckInliers (:349-383) for one correspondence: the float / double mix of the reference, expression by expression.
CORB_HD inline bool is_inlier(const double* Rt, double fu, double fv, double uc, double vc, const float* P, const float* p, float max_err) {
    const float Xc = (float)(Rt[0] * (double)P[0] + Rt[1] * (double)P[1] + Rt[2] * (double)P[2] + Rt[9]);
    const float Yc = (float)(Rt[3] * (double)P[0] + Rt[4] * (double)P[1] + Rt[5] * (double)P[2] + Rt[10]);
    const float invZc = (float)(1 / (Rt[6] * (double)P[0] + Rt[7] * (double)P[1] + Rt[8] * (double)P[2] + Rt[11]));
    const double ue = uc + fu * (double)Xc * (double)invZc;
    const double ve = vc + fv * (double)Yc * (double)invZc;
    const float distX = (float)((double)p[0] - ue);
    const float distY = (float)((double)p[1] - ve);
    const float error2 = distX * distX + distY * distY;
    return error2 < max_err;
}

}  // namespace pnp
}  // namespace corb
