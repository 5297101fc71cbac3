// Dense border block of the reduced camera system (the keyframes ordered last: separators of the band chunks and the ends of
// long-range loop / map-fusion links) on the WHOLE GPU with fp64 tensor-core MMA - the dense part of what
// g2o's LinearSolverEigen::solve (linear_solver_eigen.h:94-124) does for the reduced camera block.
//
//   k_ba_border_chol  : tiled right-looking Cholesky, in place in the block skyline. Tile = 48 x 48 doubles (8 keyframes).
//                       Per tile column K: every CTA factors tile (K, K) in shared memory (redundantly: no barrier between
//                       POTRF and TRSM) and solves its tiles X_IK = A_IK L_KK^-T; grid barrier; the trailing updates
//                       A_IJ -= X_IK X_JK^T run as m8n8k4 fp64 MMA (DMMA), one CTA per tile; grid barrier.
//                       Then one CTA substitutes forward and backward through the factor for the border right-hand side.
//   k_ba_border_to_band: x_band -= sum_j L_j,band^T x_j over the border rows j, one warp per band column.
//
// Together they replace the one-CTA left-looking sweep + the one-CTA backward pass over the border rows (0.77 + 0.61 ms
// per LM trial at P = 2000 in profiles/r01d_launches_ba.csv).
#pragma once
#include <cooperative_groups.h>

namespace corb {

constexpr int kBT = 48;         // tile edge in doubles (8 pose blocks of 6)
constexpr int kBTB = kBT / 6;   // pose blocks per tile edge
constexpr int kBLd = kBT + 1;   // shared-memory leading dimension (odd: column walks are conflict free)
constexpr int kBcThreads = 256;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct BorderView {
    BaDev d;
    int n_band, nbord, nt;
    const long long* rowbase;  // shared memory: block offset of (a, 0) in the skyline, per border row a
    // block (a, k) of the border matrix (border-local indices, k <= a) in the skyline
    __device__ __forceinline__ double* blk(int a, int k) const { return d.S + (size_t)(rowbase[a] + k) * 36; }
};

// tile (I, K), I >= K, into shared memory as a dense 48 x 48 array; padding rows / columns continue the identity
__device__ __forceinline__ void border_load_tile(const BorderView& v, int I, int K, double* s, int tid) {
    for (int e = tid; e < kBTB * kBTB * 36; e += kBcThreads) {
        const int bb = e / 36, w = e - bb * 36;
        const int br = bb / kBTB, bc = bb - br * kBTB;
        const int a = I * kBTB + br, k = K * kBTB + bc;
        const int r = br * 6 + w / 6, c = bc * 6 + w % 6;
        double x = 0.0;
        if (a < v.nbord && k < v.nbord) {
            if (k < a || (k == a && w / 6 >= w % 6)) x = v.blk(a, k)[w];
        } else if (I == K && r == c) {
            x = 1.0;
        }
        s[r * kBLd + c] = x;
    }
}

// Rank-6 update inside a shared-memory tile on the tensor cores: C[r][c] -= sum_{p < 6} A[r][p] B[c][p] for r < nr, c < nc (and
// c <= r when `lower`), as 8 x 8 output tiles of m8n8k4 fp64 MMA with the k = 6 panel padded to 8. A, B, C are row-major with
// leading dimension kBLd; every output element has exactly one owner lane, so the read-modify-write needs no atomics.
__device__ __forceinline__ void border_rank6_update(double* C, const double* A, const double* B, int nr, int nc, bool lower, int tid) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int tr = (nr + 7) >> 3, tc = (nc + 7) >> 3;
    for (int t = warp; t < tr * tc; t += kBcThreads / 32) {
        const int ti = t / tc, tj = t - ti * tc;
        if (lower && tj > ti) continue;
        const int r = ti * 8 + g, c = tj * 8 + 2 * q;      // outputs (r, c) and (r, c + 1)
        const int ra = min(ti * 8 + g, nr - 1), rb = min(tj * 8 + g, nc - 1);  // operand rows, clamped (their products are not stored)
        const bool ok0 = r < nr && c < nc && (!lower || c <= r), ok1 = r < nr && c + 1 < nc && (!lower || c + 1 <= r);
        double c0 = ok0 ? C[r * kBLd + c] : 0.0, c1 = ok1 ? C[r * kBLd + c + 1] : 0.0;
        dmma_m8n8k4(c0, c1, -A[ra * kBLd + q], B[rb * kBLd + q]);
        dmma_m8n8k4(c0, c1, q < 2 ? -A[ra * kBLd + 4 + q] : 0.0, q < 2 ? B[rb * kBLd + 4 + q] : 0.0);
        if (ok0) C[r * kBLd + c] = c0;
        if (ok1) C[r * kBLd + c + 1] = c1;
    }
}

// Cholesky of a 48 x 48 tile in shared memory (lower; the strict upper part is left untouched), 6 columns per step.
// inv[c] = 1 / L_cc. Returns false (uniformly) when a pivot is not positive.
__device__ __forceinline__ bool border_potrf(double* s, double* inv, int* fail, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    for (int b = 0; b < kBTB; b++) {
        const int c0 = b * 6;
        if (warp == 0) {  // the 6 x 6 diagonal block in registers: lane = one of its 21 lower-triangle entries
            int r = 0, c = 0;
            if (lane < 21) {
                r = lane >= 15 ? 5 : lane >= 10 ? 4 : lane >= 6 ? 3 : lane >= 3 ? 2 : lane >= 1 ? 1 : 0;
                c = lane - r * (r + 1) / 2;
            }
            double a = lane < 21 ? s[(c0 + r) * kBLd + c0 + c] : 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                double akk = __shfl_sync(0xffffffffu, a, k * (k + 1) / 2 + k);
                if (!(akk > 0.0)) { if (lane == 0) *fail = 1; akk = 1.0; }
                const double iv = rsqrt(akk), dk = akk * iv;
                if (lane == 0) inv[c0 + k] = iv;
                if (lane < 21 && c == k) a = (r == k) ? dk : a * iv;
                const double lrk = __shfl_sync(0xffffffffu, a, r * (r + 1) / 2 + min(k, r));
                const double lck = __shfl_sync(0xffffffffu, a, c * (c + 1) / 2 + min(k, c));
                if (lane < 21 && c > k) a -= lrk * lck;
            }
            if (lane < 21) s[(c0 + r) * kBLd + c0 + c] = a;
        }
        __syncthreads();
        const int below = kBT - c0 - 6;
        if (tid < below) {  // panel rows: x L_bb^T = a, one row per thread
            double* row = s + (c0 + 6 + tid) * kBLd + c0;
            double x[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double t = row[c];
#pragma unroll
                for (int p = 0; p < c; p++) t -= x[p] * s[(c0 + c) * kBLd + c0 + p];
                x[c] = t * inv[c0 + c];
            }
#pragma unroll
            for (int c = 0; c < 6; c++) row[c] = x[c];
        }
        __syncthreads();
        // trailing update (lower part): S[c0 + 6 + r][c0 + 6 + c] -= panel[r] . panel[c]
        border_rank6_update(s + (c0 + 6) * kBLd + c0 + 6, s + (c0 + 6) * kBLd + c0, s + (c0 + 6) * kBLd + c0, below, below, true, tid);
        __syncthreads();
    }
    return *fail == 0;
}

// X <- X L^-T for a 48 x 48 tile X (in place), L lower with reciprocal diagonal inv. Right-looking, 6 columns per step: the
// block column is solved (one row per thread, a chain of 6), then every remaining column takes its six terms at once
// (8 x 8 output tiles on the tensor cores).
__device__ __forceinline__ void border_trsm(double* x, const double* l, const double* inv, int tid) {
    for (int b = 0; b < kBTB; b++) {
        const int c0 = b * 6;
        if (tid < kBT) {
            double* row = x + tid * kBLd + c0;
            double v[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double t = row[c];
#pragma unroll
                for (int p = 0; p < c; p++) t -= v[p] * l[(c0 + c) * kBLd + c0 + p];
                v[c] = t * inv[c0 + c];
            }
#pragma unroll
            for (int c = 0; c < 6; c++) row[c] = v[c];
        }
        __syncthreads();
        // X[:, c0 + 6 + c] -= X[:, c0 .. c0 + 5] . L[c0 + 6 + c][c0 .. c0 + 5] for the remaining columns
        border_rank6_update(x + c0 + 6, x + c0, l + (c0 + 6) * kBLd + c0, kBT, kBT - c0 - 6, false, tid);
        __syncthreads();
    }
}

// y <- L^-1 y (forward) or y <- L^-T y (backward) for one tile, right-looking: solve 6 entries, then update the rest of the tile
__device__ __forceinline__ void border_trsv(double* y, const double* l, const double* inv, int tid, bool transposed) {
    for (int s6 = 0; s6 < kBTB; s6++) {
        const int b = transposed ? kBTB - 1 - s6 : s6, c0 = b * 6;
        if (tid == 0) {
            double v[6];
            if (!transposed) {
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double t = y[c0 + r];
#pragma unroll
                    for (int p = 0; p < r; p++) t -= l[(c0 + r) * kBLd + c0 + p] * v[p];
                    v[r] = t * inv[c0 + r];
                }
            } else {
#pragma unroll
                for (int r = 5; r >= 0; r--) {
                    double t = y[c0 + r];
#pragma unroll
                    for (int p = r + 1; p < 6; p++) t -= l[(c0 + p) * kBLd + c0 + r] * v[p];
                    v[r] = t * inv[c0 + r];
                }
            }
#pragma unroll
            for (int r = 0; r < 6; r++) y[c0 + r] = v[r];
        }
        __syncthreads();
        if (!transposed) {  // y[i] -= sum_p L[i][c0 + p] y[c0 + p] for the entries below
            const int i = c0 + 6 + tid;
            if (i < kBT) {
                double t = 0;
#pragma unroll
                for (int p = 0; p < 6; p++) t += l[i * kBLd + c0 + p] * y[c0 + p];
                y[i] -= t;
            }
        } else if (tid < c0) {  // y[i] -= sum_p L[c0 + p][i] y[c0 + p] for the entries above
            double t = 0;
#pragma unroll
            for (int p = 0; p < 6; p++) t += l[(c0 + p) * kBLd + tid] * y[c0 + p];
            y[tid] -= t;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void border_store_tile(const BorderView& v, int I, int K, const double* s, int tid, bool diag_zero_upper) {
    for (int e = tid; e < kBTB * kBTB * 36; e += kBcThreads) {
        const int bb = e / 36, w = e - bb * 36;
        const int br = bb / kBTB, bc = bb - br * kBTB;
        const int a = I * kBTB + br, k = K * kBTB + bc;
        if (a >= v.nbord || k >= v.nbord || k > a) continue;
        const int r = br * 6 + w / 6, c = bc * 6 + w % 6;
        double x = s[r * kBLd + c];
        if (k == a && w / 6 < w % 6) { if (!diag_zero_upper) continue; x = 0.0; }
        v.blk(a, k)[w] = x;
    }
}

#ifdef CORB_CHOL_TRACE
#define CH_T(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _n = clock64(); ch_t[i] += _n - ch_last; ch_last = _n; } } while (0)
#else
#define CH_T(i) do { } while (0)
#endif
__global__ void __launch_bounds__(kBcThreads) k_ba_border_chol(BaDev d, int n_band) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double bsm[];
    double* sA = bsm;                 // diagonal tile / X_I
    double* sB = sA + kBT * kBLd;     // X tile being solved / X_J
    double* sinv = sB + kBT * kBLd;   // [48]
    double* sv = sinv + kBT;          // [nbord * 6 padded] right-hand side
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    BorderView v;
    v.d = d; v.n_band = n_band; v.nbord = d.Pf - n_band; v.nt = (v.nbord + kBTB - 1) / kBTB;
    const int nt = v.nt, G = gridDim.x, me = blockIdx.x;
    long long* rowbase = reinterpret_cast<long long*>(sv + (size_t)nt * kBT);
    for (int a = tid; a < v.nbord; a += kBcThreads) rowbase[a] = (long long)(d.rowoff[n_band + a] - d.first[n_band + a]) + n_band;
    v.rowbase = rowbase;
    if (tid == 0) s_fail = d.scalars[4] != 0.0;
    __syncthreads();
    bool dead = s_fail != 0;  // the band factorisation failed: nothing to do (uniform over the grid)
#ifdef CORB_CHOL_TRACE
    long long ch_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ch_last = clock64();
#endif
    for (int K = 0; K < nt && !dead; K++) {
        // ---- POTRF (K, K) in every CTA that has a TRSM tile (CTA 0 always, it owns the write-back) + TRSM of column K
        const int ntr = nt - K - 1;
        const bool work = me == 0 || me < ntr;
        if (work) {
            border_load_tile(v, K, K, sA, tid);
            __syncthreads();
            CH_T(0);
            border_potrf(sA, sinv, &s_fail, tid);
            CH_T(1);
            if (me == 0) {
                border_store_tile(v, K, K, sA, tid, true);
                for (int c = tid; c < kBT; c += kBcThreads)
                    if (K * kBT + c < v.nbord * 6) d.invd[(size_t)n_band * 6 + K * kBT + c] = sinv[c];
                if (tid == 0 && s_fail) d.scalars[4] = 1.0;
            }
            // forward substitution rides along: y_K = L_KK^-1 b_K (every working CTA, redundantly), b_I -= X_IK y_K by the
            // CTA that owns tile row I in this step
            double* sy = sv;  // [48]
            for (int c = tid; c < kBT; c += kBcThreads) sy[c] = K * kBT + c < v.nbord * 6 ? d.xp[(size_t)n_band * 6 + K * kBT + c] : 0.0;
            __syncthreads();
            border_trsv(sy, sA, sinv, tid, false);
            if (me == 0)
                for (int c = tid; c < kBT; c += kBcThreads)
                    if (K * kBT + c < v.nbord * 6) d.xp[(size_t)n_band * 6 + K * kBT + c] = sy[c];
            CH_T(2);
            for (int t = me; t < ntr; t += G) {
                const int I = K + 1 + t;
                border_load_tile(v, I, K, sB, tid);
                __syncthreads();
                border_trsm(sB, sA, sinv, tid);
                border_store_tile(v, I, K, sB, tid, false);
                if (tid < kBT && I * kBT + tid < v.nbord * 6) {
                    double acc = 0;
                    for (int p = 0; p < kBT; p++) acc += sB[tid * kBLd + p] * sy[p];
                    d.xp[(size_t)n_band * 6 + I * kBT + tid] -= acc;
                }
                __syncthreads();
            }
        }
        CH_T(3);
        __threadfence();
        grid.sync();
        CH_T(4);
        if (d.scalars[4] != 0.0) { dead = true; break; }  // a pivot failed somewhere: every CTA sees it after the barrier
        // ---- trailing updates A_IJ -= X_IK X_JK^T (K < J <= I), one tile per CTA, fp64 MMA
        const int m = nt - K - 1, ntask = m * (m + 1) / 2;
        for (int t = me; t < ntask; t += G) {
            int ii = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while (ii * (ii + 1) / 2 > t) ii--;
            while ((ii + 1) * (ii + 2) / 2 <= t) ii++;
            const int jj = t - ii * (ii + 1) / 2;
            const int I = K + 1 + ii, J = K + 1 + jj;
            border_load_tile(v, I, K, sA, tid);
            if (J != I) border_load_tile(v, J, K, sB, tid);
            __syncthreads();
            const double* xi = sA;
            const double* xj = J == I ? sA : sB;
            const int g = lane >> 2, q = lane & 3;
            for (int sub = warp; sub < 36; sub += kBcThreads / 32) {  // 6 x 6 output sub-tiles of 8 x 8
                const int si = sub / 6, sj = sub - si * 6;
                if (J == I && sj > si) continue;
                const int r = si * 8 + g, c = sj * 8 + 2 * q;  // this lane's two outputs: (r, c), (r, c + 1)
                const int a = I * kBTB + r / 6, k = J * kBTB + c / 6;
                const bool live = a < v.nbord && k < v.nbord && k <= a;
                double* out = live ? v.blk(a, k) + (r % 6) * 6 + c % 6 : nullptr;
                double c0 = live ? out[0] : 0.0, c1 = live ? out[1] : 0.0;
#pragma unroll
                for (int k0 = 0; k0 < kBT; k0 += 4)
                    dmma_m8n8k4(c0, c1, -xi[(si * 8 + g) * kBLd + k0 + q], xj[(sj * 8 + g) * kBLd + k0 + q]);
                if (live) { out[0] = c0; out[1] = c1; }
            }
            __syncthreads();
        }
        CH_T(5);
        __threadfence();
        grid.sync();
        CH_T(6);
    }
    if (me != 0 || dead) return;
    // ---- L^T x = y for the border right-hand side (y was formed along the factorisation), tile by tile
    const int n6 = v.nbord * 6;
    for (int i = tid; i < nt * kBT; i += kBcThreads) sv[i] = i < n6 ? d.xp[(size_t)n_band * 6 + i] : 0.0;
    __syncthreads();
    for (int K = nt - 1; K >= 0; K--) {  // backward, right-looking: x_K = L_KK^-T y_K, then y_J -= L_KJ^T x_K for the tiles J < K
        border_load_tile(v, K, K, sA, tid);
        for (int c = tid; c < kBT; c += kBcThreads) sinv[c] = K * kBT + c < n6 ? d.invd[(size_t)n_band * 6 + K * kBT + c] : 1.0;
        __syncthreads();
        border_trsv(sv + K * kBT, sA, sinv, tid, true);
        // every entry (block column k < 8 K, component c) of the rows above takes its 48 terms: the blocks (a, k) of a row a
        // are contiguous in the skyline, so neighbouring threads read neighbouring memory
        const int a_lo = K * kBTB, a_hi = min(v.nbord, a_lo + kBTB);
        for (int o = tid; o < K * kBT; o += kBcThreads) {
            const int k = o / 6, c = o - k * 6;
            double t = 0;
            for (int a = a_lo; a < a_hi; a++) {
                const double* L = v.blk(a, k) + c;
                const double* x = sv + K * kBT + (a - a_lo) * 6;
#pragma unroll
                for (int p = 0; p < 6; p++) t += L[p * 6] * x[p];
            }
            sv[o] -= t;
        }
        __syncthreads();
    }
    for (int i = tid; i < n6; i += kBcThreads) d.xp[(size_t)n_band * 6 + i] = sv[i];
    CH_T(7);
#ifdef CORB_CHOL_TRACE
    if (tid == 0) printf("k_ba_border_chol nt=%d grid=%d: load diag %lld potrf %lld store+y %lld trsm %lld sync1 %lld update %lld sync2 %lld backward %lld cycles\n", nt, G, ch_t[0], ch_t[1], ch_t[2], ch_t[3], ch_t[4], ch_t[5], ch_t[6], ch_t[7]);
#endif
}

// x_i -= sum_j L_ji^T x_j for every band column i over the border rows j whose envelope reaches i (descending j)
__global__ void __launch_bounds__(256) k_ba_border_to_band(BaDev d, int n_band) {
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_band || d.scalars[4] != 0.0) return;
    const int nbord = d.Pf - n_band;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int a = nbord - 1 - lane; a >= 0; a -= 32) {
        const int j = n_band + a;
        if (d.first[j] > i) continue;
        const double* L = d.S + ((size_t)(d.rowoff[j] - d.first[j]) + i) * 36;
        const double* x = d.xp + (size_t)j * 6;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const double xr = x[r];
#pragma unroll
            for (int c = 0; c < 6; c++) acc[c] += L[r * 6 + c] * xr;
        }
    }
#pragma unroll
    for (int c = 0; c < 6; c++)
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], o);
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < 6; c++) d.xp[(size_t)i * 6 + c] -= acc[c];
}

}  // namespace corb
