// Global bundle adjustment on sm_100a (fp64): K10 linearise + J^T J block assembly, K11 Schur complement,
// K12 reduced-camera solve (block-skyline Cholesky), K13 landmark back-substitution + manifold update + chi2.
//
// Replaces Optimizer::BundleAdjustment -> g2o BlockSolver_6_3 + OptimizationAlgorithmLevenberg
// (corbslam_client/src/Optimizer.cc:54-270; Thirdparty/g2o/g2o/core/block_solver.hpp:354-604,
// optimization_algorithm_levenberg.cpp:61-189, types/types_six_dof_expmap.cpp:103-234, types/se3quat.h:217-257).
//
// Layout: edges are sorted by landmark (CSR), so a landmark's Hll, bl and its Hpl blocks are produced by one thread;
// a second CSR by pose lets one warp own a pose's Hpp/bp and one block row of the reduced system, which makes every
// floating-point sum order-deterministic without atomics. The reduced camera system of a SLAM map is block sparse
// with a narrow envelope (keyframes only share landmarks with their covisible neighbours), so it is stored as a
// block skyline and factorised right-looking by one CTA; a dense tensor-core factorisation would do orders of
// magnitude more arithmetic on structural zeros.
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <climits>
#include <atomic>
#include <chrono>
#include <numeric>
#include <mutex>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace corb {

// ------------------------------------------------------------------------------------------------ device math
__device__ __forceinline__ void quat_to_R(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
                 tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ void R_to_quat(const double* R, double* q) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        double qq[4];
        qq[i] = 0.5 * t;
        t = 0.5 / t;
        qq[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        qq[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        qq[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
        q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
    }
}
__device__ __forceinline__ void quat_normalize(double* q) {
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

struct EdgeLin {
    int D;
    double e[3], A[9], B[18];
};
// computeError / linearizeOplus of EdgeSE3ProjectXYZ (mono) and EdgeStereoSE3ProjectXYZ
template <bool kLin>
__device__ __forceinline__ void edge_eval(const double* q, const double* t, const double* cam, const double* X, const double* obs,
                                          EdgeLin& o) {
    double R[9];
    quat_to_R(q, R);
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], bf = cam[4];
    const bool stereo = !(obs[2] < 0);
    o.D = stereo ? 3 : 2;
    if (!stereo) {
        o.e[0] = obs[0] - ((x / z) * fx + cx);
        o.e[1] = obs[1] - ((y / z) * fy + cy);
        o.e[2] = 0;
    } else {  // invz is rounded to float32 in the reference (types_six_dof_expmap.cpp:150-157)
        const double invz = (double)(float)(1.0 / z);
        const double u = x * invz * fx + cx;
        o.e[0] = obs[0] - u;
        o.e[1] = obs[1] - (y * invz * fy + cy);
        o.e[2] = obs[2] - (u - bf * invz);
    }
    if (!kLin) return;
    const double z_2 = z * z;
    if (!stereo) {
        const double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                o.A[r * 3 + c] = (-1. / z) * (tmp[r * 3] * R[c] + tmp[r * 3 + 1] * R[3 + c] + tmp[r * 3 + 2] * R[6 + c]);
        o.A[6] = o.A[7] = o.A[8] = 0;
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            o.A[c] = -fx * R[c] / z + fx * x * R[6 + c] / z_2;
            o.A[3 + c] = -fy * R[3 + c] / z + fy * y * R[6 + c] / z_2;
            o.A[6 + c] = o.A[c] - bf * R[6 + c] / z_2;
        }
    }
    double* B = o.B;
    B[0] = x * y / z_2 * fx; B[1] = -(1 + (x * x / z_2)) * fx; B[2] = y / z * fx; B[3] = -1. / z * fx; B[4] = 0; B[5] = x / z_2 * fx;
    B[6] = (1 + y * y / z_2) * fy; B[7] = -x * y / z_2 * fy; B[8] = -x / z * fy; B[9] = 0; B[10] = -1. / z * fy; B[11] = y / z_2 * fy;
    if (stereo) {
        B[12] = B[0] - bf * y / z_2; B[13] = B[1] + bf * x / z_2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf / z_2;
    } else {
#pragma unroll
        for (int i = 12; i < 18; i++) B[i] = 0;
    }
}
__device__ __forceinline__ void huber(double e, double delta, double* rho0, double* rho1) {
    const double dsqr = delta * delta;
    if (e <= dsqr) { *rho0 = e; *rho1 = 1.; }
    else { const double s = sqrt(e); *rho0 = 2 * s * delta - dsqr; *rho1 = delta / s; }
}

struct BaDev {
    int P, L, E, Pf;
    double *q, *t, *X;             // estimates
    double *q0, *t0, *X0;          // backup (push/pop)
    const double* cam;
    const int *pfree, *lfree;      // free index or -1
    const int *e_pose, *e_point;   // edges sorted by landmark
    const double *e_obs, *e_info;
    const int *lm_off;             // [L+1]
    const int *pose_off, *pose_edges;  // CSR by pose: positions into the landmark-sorted edge arrays
    double *Hpp, *bp, *Hll, *bl, *W, *Dinv, *db;
    const int *first, *rowoff, *coloff, *col_rows;
    const int *coloff_b, *col_rows_b;
    const int *chunk_start;        // [n_chunks + 1] band columns of each independent chunk
    double *syrk_part;             // [border pairs][n_chunks][36] per-chunk parts of the border x border update
    const int* bstart;             // [n_border][n_chunks] first band column of chunk q a border row is coupled with (INT_MAX: none):
                                   // left of it the row's blocks in that chunk are structurally zero (fill-in only runs forward)
    double *bpart;                 // [n_border][n_chunks][6] per-chunk parts of the border rows' forward substitution  // same lists without the border rows in the band columns (bordered solve)
    double *S, *bs;                // reduced system: [S | bs] contiguous
    double *invd;                  // [Pf * 6] reciprocals of the diagonal of the Cholesky factor (triangular solves multiply)
    double *xp, *xl;
    double *partial, *scalars;     // reduction scratch; [1] scale part, [2] xx, [3] max diag, [4] fail flag, [6] chi2, [7] stop votes
    int robust;
    double delta2d, delta3d;
};

// deterministic block reduction (sum or max) of one double per thread; result valid in thread 0
template <bool kMax>
__device__ __forceinline__ double block_reduce(double v, double* sm) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double u = __shfl_down_sync(0xffffffffu, v, o);
        v = kMax ? fmax(v, u) : v + u;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? sm[lane] : (kMax ? 0.0 : 0.0);
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double u = __shfl_down_sync(0xffffffffu, v, o);
            v = kMax ? fmax(v, u) : v + u;
        }
    }
    __syncthreads();
    return v;
}
template <bool kMax>
__global__ void __launch_bounds__(256) k_reduce_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sm[32];
    double v = 0;
    for (int i = threadIdx.x; i < n; i += 256) v = kMax ? fmax(v, partial[i]) : v + partial[i];
    v = block_reduce<kMax>(v, sm);
    if (threadIdx.x == 0) *out = v;
}

// ---- chi2 over active edges (computeActiveErrors + activeRobustChi2)
__global__ void __launch_bounds__(256) k_ba_chi2(BaDev d) {
    __shared__ double sm[32];
    const int e = blockIdx.x * 256 + threadIdx.x;
    double c = 0;
    if (e < d.E) {
        const int pi = d.e_pose[e], li = d.e_point[e];
        if (d.pfree[pi] >= 0 || d.lfree[li] >= 0) {
            EdgeLin el;
            edge_eval<false>(d.q + 4 * pi, d.t + 3 * pi, d.cam + 5 * pi, d.X + 3 * li, d.e_obs + 3 * e, el);
            const double om = d.e_info[e];
            c = (el.e[0] * el.e[0] + el.e[1] * el.e[1] + el.e[2] * el.e[2]) * om;
            if (d.robust) {
                double r0, r1;
                huber(c, el.D == 2 ? d.delta2d : d.delta3d, &r0, &r1);
                c = r0;
            }
        }
    }
    c = block_reduce<false>(c, sm);
    if (threadIdx.x == 0) d.partial[blockIdx.x] = c;
}

__device__ __forceinline__ void robust_weights(const BaDev& d, const EdgeLin& el, double& om, double* omega_r) {
#pragma unroll
    for (int r = 0; r < 3; r++) omega_r[r] = -om * el.e[r];
    if (d.robust) {
        const double c = (el.e[0] * el.e[0] + el.e[1] * el.e[1] + el.e[2] * el.e[2]) * om;
        double r0, r1;
        huber(c, el.D == 2 ? d.delta2d : d.delta3d, &r0, &r1);
#pragma unroll
        for (int r = 0; r < 3; r++) omega_r[r] *= r1;
        om *= r1;
    }
}

// ---- K10a: one thread per landmark: Hll, bl and the Hpl blocks W_e = B^T Omega A of its edges
__global__ void __launch_bounds__(128) k_ba_build_lm(BaDev d) {
    const int l = blockIdx.x * 128 + threadIdx.x;
    if (l >= d.L) return;
    const bool lf = d.lfree[l] >= 0;
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    const double* X = d.X + 3 * l;
    for (int e = d.lm_off[l]; e < d.lm_off[l + 1]; e++) {
        const int pi = d.e_pose[e];
        const bool pf = d.pfree[pi] >= 0;
        double* w = d.W + (size_t)e * 18;
        if (!lf || !pf) {
#pragma unroll
            for (int i = 0; i < 18; i++) w[i] = 0;
        }
        if (!lf) continue;
        EdgeLin el;
        edge_eval<true>(d.q + 4 * pi, d.t + 3 * pi, d.cam + 5 * pi, X, d.e_obs + 3 * e, el);
        double om = d.e_info[e], omega_r[3];
        robust_weights(d, el, om, omega_r);
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int c = 0; c < 3; c++) H[a * 3 + c] += (el.A[a] * el.A[c] + el.A[3 + a] * el.A[3 + c] + el.A[6 + a] * el.A[6 + c]) * om;
            b[a] += el.A[a] * omega_r[0] + el.A[3 + a] * omega_r[1] + el.A[6 + a] * omega_r[2];
        }
        if (pf) {
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    w[a * 3 + c] = (el.B[a] * el.A[c] + el.B[6 + a] * el.A[3 + c] + el.B[12 + a] * el.A[6 + c]) * om;
        }
    }
#pragma unroll
    for (int i = 0; i < 9; i++) d.Hll[(size_t)l * 9 + i] = H[i];
#pragma unroll
    for (int i = 0; i < 3; i++) d.bl[(size_t)l * 3 + i] = b[i];
}

// ---- K10b: one warp per keyframe: Hpp and bp over the keyframe's edges (fixed lane order => deterministic sums)
__global__ void __launch_bounds__(256) k_ba_build_pose(BaDev d) {
    const int pi = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (pi >= d.P) return;
    const int pf = d.pfree[pi];
    if (pf < 0) return;
    const int lane = threadIdx.x & 31;
    double H[21], b[6];
#pragma unroll
    for (int i = 0; i < 21; i++) H[i] = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) b[i] = 0;
    for (int k = d.pose_off[pi] + lane; k < d.pose_off[pi + 1]; k += 32) {
        const int e = d.pose_edges[k];
        const int li = d.e_point[e];
        EdgeLin el;
        edge_eval<true>(d.q + 4 * pi, d.t + 3 * pi, d.cam + 5 * pi, d.X + 3 * li, d.e_obs + 3 * e, el);
        double om = d.e_info[e], omega_r[3];
        robust_weights(d, el, om, omega_r);
        int idx = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int c = 0; c <= a; c++) H[idx++] += (el.B[a] * el.B[c] + el.B[6 + a] * el.B[6 + c] + el.B[12 + a] * el.B[12 + c]) * om;
            b[a] += el.B[a] * omega_r[0] + el.B[6 + a] * omega_r[1] + el.B[12 + a] * omega_r[2];
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 21; i++) H[i] += __shfl_down_sync(0xffffffffu, H[i], o);
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] += __shfl_down_sync(0xffffffffu, b[i], o);
    }
    if (lane == 0) {
        double* Ho = d.Hpp + (size_t)pf * 36;
        int idx = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = 0; c <= a; c++) {
                Ho[a * 6 + c] = H[idx];
                Ho[c * 6 + a] = H[idx];
                idx++;
            }
#pragma unroll
        for (int a = 0; a < 6; a++) d.bp[(size_t)pf * 6 + a] = b[a];
    }
}

__global__ void __launch_bounds__(256) k_ba_copy_diag(BaDev d) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < d.Pf * 6) d.xp[i] = d.Hpp[(size_t)(i / 6) * 36 + (i % 6) * 7];
}

// ---- max |diag| of Hpp (after an optional all-reduce) and Hll, for computeLambdaInit
__global__ void __launch_bounds__(256) k_ba_maxdiag(BaDev d, int which) {
    __shared__ double sm[32];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double v = 0;
    if (which == 0) {
        if (i < d.Pf * 6) v = fabs(d.xp[i]);
    } else {
        if (i < d.L * 3 && d.lfree[i / 3] >= 0) v = fabs(d.Hll[(size_t)(i / 3) * 9 + (i % 3) * 4]);
    }
    v = block_reduce<true>(v, sm);
    if (threadIdx.x == 0) d.partial[blockIdx.x] = v;
}

// ---- K11a: Dinv = (Hll + lambda I)^-1 and Dinv * bl per landmark
__global__ void __launch_bounds__(256) k_ba_dinv(BaDev d, double lambda) {
    const int l = blockIdx.x * 256 + threadIdx.x;
    if (l >= d.L || d.lfree[l] < 0) return;
    double a[9];
#pragma unroll
    for (int i = 0; i < 9; i++) a[i] = d.Hll[(size_t)l * 9 + i];
    a[0] += lambda; a[4] += lambda; a[8] += lambda;
    const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    const double id = 1.0 / (a[0] * c0 + a[1] * c1 + a[2] * c2);
    double o[9];
    o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
#pragma unroll
    for (int i = 0; i < 9; i++) d.Dinv[(size_t)l * 9 + i] = o[i];
    const double* b = d.bl + (size_t)l * 3;
#pragma unroll
    for (int i = 0; i < 3; i++) d.db[(size_t)l * 3 + i] = o[i * 3] * b[0] + o[i * 3 + 1] * b[1] + o[i * 3 + 2] * b[2];
}

// ---- K11b: one warp per keyframe = one block row of the reduced system:
//      S(j, j') = Hpp(j) [j'=j] - sum_l W_jl Dinv_l W_j'l^T  for j' <= j,   bs(j) = bp(j) - sum_l W_jl Dinv_l bl
//      lane = entry (r,c) of the 6x6 block (lanes 0..3 also own entries 32..35); the warp walks its edges in order.
//      One CTA (8 warps) per keyframe. A row touches few distinct blocks (its own keyframe and the co-observers of its
//      landmarks) but each of them hundreds of times, and every edge costs a chain of dependent index loads
//      (edge -> landmark -> Dinv, the landmark's edges -> their poses -> W): one warp per row left that latency exposed
//      for ~500 edges in a row (2 ms per launch, set by the slowest row). Now warp w takes the edges k = w, w + 8, ...,
//      accumulates into a private block cache in shared memory (tag = block column), and the caches are merged in warp
//      order - a fixed order per entry, so the sums are deterministic. A row with more distinct blocks than slots
//      (dense covisibility) is redone by warp 0 alone, overflow blocks accumulating in global memory (the previous
//      algorithm); the host has zeroed S (cudaMemsetAsync).
constexpr int kAccSlots = 16, kBaseSlots = 32;

// the edges pose_off[pi] + w0, + step, ... of keyframe pi into the cache (acc, tags) of the calling warp; blocks that do not
// fit go to `row` in global memory when allowed (single-warp use only), else *overflow is raised and the term is dropped
__device__ __forceinline__ void schur_accumulate(const BaDev& d, int pi, int j, int w0, int step, double (*acc)[36], int* tags, int& nslots,
                                                 double& coeff, double* row, int fj, bool global_ok, int* overflow) {
    // Every edge costs a chain of dependent index loads (pose_edges -> e_point -> lfree / lm_off, then e_pose -> pfree per
    // co-observing keyframe). The chains of 32 edges are walked by the 32 lanes at once and handed round with shuffles; the
    // co-observers' indices of a landmark are fetched the same way, and the landmark's W blocks (contiguous: edges are
    // grouped by landmark) are prefetched while they resolve. A warp takes batches of 32 consecutive edges of the row
    // (batch w0, w0 + step, ...), inside a batch in order: a fixed summation order.
    const int lane = threadIdx.x & 31;
    const int r0 = lane / 6, c0 = lane - r0 * 6;  // entry `lane`
    const int c1 = 2 + lane;                      // entry 32 + lane = (5, 2 + lane) for lane < 4
    const int k_end = d.pose_off[pi + 1];
    for (int kb = d.pose_off[pi] + 32 * w0; kb < k_end; kb += 32 * step) {
        int e_l = -1, l_l = 0, o0_l = 0, o1_l = 0;
        if (kb + lane < k_end) {
            e_l = d.pose_edges[kb + lane];
            l_l = d.e_point[e_l];
            if (d.lfree[l_l] < 0) {
                e_l = -1;
            } else {
                o0_l = d.lm_off[l_l];
                o1_l = d.lm_off[l_l + 1];
            }
        }
        const int cnt = min(32, k_end - kb);
        for (int i = 0; i < cnt; i++) {
            const int e = __shfl_sync(0xffffffffu, e_l, i);
            if (e < 0) continue;
            const int l = __shfl_sync(0xffffffffu, l_l, i), p0 = __shfl_sync(0xffffffffu, o0_l, i), p1 = __shfl_sync(0xffffffffu, o1_l, i);
            {   // the landmark's W blocks, 144 B per edge: one 128-byte line per lane
                const char* wb = reinterpret_cast<const char*>(d.W + (size_t)p0 * 18);
                const size_t bytes = (size_t)(p1 - p0) * 144;
                if ((size_t)lane * 128 < bytes) asm volatile("prefetch.global.L1 [%0];" ::"l"(wb + (size_t)lane * 128));
            }
            int j2_l = -1;
            if (p0 + lane < p1) j2_l = d.pfree[d.e_pose[p0 + lane]];
            const double* Di = d.Dinv + (size_t)l * 9;
            const double* We = d.W + (size_t)e * 18;
            double bd0[3], bd1[3];
#pragma unroll
            for (int m = 0; m < 3; m++) {
                bd0[m] = We[r0 * 3] * Di[m] + We[r0 * 3 + 1] * Di[3 + m] + We[r0 * 3 + 2] * Di[6 + m];
                bd1[m] = We[15] * Di[m] + We[16] * Di[3 + m] + We[17] * Di[6 + m];
            }
            if (lane < 6) {
                const double* dbl = d.db + (size_t)l * 3;
                coeff += We[lane * 3] * dbl[0] + We[lane * 3 + 1] * dbl[1] + We[lane * 3 + 2] * dbl[2];
            }
            for (int cb = p0; cb < p1; cb += 32) {
                if (cb != p0) j2_l = cb + lane < p1 ? d.pfree[d.e_pose[cb + lane]] : -1;
                const int m2 = min(32, p1 - cb);
                for (int t = 0; t < m2; t++) {
                    const int j2 = __shfl_sync(0xffffffffu, j2_l, t);
                    if (j2 < 0 || j2 > j) continue;
                    const double* W2 = d.W + (size_t)(cb + t) * 18;
                    int slot = -1;  // warp-uniform lookup (j2 is the same for every lane)
                    const int tg = lane < kAccSlots ? tags[lane] : -2;
                    const unsigned hit = __ballot_sync(0xffffffffu, tg == j2);
                    if (hit) {
                        slot = __ffs(hit) - 1;
                    } else if (nslots < kAccSlots) {
                        slot = nslots++;
                        if (lane == 0) tags[slot] = j2;
                        __syncwarp();
                    }
                    double* blk;
                    if (slot >= 0) blk = acc[slot];
                    else if (global_ok) blk = row + (size_t)(j2 - fj) * 36;
                    else { if (lane == 0) *overflow = 1; continue; }
                    blk[lane] -= bd0[0] * W2[c0 * 3] + bd0[1] * W2[c0 * 3 + 1] + bd0[2] * W2[c0 * 3 + 2];
                    if (lane < 4) blk[32 + lane] -= bd1[0] * W2[c1 * 3] + bd1[1] * W2[c1 * 3 + 1] + bd1[2] * W2[c1 * 3 + 2];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_ba_schur_rows(BaDev d) {
    __shared__ double acc[8][kAccSlots][36];
    __shared__ int tags[8][kAccSlots];
    __shared__ int nsl[8];
    __shared__ double coeffs[8][6];
    __shared__ double base[kBaseSlots][36];
    __shared__ int btag[kBaseSlots];
    __shared__ int nbase, s_over;
    const int pi = blockIdx.x;
    const int j = d.pfree[pi];
    if (j < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int fj = d.first[j];
    double* row = d.S + (size_t)d.rowoff[j] * 36;
    for (int i = tid; i < 8 * kAccSlots * 36; i += 256) (&acc[0][0][0])[i] = 0.0;
    for (int i = tid; i < kBaseSlots * 36; i += 256) (&base[0][0])[i] = 0.0;
    if (tid < 8 * kAccSlots) (&tags[0][0])[tid] = -1;
    if (tid == 0) { s_over = 0; nbase = 0; }
    __syncthreads();
    {
        int nslots = 0;
        double coeff = 0.0;
        schur_accumulate(d, pi, j, w, 8, acc[w], tags[w], nslots, coeff, row, fj, false, &s_over);
        if (lane == 0) nsl[w] = nslots;
        if (lane < 6) coeffs[w][lane] = coeff;
    }
    __syncthreads();
    if (!s_over) {
        // merge the caches in warp order: a tag occurs once per warp table, so inside a round every base slot has one writer
        for (int ww = 0; ww < 8; ww++) {
            const int nb0 = nbase;
            for (int sl = w; sl < nsl[ww]; sl += 8) {
                const int tg = tags[ww][sl];
                int bi = -1;
                if (lane == 0) {
                    for (int i = 0; i < nb0; i++) if (btag[i] == tg) { bi = i; break; }
                    if (bi < 0) {
                        bi = atomicAdd(&nbase, 1);
                        if (bi < kBaseSlots) btag[bi] = tg; else s_over = 1;
                    }
                }
                bi = __shfl_sync(0xffffffffu, bi, 0);
                if (bi >= kBaseSlots) continue;
                base[bi][lane] += acc[ww][sl][lane];
                if (lane < 4) base[bi][32 + lane] += acc[ww][sl][32 + lane];
            }
            __syncthreads();
        }
    }
    if (!s_over) {
        const int nb = nbase;
        bool have_diag = false;
        for (int bi = w; bi < nb; bi += 8) {  // every block of the row is written exactly once
            const int j2 = btag[bi];
            double* blk = row + (size_t)(j2 - fj) * 36;
            const double h0 = j2 == j ? d.Hpp[(size_t)j * 36 + lane] : 0.0;
            blk[lane] = h0 + base[bi][lane];
            if (lane < 4) blk[32 + lane] = (j2 == j ? d.Hpp[(size_t)j * 36 + 32 + lane] : 0.0) + base[bi][32 + lane];
        }
        for (int bi = 0; bi < nb; bi++) have_diag |= btag[bi] == j;
        if (!have_diag && w == 0) {  // a keyframe without any free landmark: the diagonal is Hpp alone
            double* blk = row + (size_t)(j - fj) * 36;
            blk[lane] = d.Hpp[(size_t)j * 36 + lane];
            if (lane < 4) blk[32 + lane] = d.Hpp[(size_t)j * 36 + 32 + lane];
        }
        if (tid < 6) {
            double c = 0.0;
            for (int ww = 0; ww < 8; ww++) c += coeffs[ww][tid];
            d.bs[(size_t)j * 6 + tid] = d.bp[(size_t)j * 6 + tid] - c;
        }
        return;
    }
    // ---- dense covisibility: warp 0 alone, diagonal in slot 0 starting from Hpp, overflow blocks in global memory
    if (w != 0) return;
    for (int i = lane; i < kAccSlots * 36; i += 32) (&acc[0][0][0])[i] = 0.0;
    if (lane < kAccSlots) tags[0][lane] = lane == 0 ? j : -1;
    __syncwarp();
    acc[0][0][lane] = d.Hpp[(size_t)j * 36 + lane];
    if (lane < 4) acc[0][0][32 + lane] = d.Hpp[(size_t)j * 36 + 32 + lane];
    __syncwarp();
    int nslots = 1;
    double coeff = 0.0;
    schur_accumulate(d, pi, j, 0, 1, acc[0], tags[0], nslots, coeff, row, fj, true, &s_over);
    __syncwarp();
    for (int sl = 0; sl < nslots; sl++) {
        double* blk = row + (size_t)(tags[0][sl] - fj) * 36;
        blk[lane] = acc[0][sl][lane];
        if (lane < 4) blk[32 + lane] = acc[0][sl][32 + lane];
    }
    if (lane < 6) d.bs[(size_t)j * 6 + lane] = d.bp[(size_t)j * 6 + lane] - coeff;
}

// ---- K11c: the same block rows from SORTED PAIR LISTS (default; k_ba_schur_rows above stays as the fallback and the A/B arm).
//      Profile of K11b at the bench size (profiles/r02_ba_schur_rows_ncu.txt): 319 M warp instructions for 2.8 M block products -
//      117 per edge for the index chains, 61 per (edge, co-observer) visit for six useful DFMA: instruction bound, not bandwidth.
//      The structure (which edge of keyframe j meets which edge of keyframe j2 <= j in a landmark) does not change between LM
//      iterations, so it is enumerated ONCE per BA: per keyframe the list of (own edge k, co-observer edge p) pairs sorted by
//      (block column j2, k), cut into items of <= 32 consecutive pairs of one column. An iteration then is
//        phase 1  T_k = W_k Dinv_l for the row's edges into shared memory (thread per edge) and the ordered sum for bs
//        phase 2  nine lanes per item (lane = a 2 x 2 piece of the 6 x 6 block): acc += T_k W_p^T over the item's pairs in order,
//                 three 128-bit loads from shared memory and three from L2 per 12 DFMA; a column with one item is written to S,
//                 the others leave their partial sums in shared-memory slots
//        phase 3  columns with several items add their slots in item order
//      - every block has one writer and a fixed summation order (deterministic, no floating-point atomics).
struct SchurPairs {
    const int* row_pair_off;   // [P + 1] pairs of keyframe pi: [row_pair_off[pi], row_pair_off[pi + 1])
    const int* row_item_off;   // [P + 1] first item slot of keyframe pi (offsets by upper bound)
    int* row_item_cnt;         // [P] items of keyframe pi (a sentinel item with .y = end of the pairs follows them)
    uint2* pairs;              // x = position p of the co-observer's edge, y = local index k of the own edge in the row
    int4* items;               // x = block column j2, y = first pair (absolute), z = meta, w = 0
    int* info;                 // [0] max edges per row, [1] max pairs per row, [2] max slots per row, [3] cap exceeded
    int chunk;                 // pairs per item at most (a power of two: 32 or 64)
    int2* row_el;              // [E] in pose-CSR order: x = edge position e, y = landmark l (-1: fixed landmark) - the row's index chain, resolved once
};
constexpr int kSpThreads = 512, kSpPairCap = 8192;
constexpr int kSpMulti = 1 << 6, kSpFirst = 1 << 7, kSpLast = 1 << 8;

// exclusive scan of one int per thread over a 512-thread block (thread order); *total = block sum
__device__ __forceinline__ int sp_block_scan(int v, int* warp_tmp, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kSpThreads / 32 ? warp_tmp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        warp_tmp[lane] = w;
    }
    __syncthreads();
    const int base = wid ? warp_tmp[wid - 1] : 0;
    *total = warp_tmp[kSpThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// pairs of one edge of row j: co-observers of its landmark in keyframes j2 with 0 <= j2 <= j
__device__ __forceinline__ int sp_edge_pairs(const BaDev& d, int e, int j) {
    const int l = d.e_point[e];
    if (d.lfree[l] < 0) return 0;
    int n = 0;
    for (int p = d.lm_off[l]; p < d.lm_off[l + 1]; p++) {
        const int j2 = d.pfree[d.e_pose[p]];
        n += j2 >= 0 && j2 <= j;
    }
    return n;
}

__global__ void __launch_bounds__(256) k_ba_pairs_count(BaDev d, int* __restrict__ row_np, int* __restrict__ row_bound, int* __restrict__ info, int chunk,
                                                        int2* __restrict__ row_el) {
    __shared__ int sm[8];
    const int pi = blockIdx.x, j = d.pfree[pi];
    const int k0 = d.pose_off[pi], ne = d.pose_off[pi + 1] - k0;
    int n = 0;
    if (j >= 0)
        for (int k = threadIdx.x; k < ne; k += 256) {
            const int e = d.pose_edges[k0 + k], l = d.e_point[e];
            row_el[k0 + k] = make_int2(e, d.lfree[l] < 0 ? -1 : l);
            n += sp_edge_pairs(d, e, j);
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int np = 0;
        for (int w = 0; w < 8; w++) np += sm[w];
        row_np[pi] = np;
        // items <= distinct columns + cuts at multiples of the chunk, + the sentinel
        row_bound[pi] = j < 0 ? 0 : (np + chunk - 1) / chunk + min(np, j - d.first[j] + 1) + 1;
        if (j >= 0) {
            atomicMax(&info[0], ne);
            atomicMax(&info[1], np);
        }
    }
}

// exclusive scans of the two per-row arrays (one block; P is a few thousand keyframes)
__global__ void __launch_bounds__(1024) k_ba_pairs_scan(const int* __restrict__ a, const int* __restrict__ b, int n, int* __restrict__ oa,
                                                        int* __restrict__ ob) {
    __shared__ int wa[32], wb[32], base[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base[0] = base[1] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int va = i < n ? a[i] : 0, vb = i < n ? b[i] : 0;
        int ia = va, ib = vb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ua; ib += ub; }
        }
        if (lane == 31) { wa[wid] = ia; wb[wid] = ib; }
        __syncthreads();
        if (wid == 0) {
            int xa = wa[lane], xb = wb[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ua = __shfl_up_sync(0xffffffffu, xa, o), ub = __shfl_up_sync(0xffffffffu, xb, o);
                if (lane >= o) { xa += ua; xb += ub; }
            }
            wa[lane] = xa; wb[lane] = xb;
        }
        __syncthreads();
        const int ba = base[0] + (wid ? wa[wid - 1] : 0), bb = base[1] + (wid ? wb[wid - 1] : 0);
        if (i < n) { oa[i] = ba + ia - va; ob[i] = bb + ib - vb; }
        __syncthreads();
        if (threadIdx.x == 0) { base[0] += wa[31]; base[1] += wb[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { oa[n] = base[0]; ob[n] = base[1]; }
}

// one CTA per keyframe: enumerate the pairs in edge order, sort by (column, position) in shared memory, emit pairs and items
__global__ void __launch_bounds__(kSpThreads) k_ba_pairs_build(BaDev d, SchurPairs sp) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    uint32_t* keys = reinterpret_cast<uint32_t*>(sp_smem);            // [n2]
    uint2* tmp = reinterpret_cast<uint2*>(sp_smem + (size_t)kSpPairCap * 4);  // [np]
    __shared__ int warp_tmp[32];
    __shared__ int run_base;
    const int pi = blockIdx.x, j = d.pfree[pi], tid = threadIdx.x;
    if (j < 0) { if (tid == 0) sp.row_item_cnt[pi] = 0; return; }
    const int off = sp.row_pair_off[pi], np = sp.row_pair_off[pi + 1] - off;
    const int item0 = sp.row_item_off[pi];
    if (np > kSpPairCap || d.Pf >= (1 << 19)) {
        if (tid == 0) { sp.info[3] = 1; sp.row_item_cnt[pi] = 0; }
        return;
    }
    const int k0 = d.pose_off[pi], ne = d.pose_off[pi + 1] - k0, fj = d.first[j];
    int n2 = 1;
    while (n2 < np) n2 <<= 1;
    if (tid == 0) run_base = 0;
    for (int i = np + tid; i < n2; i += kSpThreads) keys[i] = 0xffffffffu;
    __syncthreads();
    for (int kb = 0; kb < ne; kb += kSpThreads) {  // enumeration in edge order: positions from a block scan
        const int k = kb + tid;
        int e = -1, cnt = 0;
        if (k < ne) { e = d.pose_edges[k0 + k]; cnt = sp_edge_pairs(d, e, j); }
        int total;
        int pos = run_base + sp_block_scan(cnt, warp_tmp, &total);
        if (cnt) {
            const int l = d.e_point[e];
            for (int p = d.lm_off[l]; p < d.lm_off[l + 1]; p++) {
                const int j2 = d.pfree[d.e_pose[p]];
                if (j2 >= 0 && j2 <= j) {
                    keys[pos] = (uint32_t)(j2 - fj) << 13 | (uint32_t)pos;
                    tmp[pos] = make_uint2((unsigned)p, (unsigned)k);
                    pos++;
                }
            }
        }
        __syncthreads();
        if (tid == 0) run_base += total;
        __syncthreads();
    }
    for (int kk = 2; kk <= n2; kk <<= 1)  // bitonic sort, ascending
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < (n2 >> 1); i += kSpThreads) {
                const int lo = ((i & ~(jj - 1)) << 1) | (i & (jj - 1)), hi = lo | jj;
                const bool up = (lo & kk) == 0;
                const uint32_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    // pairs in sorted order; an item starts where the column changes and at every multiple of the chunk
    if (tid == 0) run_base = 0;
    __syncthreads();
    for (int ib = 0; ib < np; ib += kSpThreads) {
        const int i = ib + tid;
        int head = 0, col = 0;
        if (i < np) {
            const uint32_t key = keys[i];
            col = (int)(key >> 13);
            sp.pairs[off + i] = tmp[key & 8191u];
            head = i == 0 || (i & (sp.chunk - 1)) == 0 || (int)(keys[i - 1] >> 13) != col;
        }
        int total;
        const int it = run_base + sp_block_scan(head, warp_tmp, &total);
        if (head) sp.items[item0 + it] = make_int4(col + fj, off + i, 0, 0);
        __syncthreads();
        if (tid == 0) run_base += total;
        __syncthreads();
    }
    const int n_items = run_base;
    if (tid == 0) {
        sp.items[item0 + n_items] = make_int4(-1, off + np, 0, 0);  // sentinel: end of the last item
        sp.row_item_cnt[pi] = n_items;
    }
    __syncthreads();
    // meta: columns with several items keep partial sums in slots (consecutive in item order)
    if (tid == 0) run_base = 0;
    __syncthreads();
    for (int ib = 0; ib < n_items; ib += kSpThreads) {
        const int it = ib + tid;
        int multi = 0, first = 0, last = 0;
        if (it < n_items) {
            const int c = sp.items[item0 + it].x;
            const bool same_prev = it > 0 && sp.items[item0 + it - 1].x == c;
            const bool same_next = it + 1 < n_items && sp.items[item0 + it + 1].x == c;
            multi = same_prev || same_next;
            first = multi && !same_prev;
            last = multi && !same_next;
        }
        int total;
        const int slot = run_base + sp_block_scan(multi, warp_tmp, &total);
        if (it < n_items) sp.items[item0 + it].z = (multi ? kSpMulti : 0) | (first ? kSpFirst : 0) | (last ? kSpLast : 0) | slot << 16;
        __syncthreads();
        if (tid == 0) run_base += total;
        __syncthreads();
    }
    if (tid == 0) atomicMax(&sp.info[2], run_base);
}

template <int TH, int CH>
__global__ void __launch_bounds__(TH, TH == 256 ? 2 : 1) k_ba_schur_pairs(BaDev d, SchurPairs sp, int t_cap) {
    constexpr int NR = (CH + 8) / 9;  // record registers per lane
    extern __shared__ __align__(16) unsigned char sp_smem[];
    double* T = reinterpret_cast<double*>(sp_smem);          // [t_cap][18]
    double* part = T + (size_t)t_cap * 18;                    // [slots][36]
    __shared__ double red[TH / 32][6];
    const int pi = blockIdx.x, j = d.pfree[pi];
    if (j < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = d.pose_off[pi], ne = d.pose_off[pi + 1] - k0;
    // ---- phase 1: T_k = W_k Dinv_l, and bs(j) = bp(j) - sum_k W_k (Dinv_l bl)
    double cf[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 2
    for (int k = tid; k < ne; k += TH) {
        const int2 el = sp.row_el[k0 + k];
        const int e = el.x, l = el.y;
        double2* Tk = reinterpret_cast<double2*>(T + (size_t)k * 18);
        if (l < 0) {
#pragma unroll
            for (int i = 0; i < 9; i++) Tk[i] = make_double2(0.0, 0.0);
            continue;
        }
        double w[18], di[9], db[3];
        const double2* Wg = reinterpret_cast<const double2*>(d.W + (size_t)e * 18);
#pragma unroll
        for (int i = 0; i < 9; i++) { const double2 v = Wg[i]; w[2 * i] = v.x; w[2 * i + 1] = v.y; }
#pragma unroll
        for (int i = 0; i < 9; i++) di[i] = d.Dinv[(size_t)l * 9 + i];
#pragma unroll
        for (int i = 0; i < 3; i++) db[i] = d.db[(size_t)l * 3 + i];
        double t[18];
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
            for (int m = 0; m < 3; m++) t[r * 3 + m] = w[r * 3] * di[m] + w[r * 3 + 1] * di[3 + m] + w[r * 3 + 2] * di[6 + m];
            cf[r] += w[r * 3] * db[0] + w[r * 3 + 1] * db[1] + w[r * 3 + 2] * db[2];
        }
#pragma unroll
        for (int i = 0; i < 9; i++) Tk[i] = make_double2(t[2 * i], t[2 * i + 1]);
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) cf[r] += __shfl_down_sync(0xffffffffu, cf[r], o);
        if (lane == 0) red[warp][r] = cf[r];
    }
    __syncthreads();
    if (tid < 6) {
        double c = 0.0;
        for (int w = 0; w < TH / 32; w++) c += red[w][tid];
        d.bs[(size_t)j * 6 + tid] = d.bp[(size_t)j * 6 + tid] - c;
    }
    // ---- phase 2: items
    const int item0 = sp.row_item_off[pi], n_items = sp.row_item_cnt[pi];
    const int fj = d.first[j];
    double* row = d.S + (size_t)d.rowoff[j] * 36;
    const int g = warp * 3 + lane / 9, gl = lane % 9, rr = gl / 3, cc = gl - rr * 3;
    const bool act = lane < 27;
    const double* Hj = d.Hpp + (size_t)j * 36;
    // Every warp takes three items per round (one per group of nine lanes). The item's <= 32 pair records are fetched by the
    // group's lanes in one go (record i sits in register i / 9 of lane i % 9) and handed round with shuffles, so the W blocks
    // of six pairs (18 independent 128-bit loads per lane) are in flight together: one L2 round trip per six block products.
    const int gbase = (lane / 9) * 9;
    for (int itb = 0; itb < n_items; itb += (TH / 32) * 3) {
        const int it = itb + g;
        const bool valid = act && it < n_items;
        int4 I = make_int4(0, 0, 0, 0);
        int cnt = 0;
        if (valid) {
            I = sp.items[item0 + it];
            cnt = sp.items[item0 + it + 1].y - I.y;
        }
        uint2 R[NR];
#pragma unroll
        for (int q = 0; q < NR; q++) {
            const int i = gl + 9 * q;
            R[q] = valid && i < cnt ? sp.pairs[I.y + i] : make_uint2(0u, 0u);
        }
        int cmax = cnt;
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, 16));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, 8));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, 4));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, 2));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, 1));
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll
        for (int b6 = 0; b6 < NR * 9; b6 += 6) {
            if (b6 >= cmax) break;  // warp-uniform
            unsigned pp[6], kk[6];
#pragma unroll
            for (int u = 0; u < 6; u++) {
                const int i = b6 + u;  // compile-time: register i / 9, source lane i % 9 of the group
                const int q = i / 9 < NR ? i / 9 : NR - 1;
                pp[u] = __shfl_sync(0xffffffffu, R[q].x, gbase + i % 9);
                kk[u] = __shfl_sync(0xffffffffu, R[q].y, gbase + i % 9);
            }
            double2 w0[6], w1[6], w2[6];
#pragma unroll
            for (int u = 0; u < 6; u++) {
                const double2* Wp = reinterpret_cast<const double2*>(d.W + (size_t)pp[u] * 18 + cc * 6);
                if (b6 + u < cnt) { w0[u] = __ldg(Wp); w1[u] = __ldg(Wp + 1); w2[u] = __ldg(Wp + 2); }
            }
#pragma unroll
            for (int u = 0; u < 6; u++) {
                if (b6 + u < cnt) {
                    const double2* Tp = reinterpret_cast<const double2*>(T + (size_t)kk[u] * 18 + rr * 6);
                    const double2 t0 = Tp[0], t1 = Tp[1], t2 = Tp[2];
                    // T rows 2rr: (t0.x, t0.y, t1.x), 2rr + 1: (t1.y, t2.x, t2.y); W rows 2cc: (w0.x, w0.y, w1.x), 2cc + 1: (w1.y, w2.x, w2.y)
                    a00 += t0.x * w0[u].x + t0.y * w0[u].y + t1.x * w1[u].x;
                    a01 += t0.x * w1[u].y + t0.y * w2[u].x + t1.x * w2[u].y;
                    a10 += t1.y * w0[u].x + t2.x * w0[u].y + t2.y * w1[u].x;
                    a11 += t1.y * w1[u].y + t2.x * w2[u].x + t2.y * w2[u].y;
                }
            }
        }
        if (valid) {
            const int e0 = (2 * rr) * 6 + 2 * cc;  // entries (2rr, 2cc), (2rr, 2cc+1), (2rr+1, 2cc), (2rr+1, 2cc+1)
            if (I.z & kSpMulti) {
                double* ps = part + (size_t)(I.z >> 16) * 36;
                ps[e0] = a00; ps[e0 + 1] = a01; ps[e0 + 6] = a10; ps[e0 + 7] = a11;
            } else {
                double* blk = row + (size_t)(I.x - fj) * 36;
                const bool dg = I.x == j;
                blk[e0] = (dg ? Hj[e0] : 0.0) - a00;
                blk[e0 + 1] = (dg ? Hj[e0 + 1] : 0.0) - a01;
                blk[e0 + 6] = (dg ? Hj[e0 + 6] : 0.0) - a10;
                blk[e0 + 7] = (dg ? Hj[e0 + 7] : 0.0) - a11;
            }
        }
    }
    __syncthreads();
    // ---- phase 3: columns with several items: their slots in item order
    if (act)
        for (int it = g; it < n_items; it += (TH / 32) * 3) {
            const int4 I = sp.items[item0 + it];
            if (!(I.z & kSpFirst)) continue;
            const int e0 = (2 * rr) * 6 + 2 * cc;
            double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
            int z = I.z;
            for (int q = it;; q++) {
                const double* ps = part + (size_t)(z >> 16) * 36;
                a00 += ps[e0]; a01 += ps[e0 + 1]; a10 += ps[e0 + 6]; a11 += ps[e0 + 7];
                if (z & kSpLast) break;
                z = sp.items[item0 + q + 1].z;
            }
            double* blk = row + (size_t)(I.x - fj) * 36;
            const bool dg = I.x == j;
            blk[e0] = (dg ? Hj[e0] : 0.0) - a00;
            blk[e0 + 1] = (dg ? Hj[e0 + 1] : 0.0) - a01;
            blk[e0 + 6] = (dg ? Hj[e0 + 6] : 0.0) - a10;
            blk[e0 + 7] = (dg ? Hj[e0 + 7] : 0.0) - a11;
        }
    // a keyframe without any free landmark: the diagonal is Hpp alone (the diagonal column sorts last)
    if (n_items == 0 || sp.items[item0 + n_items - 1].x != j) {
        if (tid < 36) row[(size_t)(j - fj) * 36 + tid] = Hj[tid];
    }
}

__global__ void __launch_bounds__(256) k_ba_add_lambda(BaDev d, double lambda) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= d.Pf * 6) return;
    const int j = i / 6, a = i - j * 6;
    d.S[(size_t)(d.rowoff[j] + j - d.first[j]) * 36 + a * 7] += lambda;
}

// ---- K12: block-skyline Cholesky + forward/backward substitution, one CTA. Row j holds block columns first[j]..j;
//      col_rows[coloff[k]..coloff[k+1]) lists the rows j > k whose envelope contains column k (ascending).
//
//      Right-looking with one-column look-ahead, two barriers per column:
//        B: every active row solves its block against L_kk (L_jk = A_jk L_kk^-T) and stages it in shared memory;
//           one thread finishes y_k = L_kk^-1 b_k (the forward substitution rides along with the factorisation)
//        C: warp 0 updates and factors the next diagonal block while the other warps apply the trailing update
//           A_ji -= L_jk L_ik^T over all active pairs from shared memory (4 pairs in flight per warp to cover the
//           L2 round trip of the read-modify-write) and b_j -= L_jk y_k
#ifndef CORB_BA_THREADS
#define CORB_BA_THREADS 1024
#endif
constexpr int kSolveThreads = CORB_BA_THREADS;
constexpr int kSolveWarps = kSolveThreads / 32;

__device__ __forceinline__ bool chol6(double* a) {  // in place, lower; upper part zeroed
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        double s = a[c * 6 + c];
#pragma unroll
        for (int p = 0; p < c; p++) s -= a[c * 6 + p] * a[c * 6 + p];
        if (!(s > 0.0)) { ok = false; s = 1.0; }
        const double dd = sqrt(s);
        a[c * 6 + c] = dd;
        const double inv = 1.0 / dd;
#pragma unroll
        for (int r = c + 1; r < 6; r++) {
            double v = a[r * 6 + c];
#pragma unroll
            for (int p = 0; p < c; p++) v -= a[r * 6 + p] * a[c * 6 + p];
            a[r * 6 + c] = v * inv;
        }
#pragma unroll
        for (int r = 0; r < c; r++) a[r * 6 + c] = 0.0;
    }
    return ok;
}

// Warp-cooperative update + Cholesky of one 6x6 diagonal block: lane t < 21 owns entry (r, c), r >= c, of the lower
// triangle in a register; columns are finalised with shuffles (no local-memory arrays on the sequential critical path).
// D <- chol(D - L0 L0^T) (L0 == nullptr: no update), result also written to Lout (full 6x6, upper part zero).
// `a_pre` (optional) = this lane's entry of D, loaded earlier to take the L2 round trip off the sequential path.
// The reciprocals 1 / L_kk go to inv_out[6] (and inv_glob): every triangular solve against this block multiplies by
// them instead of dividing (fp64 division is a ~250-cycle software sequence on the critical path of each column).
// 1 / sqrt(a) is rsqrt(a) (1 ulp) and L_kk = a * rsqrt(a): two roundings instead of sqrt's one, far below the 1e-5 px
// tolerance of the BA parity tests.
__device__ __forceinline__ bool warp_chol6(double* D, const double* L0, double* Lout, int lane, double* inv_out, double* inv_glob,
                                           bool have_pre = false, double a_pre = 0.0) {
    int r = 0, c = 0;
    if (lane < 21) {
        r = lane >= 15 ? 5 : lane >= 10 ? 4 : lane >= 6 ? 3 : lane >= 3 ? 2 : lane >= 1 ? 1 : 0;
        c = lane - r * (r + 1) / 2;
    }
    double a = have_pre ? a_pre : (lane < 21 ? D[r * 6 + c] : 0.0);
    if (L0 && lane < 21) {
        double s = 0;
#pragma unroll
        for (int p = 0; p < 6; p++) s += L0[r * 6 + p] * L0[c * 6 + p];
        a -= s;
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double akk = __shfl_sync(0xffffffffu, a, k * (k + 1) / 2 + k);
        if (!(akk > 0.0)) { ok = false; akk = 1.0; }
        const double inv = rsqrt(akk), dk = akk * inv;
        if (lane == 0) { inv_out[k] = inv; inv_glob[k] = inv; }
        if (lane < 21 && c == k) a = (r == k) ? dk : a * inv;
        const double lrk = __shfl_sync(0xffffffffu, a, r * (r + 1) / 2 + min(k, r));
        const double lck = __shfl_sync(0xffffffffu, a, c * (c + 1) / 2 + min(k, c));
        if (lane < 21 && c > k) a -= lrk * lck;
    }
    if (lane < 21) {
        D[r * 6 + c] = a;
        Lout[r * 6 + c] = a;
        if (r != c) { D[c * 6 + r] = 0.0; Lout[c * 6 + r] = 0.0; }
    }
    return ok;
}

__global__ void __launch_bounds__(kSolveThreads) k_ba_solve(BaDev d, int lact_cap, int k_begin, int k_end, int n_band, int flags,
                                                             int stage_nnz) {
    // flags: 1 = first launch (load b, clear the failure flag), 2 = defer border x border updates to k_ba_border_syrk,
    //        4 = run the backward substitution after the last column, 8 = band columns see band rows only (the border rows
    //        are swept by k_ba_border_rows), 16 = backward over the border rows only (j >= n_band), 32 = no factorisation,
    //        backward over the band rows only (j < n_band)
    extern __shared__ double Lact[];  // [lact_cap][36] staged L_jk of the active rows of the current column
    __shared__ double Lkk[2][36];
    __shared__ double Lki[2][6];  // reciprocals of the diagonal of L_kk
    __shared__ double yk[6];
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = d.Pf;
    // The envelope index arrays sit on the address path of every block access (column list -> row -> row offset ->
    // block): read from global memory that is three dependent L2 round trips per column before the first operand
    // arrives. When they fit (stage_nnz >= 0) they are copied behind Lact once.
    const int* firstp = d.first;
    const int* rowoffp = d.rowoff;
    const int* coloffp = (flags & 8) ? d.coloff_b : d.coloff;
    const int* colrowsp = (flags & 8) ? d.col_rows_b : d.col_rows;
    if (stage_nnz >= 0) {
        int* sidx = reinterpret_cast<int*>(Lact + (size_t)lact_cap * 36);
        int* sfirst = sidx, *srowoff = sidx + n, *scoloff = sidx + 2 * n, *scolrows = sidx + 3 * n + 1;
        for (int i = tid; i < n; i += kSolveThreads) { sfirst[i] = firstp[i]; srowoff[i] = rowoffp[i]; }
        for (int i = tid; i <= n; i += kSolveThreads) scoloff[i] = coloffp[i];
        for (int i = tid; i < stage_nnz; i += kSolveThreads) scolrows[i] = colrowsp[i];
        firstp = sfirst; rowoffp = srowoff; coloffp = scoloff; colrowsp = scolrows;
    }
    if (flags & 64) {  // one CTA per independent band chunk
        k_begin = d.chunk_start[blockIdx.x];
        k_end = d.chunk_start[blockIdx.x + 1];
    }
    if (tid == 0) s_fail = (flags & 1) ? 0 : (d.scalars[4] != 0.0);
    if (flags & 1) {
        if (flags & 64) {
            for (int i = k_begin * 6 + tid; i < k_end * 6; i += kSolveThreads) d.xp[i] = d.bs[i];
            if (blockIdx.x == 0)
                for (int i = n_band * 6 + tid; i < n * 6; i += kSolveThreads) d.xp[i] = d.bs[i];
        } else {
            for (int i = tid; i < n * 6; i += kSolveThreads) d.xp[i] = d.bs[i];
        }
    }
    __syncthreads();
    if (warp == 0 && k_begin < k_end && !s_fail && !(flags & 32)) {
        double* D = d.S + (size_t)(rowoffp[k_begin] + k_begin - firstp[k_begin]) * 36;
        if (!warp_chol6(D, nullptr, Lkk[k_begin & 1], lane, Lki[k_begin & 1], d.invd + (size_t)k_begin * 6) && lane == 0) s_fail = 1;
    }
    __syncthreads();
#ifdef CORB_BA_TRACE
    long long trB = 0, trC = 0, trT = clock64(), nact_sum = 0, trBw = 0, trChol = 0;
#endif
    for (int k = k_begin; k < ((flags & 32) ? k_begin : k_end); k++) {
        if (s_fail) break;
#ifdef CORB_BA_TRACE
        const long long trs = clock64();
#endif
        const double* Lk = Lkk[k & 1];
        const double* Li = Lki[k & 1];
        const int cb = coloffp[k], nact = coloffp[k + 1] - cb;
        const int* rows = colrowsp + cb;
        const bool staged = nact <= lact_cap;
        // ---- phase B (warp 0 has no rows here when there are other warps: it fetches its entry of the next diagonal block)
        double dpre = 0.0;
        if (warp == 0 && k + 1 < k_end && lane < 21) {
            const int rr = lane >= 15 ? 5 : lane >= 10 ? 4 : lane >= 6 ? 3 : lane >= 3 ? 2 : lane >= 1 ? 1 : 0;
            const int cc = lane - rr * (rr + 1) / 2;
            dpre = d.S[(size_t)(rowoffp[k + 1] + (k + 1) - firstp[k + 1]) * 36 + rr * 6 + cc];
        }
        for (int it = tid; it < nact * 6; it += kSolveThreads) {
            const int a = it / 6, r = it - a * 6;
            const int j = rows[a];
            double* B = d.S + (size_t)(rowoffp[j] + k - firstp[j]) * 36 + r * 6;
            double v[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double s = B[c];
#pragma unroll
                for (int p = 0; p < c; p++) s -= v[p] * Lk[c * 6 + p];
                v[c] = s * Li[c];
            }
#pragma unroll
            for (int c = 0; c < 6; c++) B[c] = v[c];
            if (staged) {
#pragma unroll
                for (int c = 0; c < 6; c++) Lact[a * 36 + r * 6 + c] = v[c];
            }
        }
        if (tid == kSolveThreads - 1) {  // y_k = L_kk^-1 b_k
            double v[6];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double s = d.xp[k * 6 + r];
#pragma unroll
                for (int p = 0; p < r; p++) s -= Lk[r * 6 + p] * v[p];
                v[r] = s * Li[r];
            }
#pragma unroll
            for (int r = 0; r < 6; r++) { d.xp[k * 6 + r] = v[r]; yk[r] = v[r]; }
        }
#ifdef CORB_BA_TRACE
        const long long trb1 = clock64();
#endif
        __syncthreads();
#ifdef CORB_BA_TRACE
        const long long trm = clock64();
        trB += trb1 - trs; trBw += trm - trb1; nact_sum += nact;
#endif
        // ---- phase C
        const bool next_active = nact > 0 && rows[0] == k + 1;  // rows ascend, so k+1 can only be the first entry
        if (warp == 0) {
            if (k + 1 < k_end) {
                double* D = d.S + (size_t)(rowoffp[k + 1] + (k + 1) - firstp[k + 1]) * 36;
                const double* L0 = !next_active ? nullptr : staged ? Lact : d.S + (size_t)(rowoffp[k + 1] + k - firstp[k + 1]) * 36;
                if (!warp_chol6(D, L0, Lkk[(k + 1) & 1], lane, Lki[(k + 1) & 1], d.invd + (size_t)(k + 1) * 6, true, dpre) && lane == 0)
                    s_fail = 1;
            }
#ifdef CORB_BA_TRACE
            trChol += clock64() - trm;
#endif
        } else if (warp == 1) {  // b_j -= L_jk y_k
            for (int it = lane; it < nact * 6; it += 32) {
                const int a = it / 6, r = it - a * 6;
                const int j = rows[a];
                const double* Lr = staged ? Lact + a * 36 + r * 6 : d.S + (size_t)(rowoffp[j] + k - firstp[j]) * 36 + r * 6;
                double s = 0;
#pragma unroll
                for (int c = 0; c < 6; c++) s += Lr[c] * yk[c];
                d.xp[j * 6 + r] -= s;
            }
        } else {
            // pairs (a >= b) of active rows, linear index p = a (a + 1) / 2 + b; pair 0 = (0, 0) is warp 0's when next_active
            const int npairs = nact * (nact + 1) / 2;
            const int p_begin = next_active ? 1 : 0;
            const bool defer = (flags & 2) != 0;  // pairs of two border rows are applied later by k_ba_border_syrk
            const int r0 = lane / 6, c0 = lane - r0 * 6, c1 = 2 + lane;  // entries `lane` and `32 + lane` (lane < 4)
            constexpr int U = 4;
            for (int base = p_begin + (warp - 2) * U; base < npairs; base += (kSolveWarps - 2) * U) {
                int pa[U], pb[U], toff[U];
                double t0[U], t1[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int p = base + u;
                    toff[u] = -1;
                    if (p < npairs) {
                        int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
                        while (a * (a + 1) / 2 > p) a--;
                        while ((a + 1) * (a + 2) / 2 <= p) a++;
                        const int b = p - a * (a + 1) / 2;
                        const int j = rows[a];
                        if (defer && rows[b] >= n_band) continue;  // rows ascend: b is the smaller row of the pair
                        pa[u] = a;
                        pb[u] = b;
                        toff[u] = (rowoffp[j] + rows[b] - firstp[j]) * 36;
                        t0[u] = d.S[(size_t)toff[u] + lane];
                        t1[u] = lane < 4 ? d.S[(size_t)toff[u] + 32 + lane] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (toff[u] >= 0) {
                        const int ja = rows[pa[u]], jb = rows[pb[u]];
                        const double* La = staged ? Lact + pa[u] * 36 : d.S + (size_t)(rowoffp[ja] + k - firstp[ja]) * 36;
                        const double* Lb = staged ? Lact + pb[u] * 36 : d.S + (size_t)(rowoffp[jb] + k - firstp[jb]) * 36;
                        double s0 = 0, s1 = 0;
#pragma unroll
                        for (int q = 0; q < 6; q++) s0 += La[r0 * 6 + q] * Lb[c0 * 6 + q];
                        d.S[(size_t)toff[u] + lane] = t0[u] - s0;
                        if (lane < 4) {
#pragma unroll
                            for (int q = 0; q < 6; q++) s1 += La[30 + q] * Lb[c1 * 6 + q];
                            d.S[(size_t)toff[u] + 32 + lane] = t1[u] - s1;
                        }
                    }
                }
            }
        }
        __syncthreads();
#ifdef CORB_BA_TRACE
        trC += clock64() - trm;
#endif
    }
#ifdef CORB_BA_TRACE
    if (tid == 0 && k_end > k_begin)
        printf("k_ba_solve cols %d..%d: phase B work %lld + barrier wait %lld cyc/col, phase C %lld (chol %lld) cyc/col, mean nact %.1f, total %lld cyc\n", k_begin, k_end,
               trB / (k_end - k_begin), trBw / (k_end - k_begin), trC / (k_end - k_begin), trChol / (k_end - k_begin), (double)nact_sum / (k_end - k_begin), clock64() - trT);
    const long long trb0 = clock64();
#endif
    if (s_fail) {
        for (int i = tid; i < n * 6; i += kSolveThreads) d.xp[i] = 0.0;
        if (tid == 0) d.scalars[4] = 1.0;
        return;
    }
    if (tid == 0 && !(flags & 64)) d.scalars[4] = 0.0;  // (chunk CTAs only ever raise the flag; the host clears it)
    if (!(flags & 4)) return;
    // ---- backward: L^T x = y (row oriented); xp already holds y from the fused forward pass
    const int j_hi = (flags & 32) ? n_band - 1 : n - 1, j_lo = (flags & 16) ? n_band : 0;
    for (int j = j_hi; j >= j_lo; j--) {
        const int fj = firstp[j];
        const double* rowp = d.S + (size_t)rowoffp[j] * 36;
        const double* D = rowp + (size_t)(j - fj) * 36;
        if (tid == 0) {
            double v[6];
#pragma unroll
            for (int r = 5; r >= 0; r--) {
                double s = d.xp[j * 6 + r];
#pragma unroll
                for (int p = r + 1; p < 6; p++) s -= D[p * 6 + r] * v[p];
                v[r] = s * d.invd[j * 6 + r];
            }
#pragma unroll
            for (int r = 0; r < 6; r++) { d.xp[j * 6 + r] = v[r]; yk[r] = v[r]; }
        }
        __syncthreads();
        for (int it = tid; it < (j - fj) * 6; it += kSolveThreads) {
            const int i = fj + it / 6, c = it % 6;
            const double* Lb = rowp + (size_t)(i - fj) * 36;
            double s = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) s += Lb[r * 6 + c] * yk[r];
            d.xp[i * 6 + c] -= s;
        }
        __syncthreads();
    }
#ifdef CORB_BA_TRACE
    if (tid == 0) printf("k_ba_solve backward over %d rows: %lld cyc\n", n, clock64() - trb0);
#endif
}

// ---- K12e: the dense border block (border columns n_band .. n-1), left-looking, one CTA.
//      The right-looking sweep applied every column's trailing update as read-modify-writes on the whole remaining
//      block: with ~50 active rows that is 1 400 blocks per column through one SM's L2 port, and each pass waits for
//      the previous one (2 ms per solve). Left-looking, block (a, k) is formed when column k is reached:
//      T_ak = A_ak - sum_{c < k} L_ac L_kc^T reads only finished blocks (contiguous in each row's storage), one warp per
//      block, no read-modify-write; warp 0 forms the diagonal block first and factors it while the other warps form the
//      rest of the column. Per entry the terms are subtracted in ascending column order with the same 6-term sums, so
//      the factor is bit-identical to the right-looking one. The band-column terms of these blocks were applied by
//      k_ba_border_syrk; the forward substitution (y_k) rides along.
constexpr int kDenseU = 8;  // blocks of a row fetched per batch
__global__ void __launch_bounds__(1024) k_ba_border_dense(BaDev d, int n_band) {
    extern __shared__ double dsm[];  // T[nbord + 1][36] | ys[nbord][6] | Lrow[nbord][36] | stage[32 warps][kDenseU][36]
    __shared__ double Lkk[36], Lki[6], bk[6];
    __shared__ int s_fail;
    const int n = d.Pf, nbord = n - n_band;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* T = dsm;
    double* ys = dsm + (size_t)(nbord + 1) * 36;
    double* Lrow = ys + (size_t)nbord * 6;                                   // row k's finished blocks L_kc, c = ck0 .. k-1
    double* stage = Lrow + (size_t)nbord * 36 + (size_t)warp * kDenseU * 36;  // this warp's batch of L_ac blocks
    if (tid == 0) s_fail = d.scalars[4] != 0.0;
    __syncthreads();
    const int r0 = lane / 6, c0 = lane - r0 * 6, c1 = 2 + lane;  // entries `lane` and `32 + lane` (lane < 4)
    for (int k = n_band; k < n && !s_fail; k++) {
        const int cb = d.coloff[k], nact = d.coloff[k + 1] - cb;
        const int* rows = d.col_rows + cb;
        const int fk = d.first[k];
        const double* rowk = d.S + (size_t)(d.rowoff[k] - fk) * 36;  // block (k, c) at rowk + c * 36
        const int ck0 = max(n_band, fk);
        for (int i = tid; i < (k - ck0) * 36; i += 1024) Lrow[i] = rowk[(size_t)ck0 * 36 + i];  // shared by every block of the column
        __syncthreads();
        // ---- form the column: item 0 = diagonal, item i = (rows[i - 1], k)
        for (int it = warp; it <= nact; it += 32) {
            const int a = it == 0 ? k : rows[it - 1];
            const int fa = d.first[a];
            const double* rowa = d.S + (size_t)(d.rowoff[a] - fa) * 36;
            double t0 = rowa[(size_t)k * 36 + lane];
            double t1 = lane < 4 ? rowa[(size_t)k * 36 + 32 + lane] : 0.0;
            const int clo = max(n_band, max(fa, fk));
            for (int cb0 = clo; cb0 < k; cb0 += kDenseU) {
                // row a's blocks are contiguous: fetch a batch with all loads in flight, then use it from shared memory
                const int nb = min(kDenseU, k - cb0);
                double v[(kDenseU * 36 + 31) / 32];
#pragma unroll
                for (int u = 0; u < (kDenseU * 36 + 31) / 32; u++)
                    v[u] = lane + 32 * u < nb * 36 ? rowa[(size_t)cb0 * 36 + lane + 32 * u] : 0.0;
#pragma unroll
                for (int u = 0; u < (kDenseU * 36 + 31) / 32; u++)
                    if (lane + 32 * u < nb * 36) stage[lane + 32 * u] = v[u];
                __syncwarp();
                for (int c = cb0; c < cb0 + nb; c++) {
                    const double* La = stage + (size_t)(c - cb0) * 36;
                    const double* Lb = Lrow + (size_t)(c - ck0) * 36;
                    double s0 = 0, s1 = 0;
#pragma unroll
                    for (int q = 0; q < 6; q++) s0 += La[r0 * 6 + q] * Lb[c0 * 6 + q];
                    t0 = t0 - s0;
                    if (lane < 4) {
#pragma unroll
                        for (int q = 0; q < 6; q++) s1 += La[30 + q] * Lb[c1 * 6 + q];
                        t1 = t1 - s1;
                    }
                }
                __syncwarp();
            }
            T[it * 36 + lane] = t0;
            if (lane < 4) T[it * 36 + 32 + lane] = t1;
            if (it == 0) {  // warp 0: factor the diagonal block right away
                __syncwarp();
                double* D = const_cast<double*>(rowk) + (size_t)k * 36;
                if (!warp_chol6(T, nullptr, Lkk, lane, Lki, d.invd + (size_t)k * 6) && lane == 0) s_fail = 1;
                if (lane < 21) {  // warp_chol6 left L_kk in T[0..35] and Lkk; the factor's home is the skyline
                    const int rr = lane >= 15 ? 5 : lane >= 10 ? 4 : lane >= 6 ? 3 : lane >= 3 ? 2 : lane >= 1 ? 1 : 0;
                    const int cc = lane - rr * (rr + 1) / 2;
                    D[rr * 6 + cc] = Lkk[rr * 6 + cc];
                    if (rr != cc) D[cc * 6 + rr] = 0.0;
                }
            }
        }
        if (warp == 31 && lane < 6) {  // b_k -= sum_c L_kc y_c over the earlier border columns (ascending c)
            double b = d.xp[k * 6 + lane];
            for (int c = max(n_band, fk); c < k; c++) {
                const double* Lr = rowk + (size_t)c * 36 + lane * 6;
                const double* yc = ys + (size_t)(c - n_band) * 6;
                double s2 = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) s2 += Lr[q] * yc[q];
                b -= s2;
            }
            bk[lane] = b;
        }
        __syncthreads();
        if (s_fail) break;
        // ---- L_ak = T_ak L_kk^-T, one thread per row of a block; y_k = L_kk^-1 b_k
        for (int it = tid; it < nact * 6; it += 1024) {
            const int i = it / 6, r = it - i * 6;
            const int a = rows[i];
            const double* Tr = T + (size_t)(i + 1) * 36 + r * 6;
            double* out = d.S + (size_t)(d.rowoff[a] + k - d.first[a]) * 36 + r * 6;
            double v[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double s2 = Tr[c];
#pragma unroll
                for (int p = 0; p < c; p++) s2 -= v[p] * Lkk[c * 6 + p];
                v[c] = s2 * Lki[c];
            }
#pragma unroll
            for (int c = 0; c < 6; c++) out[c] = v[c];
        }
        if (tid == 1023) {
            double v[6];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double s2 = bk[r];
#pragma unroll
                for (int p = 0; p < r; p++) s2 -= Lkk[r * 6 + p] * v[p];
                v[r] = s2 * Lki[r];
            }
#pragma unroll
            for (int r = 0; r < 6; r++) { d.xp[k * 6 + r] = v[r]; ys[(size_t)(k - n_band) * 6 + r] = v[r]; }
        }
        __syncthreads();
    }
    if (s_fail && tid == 0) d.scalars[4] = 1.0;
}

// ---- border structure: bstart[b][q] = the first band column of chunk q that border row b shares a landmark with
__global__ void __launch_bounds__(256) k_ba_border_starts(BaDev d, int n_band, int nq, int* __restrict__ bstart) {
    const int l = blockIdx.x * 256 + threadIdx.x;
    if (l >= d.L) return;
    const int e0 = d.lm_off[l], e1 = d.lm_off[l + 1];
    bool any = false;
    for (int e = e0; e < e1; e++) any |= d.pfree[d.e_pose[e]] >= n_band;
    if (!any) return;
    for (int eb = e0; eb < e1; eb++) {
        const int b = d.pfree[d.e_pose[eb]];
        if (b < n_band) continue;
        for (int e = e0; e < e1; e++) {
            const int p = d.pfree[d.e_pose[e]];
            if (p < 0 || p >= n_band) continue;
            int q = 0;
            while (q + 1 < nq && d.chunk_start[q + 1] <= p) q++;
            atomicMin(&bstart[(size_t)(b - n_band) * nq + q], p);
        }
    }
}
__global__ void __launch_bounds__(256) k_ba_fill_int(int* p, int n, int v) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}
// ---- set-up: neighbour ranges of the keyframes on the device. nbr[j] = smallest, nbr[Pf + j] = largest free keyframe (in the
//      numbering `idx`) that shares a free landmark with keyframe j. The ordering derives the border, the chunks and the
//      envelope from them; it asks three times (natural order, border last, separators last), and walking the 1 M edges on
//      the host each time was half of the set-up.
__global__ void __launch_bounds__(256) k_ba_nbr_init(int* __restrict__ nbr, int Pf) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < Pf) { nbr[j] = j; nbr[Pf + j] = j; }
}
__global__ void __launch_bounds__(256) k_ba_nbr_range(const int* __restrict__ e_pose, const int* __restrict__ lm_off, const int* __restrict__ lfree,
                                                      const int* __restrict__ idx, int L, int Pf, int* __restrict__ nbr) {
    const int l = blockIdx.x * 256 + threadIdx.x;
    if (l >= L || lfree[l] < 0) return;
    const int k0 = lm_off[l], k1 = lm_off[l + 1];
    int lo = Pf, hi = -1;
    for (int k = k0; k < k1; k++) {
        const int pj = idx[e_pose[k]];
        if (pj >= 0) { lo = min(lo, pj); hi = max(hi, pj); }
    }
    if (hi < 0) return;
    for (int k = k0; k < k1; k++) {
        const int pj = idx[e_pose[k]];
        if (pj >= 0) {
            if (lo < pj) atomicMin(&nbr[pj], lo);  // (min / max commute: the result does not depend on the order of the atomics)
            if (hi > pj) atomicMax(&nbr[Pf + pj], hi);
        }
    }
}

__global__ void __launch_bounds__(256) k_ba_int_to_double(const int* a, double* b, int n, int back, int* a_out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (back) a_out[i] = (int)b[i]; else b[i] = (double)a[i];
}

// ---- K12c: the border rows (keyframes with long-range links, ordered last) against the finished band factor.
//      L_jk = (A_jk - sum_{c<k} L_jc L_kc^T) L_kk^-T only needs row j itself and the band rows, so every border row is
//      swept left to right on its own: one small CTA per border row, all rows concurrently, instead of riding along as
//      active rows of every band column in the one-CTA sweep (they were 80 % of its active rows). Per entry the
//      updates are subtracted in ascending column order with the same 6-term sums as the right-looking sweep, and the
//      right-hand side follows the same way (b_j -= L_jk y_k), so the result is bit-identical to it.
//      Operands of column k + 1 (A_j,k+1, the L_k+1,c blocks, L_k+1,k+1, y_k+1) do not depend on the running row and
//      are fetched while column k is computed.
__global__ void __launch_bounds__(64) k_ba_border_rows(BaDev d, int n_band, int wmax, int stage) {
    // wmax = longest band envelope (blocks left of the diagonal); dynamic shared memory: ring[wmax + 1] | Lkc[2][wmax]
    extern __shared__ double bsm[];
    const int kBorderRing = wmax + 1;
    double (*ring)[36] = reinterpret_cast<double (*)[36]>(bsm);  // L_jc of the last columns, slot c % kBorderRing
    double (*Lkc0)[36] = reinterpret_cast<double (*)[36]>(bsm + (size_t)kBorderRing * 36);  // band blocks L_kc, c = lo..k-1,
    double (*Lkc1)[36] = Lkc0 + wmax;                                                       // of the column computed / fetched
    __shared__ double Lkk[2][36];
    __shared__ double yk[2][6];
    __shared__ double Lki[2][6];
    __shared__ double T[36];
    const int j = n_band + blockIdx.x;
    const int tid = threadIdx.x;
    const int nq = gridDim.y, q = blockIdx.y;  // this CTA sweeps the columns of band chunk q only (chunks are decoupled)
#ifdef CORB_BA_TRACE
    const long long tr0 = clock64();
#endif
    if (d.scalars[4] != 0.0) return;  // the band factorisation failed
    // first / rowoff of the band rows sit on the address path of every fetch: staged in shared memory when they fit
    const int* firstp = d.first;
    const int* rowoffp = d.rowoff;
    if (stage) {
        int* sidx = reinterpret_cast<int*>(Lkc1 + wmax);
        for (int i = tid; i < n_band; i += 64) { sidx[i] = d.first[i]; sidx[n_band + i] = d.rowoff[i]; }
        firstp = sidx;
        rowoffp = sidx + n_band;
        __syncthreads();
    }
    const int fj = d.first[j];
    double* rowp = d.S + (size_t)d.rowoff[j] * 36;  // block of column k at rowp + (k - fj) * 36
    const int r = tid / 6, c0 = tid - r * 6;
    // Operands of a column are fetched into registers while the previous column is computed and committed to the
    // shared buffers afterwards (a shared-memory store right behind its load would stall the in-order thread for the
    // whole L2 round trip). Up to kPre band blocks travel in registers; longer envelopes commit directly.
    constexpr int kPre = 6;
    double pa = 0, pkk = 0, py = 0, pl[kPre];
    auto fetch_issue = [&](int k, int buf) {
        const int fk = firstp[k];
        const int lo = max(fj, fk);
        const double* bandrow = d.S + (size_t)rowoffp[k] * 36;
        if (tid < 36) {
            pa = rowp[(size_t)(k - fj) * 36 + tid];
            pkk = bandrow[(size_t)(k - fk) * 36 + tid];
#pragma unroll
            for (int i = 0; i < kPre; i++) pl[i] = lo + i < k ? bandrow[(size_t)(lo + i - fk) * 36 + tid] : 0.0;
            double (*Lb)[36] = buf ? Lkc1 : Lkc0;
            for (int c = lo + kPre; c < k; c++) Lb[c - lo][tid] = bandrow[(size_t)(c - fk) * 36 + tid];
        } else if (tid < 42) {
            py = d.xp[k * 6 + tid - 36];
        } else if (tid < 48) {
            py = d.invd[k * 6 + tid - 42];
        }
    };
    auto fetch_commit = [&](int k, int buf) {
        const int lo = max(fj, firstp[k]);
        if (tid < 36) {
            double (*Lb)[36] = buf ? Lkc1 : Lkc0;
#pragma unroll
            for (int i = 0; i < kPre; i++)
                if (lo + i < k) Lb[i][tid] = pl[i];
            Lkk[buf][tid] = pkk;
        } else if (tid < 42) {
            yk[buf][tid - 36] = py;
        } else if (tid < 48) {
            Lki[buf][tid - 42] = py;
        }
    };
    double a_cur = 0;
    double bj = 0.0;  // this chunk's part of sum_k L_jk y_k (k_ba_border_rhs subtracts the parts in chunk order)
    // the row has no entry in this chunk left of its first coupling there (bstart); columns skipped that way read as zero
    // from the ring, and their blocks in the skyline stay zero from the memset of the trial
    const int k_hi = min(d.chunk_start[q + 1], n_band);
    const int k_lo = min(k_hi, max(max(fj, d.chunk_start[q]), d.bstart[(size_t)blockIdx.x * nq + q]));
    for (int i = tid; i < kBorderRing * 36; i += 64) ring[0][i] = 0.0;
    __syncthreads();
    if (k_lo < k_hi) {
        fetch_issue(k_lo, 0);
        fetch_commit(k_lo, 0);
        a_cur = pa;
    }
    __syncthreads();
    for (int k = k_lo; k < k_hi; k++) {
        const int buf = (k - k_lo) & 1;
        if (k + 1 < k_hi) fetch_issue(k + 1, buf ^ 1);
        const int lo = max(fj, firstp[k]);
        if (tid < 36) {
            double t = a_cur;
            for (int c = lo; c < k; c++) {
                const double* La = ring[c % kBorderRing];
                const double* Lb = (buf ? Lkc1 : Lkc0)[c - lo];
                double s0 = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) s0 += La[r * 6 + q] * Lb[c0 * 6 + q];
                t = t - s0;
            }
            T[tid] = t;
        }
        __syncthreads();
        if (tid < 6) {  // row `tid` of L_jk = T L_kk^-T, then b_j[tid] -= L_jk[tid, :] y_k
            const double* Lk = Lkk[buf];
            double v[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double s = T[tid * 6 + c];
#pragma unroll
                for (int p = 0; p < c; p++) s -= v[p] * Lk[c * 6 + p];
                v[c] = s * Lki[buf][c];
            }
            double* out = rowp + (size_t)(k - fj) * 36 + tid * 6;
            double* rg = ring[k % kBorderRing] + tid * 6;
            double s = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                out[c] = v[c];
                rg[c] = v[c];
                s += v[c] * yk[buf][c];
            }
            bj += s;
        }
        if (k + 1 < k_hi) {
            fetch_commit(k + 1, buf ^ 1);
            a_cur = pa;
        }
        __syncthreads();
    }
    if (tid < 6) d.bpart[((size_t)blockIdx.x * nq + q) * 6 + tid] = bj;
#ifdef CORB_BA_TRACE
    if (tid == 0 && blockIdx.x == 0 && q == 0) printf("k_ba_border_rows row %d chunk 0: %d columns, %lld cyc\n", j, k_hi - k_lo, clock64() - tr0);
#endif
}

// b_j -= sum over the chunks (in chunk order) of their parts of sum_k L_jk y_k
__global__ void __launch_bounds__(256) k_ba_border_rhs(BaDev d, int n_band, int nq) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= (d.Pf - n_band) * 6 || d.scalars[4] != 0.0) return;
    const int b = i / 6, r = i - b * 6;
    double x = d.xp[(size_t)n_band * 6 + i];
    for (int q = 0; q < nq; q++) x -= d.bpart[((size_t)b * nq + q) * 6 + r];
    d.xp[(size_t)n_band * 6 + i] = x;
}

// ---- K12d: backward substitution over the band rows, one warp per independent chunk. Row j only touches the <= wmax rows of its envelope,
//      so the running right-hand sides live in a shared ring and the blocks of row j - 1 are fetched while row j is
//      solved; same operation order as the row loop at the end of k_ba_solve (bit-identical), without its two block
//      barriers and global read-modify-writes per row.
__global__ void __launch_bounds__(32) k_ba_band_backward(BaDev d, int n_band, int wmax, int stage) {
    extern __shared__ double bsm[];
    const int R = wmax + 1;
    double* xs = bsm;                       // [R][6] ring, slot j % R: xp of rows j - wmax .. j
    double* blk0 = bsm + (size_t)R * 6;     // [R][36] blocks of the row being solved: envelope blocks then the diagonal
    double* blk1 = blk0 + (size_t)R * 36;   // ... of the next row (being fetched)
    __shared__ double v[6];
    __shared__ double inv2[2][6];  // reciprocal diagonal of the row being solved / fetched
    const int lane = threadIdx.x;
#ifdef CORB_BA_TRACE
    const long long tr0 = clock64();
#endif
    if (d.scalars[4] != 0.0 || n_band <= 0) return;
    const int* firstp = d.first;
    const int* rowoffp = d.rowoff;
    if (stage) {
        int* sidx = reinterpret_cast<int*>(blk1 + (size_t)R * 36);
        for (int i = lane; i < n_band; i += 32) { sidx[i] = d.first[i]; sidx[n_band + i] = d.rowoff[i]; }
        firstp = sidx;
        rowoffp = sidx + n_band;
        __syncwarp();
    }
    constexpr int kPre = 8;  // doubles per lane that travel in registers (covers wmax <= 6)
    double pre[kPre], px = 0, pinv = 0;
    auto row_blocks = [&](int j) { return j - firstp[j] + 1; };
    auto issue = [&](int j, double* dst) {  // row j's blocks first[j]..j (contiguous in S) -> registers / dst
        const double* src = d.S + (size_t)rowoffp[j] * 36;
        const int n = row_blocks(j) * 36;
#pragma unroll
        for (int i = 0; i < kPre; i++) pre[i] = lane + 32 * i < n ? src[lane + 32 * i] : 0.0;
        if (lane >= 26) pinv = d.invd[j * 6 + lane - 26];
        for (int i = lane + 32 * kPre; i < n; i += 32) dst[i] = src[i];
    };
    auto commit = [&](int j, double* dst) {
        const int n = row_blocks(j) * 36;
#pragma unroll
        for (int i = 0; i < kPre; i++)
            if (lane + 32 * i < n) dst[lane + 32 * i] = pre[i];
        if (lane >= 26) inv2[dst == blk0 ? 0 : 1][lane - 26] = pinv;
    };
    const int j_lo = d.chunk_start[blockIdx.x], j_hi = d.chunk_start[blockIdx.x + 1] - 1;
    if (j_hi < j_lo) return;
    // rows j_hi .. j_hi - R + 1 into the ring
    for (int i = lane; i < R * 6; i += 32) {
        const int j = j_hi - i / 6;
        if (j >= 0) xs[(j % R) * 6 + i % 6] = d.xp[j * 6 + i % 6];
    }
    issue(j_hi, blk0);
    commit(j_hi, blk0);
    __syncwarp();
#ifdef CORB_BA_TRACE
    long long tr_solve = 0, tr_upd = 0, tr_commit = 0, tr_issue = 0;
#endif
    for (int j = j_hi; j >= j_lo; j--) {
#ifdef CORB_BA_TRACE
        const long long tz = clock64();
#endif
        double* cur = ((j_hi - j) & 1) ? blk1 : blk0;
        double* nxt = ((j_hi - j) & 1) ? blk0 : blk1;
        const int fj = firstp[j], nb = j - fj;
        if (j > j_lo) issue(j - 1, nxt);
        const int jin = j - R;  // row entering the ring once row j is done (its slot)
        if (jin >= 0 && lane < 6) px = d.xp[jin * 6 + lane];
        const double* D = cur + (size_t)nb * 36;
#ifdef CORB_BA_TRACE
        const long long ta = clock64();
#endif
        if (lane == 0) {
            double w[6];
#pragma unroll
            for (int r = 5; r >= 0; r--) {
                double s = xs[(j % R) * 6 + r];
#pragma unroll
                for (int p = r + 1; p < 6; p++) s -= D[p * 6 + r] * w[p];
                w[r] = s * inv2[cur == blk0 ? 0 : 1][r];
            }
#pragma unroll
            for (int r = 0; r < 6; r++) { d.xp[j * 6 + r] = w[r]; v[r] = w[r]; }
        }
        __syncwarp();
#ifdef CORB_BA_TRACE
        const long long tb = clock64();
#endif
        for (int it = lane; it < nb * 6; it += 32) {
            const int i = fj + it / 6, c = it % 6;
            const double* Lb = cur + (size_t)(i - fj) * 36;
            double s = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) s += Lb[r * 6 + c] * v[r];
            xs[(i % R) * 6 + c] -= s;
        }
        __syncwarp();
#ifdef CORB_BA_TRACE
        const long long tc = clock64();
#endif
        if (j > j_lo) commit(j - 1, nxt);
        if (jin >= 0 && lane < 6) xs[(j % R) * 6 + lane] = px;
        __syncwarp();
#ifdef CORB_BA_TRACE
        tr_solve += tb - ta; tr_upd += tc - tb; tr_commit += clock64() - tc; tr_issue += ta - tz;
#endif
    }
#ifdef CORB_BA_TRACE
    if (lane == 0 && blockIdx.x == 0) printf("k_ba_band_backward %d rows: %lld cyc; per row issue %lld solve %lld update %lld commit %lld\n", n_band, clock64() - tr0,
                          tr_issue / n_band, tr_solve / n_band, tr_upd / n_band, tr_commit / n_band);
#endif
}

// ---- K12b: deferred update of the border block (keyframes with long-range links, ordered last):
//      S(j, i) -= sum_k L_jk L_ik^T over the band columns k shared by the envelopes of border rows j >= i.
//      One warp per (j, i) pair; lane = entry of the 6x6 block; fixed k order => deterministic.
//      The band columns are independent chunks, so a pair's sum is cut at the chunk boundaries: one warp per
//      (pair, chunk) writes a partial block (nq > 1), and k_ba_border_syrk_apply subtracts the partials in chunk order.
__global__ void __launch_bounds__(256) k_ba_border_syrk(BaDev d, int n_band, int nq) {
    const int nbord = d.Pf - n_band;
    const int wi = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int p = wi / nq, q = wi - p * nq;
    if (p >= nbord * (nbord + 1) / 2) return;
    const int lane = threadIdx.x & 31;
    int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > p) a--;
    while ((a + 1) * (a + 2) / 2 <= p) a++;
    const int b = p - a * (a + 1) / 2;
    const int j = n_band + a, i = n_band + b;
    const int k1 = min(n_band, d.chunk_start[q + 1]);
    const int k0 = min(k1, max(max(max(d.first[j], d.first[i]), d.chunk_start[q]), max(d.bstart[(size_t)a * nq + q], d.bstart[(size_t)b * nq + q])));
    const double* Lj = d.S + (size_t)(d.rowoff[j] - d.first[j]) * 36;
    const double* Li = d.S + (size_t)(d.rowoff[i] - d.first[i]) * 36;
    const int r0 = lane / 6, c0 = lane - r0 * 6, c1 = 2 + lane;
    double s0 = 0, s1 = 0;
    for (int k = k0; k < k1; k++) {
        const double* A = Lj + (size_t)k * 36;
        const double* B = Li + (size_t)k * 36;
#pragma unroll
        for (int q2 = 0; q2 < 6; q2++) s0 += A[r0 * 6 + q2] * B[c0 * 6 + q2];
        if (lane < 4) {
#pragma unroll
            for (int q2 = 0; q2 < 6; q2++) s1 += A[30 + q2] * B[c1 * 6 + q2];
        }
    }
    double* out = d.syrk_part + ((size_t)p * nq + q) * 36;
    out[lane] = s0;
    if (lane < 4) out[32 + lane] = s1;
}

__global__ void __launch_bounds__(256) k_ba_border_syrk_apply(BaDev d, int n_band, int nq) {
    const int nbord = d.Pf - n_band;
    const int p = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (p >= nbord * (nbord + 1) / 2) return;
    const int lane = threadIdx.x & 31;
    int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > p) a--;
    while ((a + 1) * (a + 2) / 2 <= p) a++;
    const int b = p - a * (a + 1) / 2;
    const int j = n_band + a, i = n_band + b;
    double* T = d.S + (size_t)(d.rowoff[j] + i - d.first[j]) * 36;
    double s0 = 0, s1 = 0;
    for (int q = 0; q < nq; q++) {
        const double* part = d.syrk_part + ((size_t)p * nq + q) * 36;
        s0 += part[lane];
        if (lane < 4) s1 += part[32 + lane];
    }
    T[lane] -= s0;
    if (lane < 4) T[32 + lane] -= s1;
}

// ---- K13a: xl = Dinv (bl - Hpl^T xp) per landmark
__global__ void __launch_bounds__(256) k_ba_backsub(BaDev d) {
    const int l = blockIdx.x * 256 + threadIdx.x;
    if (l >= d.L) return;
    double o[3] = {0, 0, 0};
    if (d.lfree[l] >= 0) {
        double c[3] = {d.bl[(size_t)l * 3], d.bl[(size_t)l * 3 + 1], d.bl[(size_t)l * 3 + 2]};
        for (int e = d.lm_off[l]; e < d.lm_off[l + 1]; e++) {
            const int j = d.pfree[d.e_pose[e]];
            if (j < 0) continue;
            const double* w = d.W + (size_t)e * 18;
            const double* x = d.xp + (size_t)j * 6;
#pragma unroll
            for (int a = 0; a < 6; a++) {
                c[0] -= w[a * 3] * x[a]; c[1] -= w[a * 3 + 1] * x[a]; c[2] -= w[a * 3 + 2] * x[a];
            }
        }
        const double* Di = d.Dinv + (size_t)l * 9;
#pragma unroll
        for (int a = 0; a < 3; a++) o[a] = Di[a * 3] * c[0] + Di[a * 3 + 1] * c[1] + Di[a * 3 + 2] * c[2];
    }
#pragma unroll
    for (int a = 0; a < 3; a++) d.xl[(size_t)l * 3 + a] = o[a];
}

// ---- K13b: manifold update T <- exp(delta) T (se3quat.h:217-257,108-113), X <- X + delta, from the backup copies
__global__ void __launch_bounds__(256) k_ba_update(BaDev d) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < d.P) {
        const int pf = d.pfree[i];
        double q[4] = {d.q0[4 * i], d.q0[4 * i + 1], d.q0[4 * i + 2], d.q0[4 * i + 3]};
        double t[3] = {d.t0[3 * i], d.t0[3 * i + 1], d.t0[3 * i + 2]};
        if (pf >= 0) {
            const double* dl = d.xp + (size_t)pf * 6;
            const double om[3] = {dl[0], dl[1], dl[2]}, up[3] = {dl[3], dl[4], dl[5]};
            const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
            const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
            double O2[9], R[9], V[9];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) O2[r * 3 + c] = O[r * 3] * O[c] + O[r * 3 + 1] * O[3 + c] + O[r * 3 + 2] * O[6 + c];
            if (theta < 0.00001) {
#pragma unroll
                for (int k = 0; k < 9; k++) { R[k] = (k % 4 == 0 ? 1.0 : 0.0) + O[k] + O2[k]; V[k] = R[k]; }
            } else {
                const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / (theta * theta * theta);
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    R[k] = (k % 4 == 0 ? 1.0 : 0.0) + a * O[k] + b * O2[k];
                    V[k] = (k % 4 == 0 ? 1.0 : 0.0) + b * O[k] + c * O2[k];
                }
            }
            double qd[4], td[3], Rd[9];
            R_to_quat(R, qd);
            quat_normalize(qd);
#pragma unroll
            for (int r = 0; r < 3; r++) td[r] = V[r * 3] * up[0] + V[r * 3 + 1] * up[1] + V[r * 3 + 2] * up[2];
            quat_to_R(qd, Rd);
            double nt[3], nq[4];
#pragma unroll
            for (int r = 0; r < 3; r++) nt[r] = td[r] + Rd[r * 3] * t[0] + Rd[r * 3 + 1] * t[1] + Rd[r * 3 + 2] * t[2];
            nq[3] = qd[3] * q[3] - qd[0] * q[0] - qd[1] * q[1] - qd[2] * q[2];
            nq[0] = qd[3] * q[0] + qd[0] * q[3] + qd[1] * q[2] - qd[2] * q[1];
            nq[1] = qd[3] * q[1] + qd[1] * q[3] + qd[2] * q[0] - qd[0] * q[2];
            nq[2] = qd[3] * q[2] + qd[2] * q[3] + qd[0] * q[1] - qd[1] * q[0];
            quat_normalize(nq);
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = nq[k];
#pragma unroll
            for (int k = 0; k < 3; k++) t[k] = nt[k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) d.q[4 * i + k] = q[k];
#pragma unroll
        for (int k = 0; k < 3; k++) d.t[3 * i + k] = t[k];
    }
    for (int l = i; l < d.L; l += gridDim.x * 256) {
        const bool f = d.lfree[l] >= 0;
#pragma unroll
        for (int a = 0; a < 3; a++) d.X[(size_t)l * 3 + a] = d.X0[(size_t)l * 3 + a] + (f ? d.xl[(size_t)l * 3 + a] : 0.0);
    }
}

// ---- computeScale parts: [0] sum xp.bp + sum xl (lambda xl + bl), [1] sum xp^2
__global__ void __launch_bounds__(256) k_ba_scale(BaDev d, double lambda, int nb) {
    __shared__ double sm[32];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double part = 0, xx = 0;
    if (i < d.Pf * 6) { part = d.xp[i] * d.bp[i]; xx = d.xp[i] * d.xp[i]; }
    for (int k = i; k < d.L * 3; k += nb * 256)
        if (d.lfree[k / 3] >= 0) part += d.xl[k] * (lambda * d.xl[k] + d.bl[k]);
    part = block_reduce<false>(part, sm);
    xx = block_reduce<false>(xx, sm);
    if (threadIdx.x == 0) { d.partial[blockIdx.x] = part; d.partial[nb + blockIdx.x] = xx; }
}

}  // namespace corb

#include "ba_border_chol.cuh"

using namespace corb;

namespace {

// Device memory of a BA call comes from a per-device arena that outlives the call: a global BA allocates ~40 buffers
// (some > 100 MB) and cudaMalloc / cudaFree cost tens of milliseconds in total, which is comparable to the whole
// optimisation on a B200. The arena grows by slabs and is reused by the next call on the same device; a concurrent call
// on the same device (arena busy) falls back to plain cudaMalloc.
__global__ void __launch_bounds__(256) k_ba_iota_hist(const int* __restrict__ e_pose, int n, int* __restrict__ iota, int* __restrict__ cnt) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < n) {
        iota[k] = k;
        atomicAdd(&cnt[e_pose[k]], 1);
    }
}

struct BaArena {
    std::mutex mu;
    bool busy = false;
    bool release_when_idle = false;
    // the call's stream, timing events and pinned scalar mailbox live with the arena too (~2 ms of driver calls per BA otherwise)
    cudaStream_t stream = nullptr;
    cudaStream_t up_stream = nullptr;  // edge uploads of a sharded solve: beside the collectives of the set-up, not behind them
    cudaEvent_t ev_up = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double* h_scalars = nullptr;
    // page-locked staging slots for the upload of the caller's (pageable) edge arrays: kStageSlots threads each copy
    // chunks into their slot and enqueue the DMA from there - several times the rate of cudaMemcpy from pageable memory
    static constexpr int kStageSlots = 8;
    static constexpr size_t kStageBytes = (size_t)1 << 20;
    uint8_t* h_stage = nullptr;
    cudaEvent_t stage_ev[kStageSlots] = {};
    void free_slabs() {  // caller holds mu and the arena is idle
        for (auto& s : slabs) cudaFree(s.first);
        slabs.clear();
        cur = off = 0;
        release_when_idle = false;
        if (h_scalars) cudaFreeHost(h_scalars), h_scalars = nullptr;
        if (h_stage) cudaFreeHost(h_stage), h_stage = nullptr;
        for (auto& e : stage_ev) if (e) cudaEventDestroy(e), e = nullptr;
        if (ev0) cudaEventDestroy(ev0), ev0 = nullptr;
        if (ev1) cudaEventDestroy(ev1), ev1 = nullptr;
        if (stream) cudaStreamDestroy(stream), stream = nullptr;
        if (up_stream) cudaStreamDestroy(up_stream), up_stream = nullptr;
        if (ev_up) cudaEventDestroy(ev_up), ev_up = nullptr;
    }
    std::vector<std::pair<char*, size_t>> slabs;
    size_t cur = 0, off = 0;
    void* take(size_t bytes) {  // caller holds `busy`
        bytes = (bytes + 255) & ~(size_t)255;
        for (;;) {
            if (cur < slabs.size() && off + bytes <= slabs[cur].second) {
                void* q = slabs[cur].first + off;
                off += bytes;
                return q;
            }
            if (cur + 1 < slabs.size()) { cur++; off = 0; continue; }
            const size_t sz = std::max(bytes, (size_t)256 << 20);
            void* q = nullptr;
            if (cudaMalloc(&q, sz) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            slabs.emplace_back((char*)q, sz);
            cur = slabs.size() - 1;
            off = 0;
        }
    }
};
static BaArena g_arena[16];

struct BaHost {
    int device;
    BaArena* arena = nullptr;  // non-null while this call owns its device's arena
    bool arena_owns_handles = false;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    BaDev d;
    size_t s_doubles = 0;  // S blocks * 36
    int lact_cap = 0;      // active rows of one column that fit the solve kernel's shared memory
    int band_wmax = 1;
    int stage_nnz_b = -1;  // entries of the band column lists when the index arrays fit the solve kernel's shared memory
    size_t stage_bytes_b = 0;
    size_t band_idx_bytes = 0;
    size_t border_dense_bytes = 0;
    bool border_dense_ok = false;  // the left-looking border kernel's shared staging fits
    bool border_tiled = false;     // tiled multi-CTA border factorisation (k_ba_border_chol) usable
    int border_grid = 1;
    size_t border_tiled_bytes = 0;
    int n_chunks = 1;           // independent band chunks (columns chunk_start[q] .. chunk_start[q + 1])  // first[] / rowoff[] of the band rows staged by the border-row and band-backward kernels     // longest band envelope (blocks left of the diagonal)
    int n_band = 0;        // free keyframes before the border block (== Pf when there is no border)
    SchurPairs sp = {};      // sorted pair lists of the Schur rows (k_ba_schur_pairs); sp_ok: built and within the kernel's limits
    bool sp_ok = false;
    int sp_tcap = 0, sp_threads = 512;
    size_t sp_smem = 0;
    corb_allreduce_fn ar = nullptr;
    void* ar_user = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms_solve = 0;
    double* h_scalars = nullptr;  // pinned [8]
    // Sharded solve: every rank must take the same branch on the caller's stop flag (each chi2 / trial is a collective), so
    // the flag is sampled where chi2 is reduced, summed over the ranks next to it, and only the reduced value is branched on.
    const volatile uint8_t* stop_flag = nullptr;
    bool agreed_stop = false;
    bool stop_requested() const { return ar ? agreed_stop : (stop_flag && *stop_flag); }

    ~BaHost() {
        if (stream) cudaStreamSynchronize(stream);
        for (void* p : allocs) cudaFree(p);
        if (arena) {
            std::lock_guard<std::mutex> lk(arena->mu);
            arena->busy = false;
            if (arena->release_when_idle) arena->free_slabs();
        }
        if (!arena_owns_handles) {
            if (h_scalars) cudaFreeHost(h_scalars);
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
            if (stream) cudaStreamDestroy(stream);
        }
    }
    template <typename T>
    int alloc(T** p, size_t n) {
        const size_t bytes = std::max<size_t>(n, 1) * sizeof(T) + 256;
        void* q = arena ? arena->take(bytes) : nullptr;
        if (!q) {
            CORB_CUDA(cudaMalloc(&q, bytes));
            allocs.push_back(q);
        }
        *p = (T*)q;
        return CORB_OK;
    }
    template <typename T>
    int upload_raw(const T** p, const T* src, size_t n) {
        T* q;
        int rc = alloc(&q, n);
        if (rc != CORB_OK) return rc;
        if (n) CORB_CUDA(cudaMemcpyAsync(q, src, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        *p = q;
        return CORB_OK;
    }
    // pose_off / pose_edges: positions of every keyframe's edges inside the landmark-ordered edge arrays, ascending
    // (a stable sort of the edge index by keyframe; the fixed order keeps k_ba_build_pose's sums deterministic)
    int build_pose_csr() {
        int *keys_out, *vals_in, *vals_out, *off;
        int rc;
        if ((rc = alloc(&keys_out, (size_t)d.E)) != CORB_OK || (rc = alloc(&vals_in, (size_t)d.E)) != CORB_OK ||
            (rc = alloc(&vals_out, (size_t)d.E)) != CORB_OK || (rc = alloc(&off, (size_t)d.P + 1)) != CORB_OK)
            return rc;
        CORB_CUDA(cudaMemsetAsync(off, 0, ((size_t)d.P + 1) * sizeof(int), stream));
        if (d.E > 0) {
            k_ba_iota_hist<<<(d.E + 255) / 256, 256, 0, stream>>>(d.e_pose, d.E, vals_in, off + 1);
            int bits = 1;
            while ((1 << bits) < d.P) bits++;
            size_t tmp_bytes = 0;
            CORB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d.e_pose, keys_out, vals_in, vals_out, d.E, 0, bits, stream));
            size_t scan_bytes = 0;
            CORB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, off + 1, off + 1, d.P, stream));
            char* tmp;
            if ((rc = alloc(&tmp, std::max(tmp_bytes, scan_bytes))) != CORB_OK) return rc;
            CORB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d.e_pose, keys_out, vals_in, vals_out, d.E, 0, bits, stream));
            CORB_CUDA(cub::DeviceScan::InclusiveSum(tmp, scan_bytes, off + 1, off + 1, d.P, stream));
        }
        d.pose_off = off;
        d.pose_edges = vals_out;
        return CORB_OK;
    }
    // Sorted pair lists for k_ba_schur_pairs, once per call (the structure is the same in every LM iteration).
    // `pairs_bound` = sum over landmarks of k (k + 1) / 2 observations pairs (exact when nothing is fixed), `nblocks` = envelope blocks.
    int build_schur_pairs(size_t pairs_bound, size_t nblocks) {
        sp_ok = false;
        const char* env = getenv("CORB_BA_SCHUR");
        if ((env && !strcmp(env, "old")) || d.P <= 0 || d.E <= 0 || pairs_bound == 0 || pairs_bound > ((size_t)1 << 27)) return CORB_OK;  // (1 GB of pair records at most; denser maps keep the edge walk)
        int th = 512, ch = 32;  // CORB_BA_SP=threads,chunk: A/B switch of the kernel shape
        if (const char* e2 = getenv("CORB_BA_SP")) sscanf(e2, "%d,%d", &th, &ch);
        if (th != 256) th = 512;
        if (ch != 64) ch = 32;
        sp_threads = th;
        sp.chunk = ch;
        const size_t items_bound = pairs_bound / ch + nblocks + 2 * (size_t)d.P + 16;
        int *row_np, *row_bound, *pair_off, *item_off, *item_cnt, *info;
        int2* row_el;
        uint2* pairs;
        int4* items;
        int rc;
        if ((rc = alloc(&row_np, (size_t)d.P)) != CORB_OK || (rc = alloc(&row_bound, (size_t)d.P)) != CORB_OK ||
            (rc = alloc(&pair_off, (size_t)d.P + 1)) != CORB_OK || (rc = alloc(&item_off, (size_t)d.P + 1)) != CORB_OK ||
            (rc = alloc(&item_cnt, (size_t)d.P)) != CORB_OK || (rc = alloc(&info, 4)) != CORB_OK ||
            (rc = alloc(&pairs, pairs_bound)) != CORB_OK || (rc = alloc(&items, items_bound)) != CORB_OK ||
            (rc = alloc(&row_el, (size_t)d.E)) != CORB_OK)
            return rc;
        CORB_CUDA(cudaMemsetAsync(info, 0, 4 * sizeof(int), stream));
        k_ba_pairs_count<<<d.P, 256, 0, stream>>>(d, row_np, row_bound, info, ch, row_el);
        k_ba_pairs_scan<<<1, 1024, 0, stream>>>(row_np, row_bound, d.P, pair_off, item_off);
        sp.row_pair_off = pair_off; sp.row_item_off = item_off; sp.row_item_cnt = item_cnt; sp.pairs = pairs; sp.items = items; sp.info = info; sp.row_el = row_el;
        CORB_SMEM_OPT_IN(k_ba_pairs_build);
        CORB_SMEM_OPT_IN((k_ba_schur_pairs<512, 32>));
        CORB_SMEM_OPT_IN((k_ba_schur_pairs<512, 64>));
        CORB_SMEM_OPT_IN((k_ba_schur_pairs<256, 32>));
        CORB_SMEM_OPT_IN((k_ba_schur_pairs<256, 64>));
        k_ba_pairs_build<<<d.P, kSpThreads, (size_t)kSpPairCap * 12, stream>>>(d, sp);
        CORB_CUDA(cudaGetLastError());
        int h_info[4] = {0, 0, 0, 1};
        CORB_CUDA(cudaMemcpyAsync(h_info, info, sizeof(h_info), cudaMemcpyDeviceToHost, stream));
        CORB_CUDA(cudaStreamSynchronize(stream));
        sp_tcap = std::max(1, h_info[0]);
        sp_smem = ((size_t)sp_tcap * 18 + (size_t)std::max(1, h_info[2]) * 36) * sizeof(double);
        sp_ok = !h_info[3] && sp_smem <= 200 * 1024;
        return CORB_OK;
    }
    template <typename T>
    int upload(const T** p, const std::vector<T>& v) {
        T* q;
        int rc = alloc(&q, v.size());
        if (rc != CORB_OK) return rc;
        if (!v.empty()) CORB_CUDA(cudaMemcpyAsync(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        *p = q;
        return CORB_OK;
    }
    int reduce(double* d_buf, size_t n, int op) {
        if (!ar || n == 0) return CORB_OK;
        CORB_CHECK(ar(ar_user, d_buf, n, op, (void*)stream) == 0, CORB_ERR_CUDA, "all-reduce callback failed");
        return CORB_OK;
    }
    int read_scalars(int n) {
        CORB_CUDA(cudaMemcpyAsync(h_scalars, d.scalars, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CORB_CUDA(cudaStreamSynchronize(stream));
        return CORB_OK;
    }
    int chi2(double* out) {
        const int nb = (d.E + 255) / 256;
        if (nb > 0) k_ba_chi2<<<nb, 256, 0, stream>>>(d);
        k_reduce_final<false><<<1, 256, 0, stream>>>(d.partial, nb, d.scalars + 6);
        if (ar) {  // scalars[7] = this rank's vote; h_scalars[5] is never a read-back target, the stream is idle between calls
            h_scalars[5] = (stop_flag && *stop_flag) ? 1.0 : 0.0;
            CORB_CUDA(cudaMemcpyAsync(d.scalars + 7, h_scalars + 5, sizeof(double), cudaMemcpyHostToDevice, stream));
        }
        int rc = reduce(d.scalars + 6, ar ? 2 : 1, 0);
        if (rc != CORB_OK) return rc;
        CORB_CUDA(cudaMemcpyAsync(h_scalars + 6, d.scalars + 6, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CORB_CUDA(cudaStreamSynchronize(stream));
        *out = h_scalars[6];
        if (ar) agreed_stop = h_scalars[7] != 0.0;
        return rc;
    }
    int build() {
        if (d.L > 0) k_ba_build_lm<<<(d.L + 127) / 128, 128, 0, stream>>>(d);
        if (d.P > 0) k_ba_build_pose<<<(d.P * 32 + 255) / 256, 256, 0, stream>>>(d);
        CORB_CUDA(cudaGetLastError());
        return CORB_OK;
    }
    int lambda_init(double* out) {  // computeLambdaInit: tau * max |diag H| over all free vertices, all ranks
        double mx = 0;
        const int nb0 = (d.Pf * 6 + 255) / 256, nb1 = (d.L * 3 + 255) / 256;
        if (nb0 > 0) {
            k_ba_copy_diag<<<nb0, 256, 0, stream>>>(d);  // xp <- diag(Hpp); summed over ranks before taking the max
            int rc = reduce(d.xp, (size_t)d.Pf * 6, 0);
            if (rc != CORB_OK) return rc;
            k_ba_maxdiag<<<nb0, 256, 0, stream>>>(d, 0);
            k_reduce_final<true><<<1, 256, 0, stream>>>(d.partial, nb0, d.scalars + 3);
            rc = read_scalars(4);
            if (rc != CORB_OK) return rc;
            mx = std::max(mx, h_scalars[3]);
        }
        {
            if (nb1 > 0) k_ba_maxdiag<<<nb1, 256, 0, stream>>>(d, 1);
            k_reduce_final<true><<<1, 256, 0, stream>>>(d.partial, nb1, d.scalars + 3);
            int rc = reduce(d.scalars + 3, 1, 2);
            if (rc != CORB_OK) return rc;
            rc = read_scalars(4);
            if (rc != CORB_OK) return rc;
            mx = std::max(mx, h_scalars[3]);
        }
        *out = 1e-5 * mx;  // _tau * maxDiagonal
        return CORB_OK;
    }

    // setLambda + BlockSolver::solve + update + computeScale; leaves the tentative estimate in q/t/X
    int trial(double lambda, bool* ok, double* scale) {
        if (d.L > 0) k_ba_dinv<<<(d.L + 255) / 256, 256, 0, stream>>>(d, lambda);
        if (d.P > 0) {
            CORB_CUDA(cudaMemsetAsync(d.S, 0, s_doubles * sizeof(double), stream));  // rows are accumulated into, not zeroed, by the kernel
            if (sp_ok) {  // one CTA per keyframe, sorted pair lists
                if (sp_threads == 512 && sp.chunk == 32) k_ba_schur_pairs<512, 32><<<d.P, 512, sp_smem, stream>>>(d, sp, sp_tcap);
                else if (sp_threads == 512) k_ba_schur_pairs<512, 64><<<d.P, 512, sp_smem, stream>>>(d, sp, sp_tcap);
                else if (sp.chunk == 32) k_ba_schur_pairs<256, 32><<<d.P, 256, sp_smem, stream>>>(d, sp, sp_tcap);
                else k_ba_schur_pairs<256, 64><<<d.P, 256, sp_smem, stream>>>(d, sp, sp_tcap);
            }
            else k_ba_schur_rows<<<d.P, 256, 0, stream>>>(d);                                    // one CTA per keyframe, edge walk
        }
        int rc = reduce(d.S, s_doubles + (size_t)d.Pf * 6, 0);
        if (rc != CORB_OK) return rc;
        if (d.Pf > 0) k_ba_add_lambda<<<(d.Pf * 6 + 255) / 256, 256, 0, stream>>>(d, lambda);
        cudaEventRecord(ev0, stream);
        {   // band columns (border x border updates deferred) -> border SYRK on the whole GPU -> border columns + backward
            const size_t sm = (size_t)lact_cap * 36 * sizeof(double);
            const int nbord = d.Pf - n_band;
            if (nbord > 0) {
                // band rows alone -> every border row against the band factor, concurrently -> border x border SYRK ->
                // dense border block + backward over the border rows -> backward over the band rows
                CORB_CUDA(cudaMemsetAsync(d.scalars + 4, 0, sizeof(double), stream));
                k_ba_solve<<<n_chunks, kSolveThreads, sm + stage_bytes_b, stream>>>(d, lact_cap, 0, n_band, n_band, 1 | 8 | 64, stage_nnz_b);
                k_ba_border_rows<<<dim3(nbord, n_chunks), 64, (size_t)(3 * band_wmax + 1) * 36 * sizeof(double) + band_idx_bytes, stream>>>(
                    d, n_band, band_wmax, band_idx_bytes > 0);
                k_ba_border_rhs<<<(nbord * 6 + 255) / 256, 256, 0, stream>>>(d, n_band, n_chunks);
                const int npairs = nbord * (nbord + 1) / 2;
                k_ba_border_syrk<<<(int)(((size_t)npairs * n_chunks * 32 + 255) / 256), 256, 0, stream>>>(d, n_band, n_chunks);
                k_ba_border_syrk_apply<<<(npairs * 32 + 255) / 256, 256, 0, stream>>>(d, n_band, n_chunks);
                if (border_tiled) {
                    // the dense border block on the whole GPU: tiled Cholesky with fp64 tensor-core updates + the border's
                    // forward / backward substitution, then the border solution's contribution to the band right-hand side
                    BaDev dd = d;
                    int nb_arg = n_band;
                    void* args[] = {&dd, &nb_arg};
                    CORB_CUDA(cudaLaunchCooperativeKernel((void*)k_ba_border_chol, dim3(border_grid), dim3(kBcThreads), args, border_tiled_bytes, stream));
                    k_ba_border_to_band<<<(n_band * 32 + 255) / 256, 256, 0, stream>>>(d, n_band);
                } else if (getenv("CORB_BA_BORDER_RL") || !border_dense_ok) {  // A/B switch / border too large for the staging buffer
                    k_ba_solve<<<1, kSolveThreads, sm, stream>>>(d, lact_cap, n_band, d.Pf, n_band, 4 | 16, -1);
                } else {
                    k_ba_border_dense<<<1, 1024, border_dense_bytes, stream>>>(d, n_band);
                    k_ba_solve<<<1, kSolveThreads, sm, stream>>>(d, lact_cap, d.Pf, d.Pf, n_band, 4 | 16, -1);  // backward over the border rows
                }
                k_ba_band_backward<<<n_chunks, 32, (size_t)(band_wmax + 1) * (6 + 72) * sizeof(double) + band_idx_bytes, stream>>>(
                    d, n_band, band_wmax, band_idx_bytes > 0);
            } else {
                k_ba_solve<<<1, kSolveThreads, sm + stage_bytes_b, stream>>>(d, lact_cap, 0, d.Pf, n_band, 1 | 4 | 8, stage_nnz_b);
            }
        }
        cudaEventRecord(ev1, stream);
        if (d.L > 0) k_ba_backsub<<<(d.L + 255) / 256, 256, 0, stream>>>(d);
        const int nbu = std::max(1, (std::max(d.P, std::min(d.L, 1 << 20)) + 255) / 256);
        k_ba_update<<<nbu, 256, 0, stream>>>(d);
        const int nb = std::max(1, std::min(1024, (std::max(d.Pf * 6, d.L * 3) + 255) / 256));
        const int nbs = std::max(nb, (d.Pf * 6 + 255) / 256);
        k_ba_scale<<<nbs, 256, 0, stream>>>(d, lambda, nbs);
        k_reduce_final<false><<<1, 256, 0, stream>>>(d.partial, nbs, d.scalars + 1);
        k_reduce_final<false><<<1, 256, 0, stream>>>(d.partial + nbs, nbs, d.scalars + 2);
        CORB_CUDA(cudaGetLastError());
        rc = reduce(d.scalars + 1, 1, 0);
        if (rc != CORB_OK) return rc;
        rc = read_scalars(5);
        if (rc != CORB_OK) return rc;
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        ms_solve += ms;
        *ok = h_scalars[4] == 0.0;
        *scale = h_scalars[1] + lambda * h_scalars[2];
        return CORB_OK;
    }
};

}  // namespace

// splits [0, n) over a few host threads (the structure build is memory bound: more than 8 threads do not help)
template <typename F>
static void host_parallel_for(int n, F fn) {
    const int nt = std::max(1, std::min(8, std::min((int)std::thread::hardware_concurrency(), n / 65536)));
    if (nt <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=]() { fn((int)((long long)n * t / nt), (int)((long long)n * (t + 1) / nt)); });
    for (auto& t : th) t.join();
}

extern "C" int corb_host_alloc(size_t bytes, void** out) {
    CORB_CHECK(out, CORB_ERR_INVALID, "bad argument");
    *out = nullptr;
    void* p = nullptr;
    CORB_CUDA(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable));
    *out = p;
    return CORB_OK;
}
extern "C" int corb_host_free(void* p) {
    if (p) CORB_CUDA(cudaFreeHost(p));
    return CORB_OK;
}

extern "C" int corb_ba_release_cache(int device) {
    CORB_CHECK(device >= 0 && device < 16, CORB_ERR_INVALID, "device %d out of range", device);
    std::lock_guard<std::mutex> lk(g_arena[device].mu);
    if (g_arena[device].busy) {
        g_arena[device].release_when_idle = true;
        return CORB_OK;
    }
    if (!g_arena[device].slabs.empty()) {
        CORB_CUDA(cudaSetDevice(device));
        g_arena[device].free_slabs();
    }
    return CORB_OK;
}

extern "C" int corb_ba_solve(corb_ba_problem* p, int iterations, const volatile uint8_t* stop, int robust, int device,
                             corb_ba_result* res, corb_allreduce_fn allreduce, void* allreduce_user) {
    const auto t_start = std::chrono::steady_clock::now();
    CORB_CHECK(p && res, CORB_ERR_INVALID, "problem/result is NULL");
    memset(res, 0, sizeof(*res));
    const int P = p->n_poses, L = p->n_points, E = p->n_edges;
    CORB_CHECK(P >= 0 && L >= 0 && E >= 0 && iterations >= 0, CORB_ERR_INVALID, "negative size");
    CORB_CHECK(P == 0 || (p->pose_q && p->pose_t && p->pose_fixed && p->pose_cam), CORB_ERR_INVALID, "pose arrays are NULL");
    CORB_CHECK(L == 0 || (p->point_xyz && p->point_fixed), CORB_ERR_INVALID, "point arrays are NULL");
    CORB_CHECK(E == 0 || (p->edge_pose && p->edge_point && p->edge_obs && p->edge_inv_sigma2), CORB_ERR_INVALID, "edge arrays are NULL");
    // one pass over the edges on a few threads: range check, and is the list grouped by landmark already (the order
    // Optimizer::BundleAdjustment creates it in, Optimizer.cc:131-196)?
    std::atomic<int> first_bad(E);
    std::atomic<bool> grouped_a(true);
    host_parallel_for(E, [&](int e0, int e1) {
        int bad = E;
        bool g = true;
        for (int e = e0; e < e1; e++) {
            const int a = p->edge_pose[e], b = p->edge_point[e];
            if ((unsigned)a >= (unsigned)P || (unsigned)b >= (unsigned)L) { bad = std::min(bad, e); continue; }
            if (e > 0 && p->edge_point[e - 1] > b) g = false;
        }
        if (!g) grouped_a.store(false);
        int cur = first_bad.load();
        while (bad < cur && !first_bad.compare_exchange_weak(cur, bad)) {}
    });
    {
        const int e = first_bad.load();
        CORB_CHECK(e >= E, CORB_ERR_INVALID, "edge %d references pose %d / point %d out of range", e, p->edge_pose[e], p->edge_point[e]);
    }
    int ndev = 0;
    CORB_CUDA(cudaGetDeviceCount(&ndev));
    CORB_CHECK(device >= 0 && device < ndev, CORB_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    CORB_CUDA(cudaSetDevice(device));
#ifdef CORB_BA_TRACE
    auto lap = [&](const char* what) {
        printf("  setup %-28s at %.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
#else
    auto lap = [&](const char*) {};
#endif
    lap("validated edges");
    BaHost H;
    H.device = device;
    if (device < 16 && !getenv("CORB_BA_NO_ARENA")) {
        std::lock_guard<std::mutex> lk(g_arena[device].mu);
        if (!g_arena[device].busy) {
            g_arena[device].busy = true;
            g_arena[device].cur = 0;
            g_arena[device].off = 0;
            H.arena = &g_arena[device];
        }
    }
    H.ar = allreduce;
    H.ar_user = allreduce_user;
    H.stop_flag = stop;
    if (H.arena) {
        BaArena& A = *H.arena;
        if (!A.stream) CORB_CUDA(cudaStreamCreateWithFlags(&A.stream, cudaStreamNonBlocking));
        if (!A.ev0) CORB_CUDA(cudaEventCreate(&A.ev0));
        if (!A.ev1) CORB_CUDA(cudaEventCreate(&A.ev1));
        if (!A.up_stream) {
            CORB_CUDA(cudaStreamCreateWithFlags(&A.up_stream, cudaStreamNonBlocking));
            CORB_CUDA(cudaEventCreateWithFlags(&A.ev_up, cudaEventDisableTiming));
        }
        if (!A.h_scalars) CORB_CUDA(cudaMallocHost(&A.h_scalars, 8 * sizeof(double)));
        if (!A.h_stage && (size_t)E * 40 >= ((size_t)16 << 20)) {
            CORB_CUDA(cudaMallocHost(&A.h_stage, BaArena::kStageSlots * BaArena::kStageBytes));
            for (auto& e : A.stage_ev) CORB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        H.stream = A.stream; H.ev0 = A.ev0; H.ev1 = A.ev1; H.h_scalars = A.h_scalars;
        H.arena_owns_handles = true;
    } else {
        CORB_CUDA(cudaStreamCreateWithFlags(&H.stream, cudaStreamNonBlocking));
        CORB_CUDA(cudaEventCreate(&H.ev0));
        CORB_CUDA(cudaEventCreate(&H.ev1));
        CORB_CUDA(cudaMallocHost(&H.h_scalars, 8 * sizeof(double)));
    }
    BaDev& d = H.d;
    memset(&d, 0, sizeof(d));
    d.P = P; d.L = L; d.E = E; d.robust = robust != 0;
    d.delta2d = (double)(float)sqrt(5.99);   // thHuber2D / thHuber3D are floats in the reference (Optimizer.cc:101-102)
    d.delta3d = (double)(float)sqrt(7.815);
    lap("cuda context / stream / events");
    // ---- host-side structure (the analogue of g2o's buildStructure, block_solver.hpp:143-295)
    std::vector<int> pfree(P), lfree(L);
    int Pf = 0, Lf = 0;
    for (int i = 0; i < P; i++) pfree[i] = p->pose_fixed[i] ? -1 : Pf++;
    for (int i = 0; i < L; i++) lfree[i] = p->point_fixed[i] ? -1 : Lf++;
    d.Pf = Pf;
    // Edges grouped by landmark (for every map point, its observations) are used in place; anything else is stably
    // sorted by landmark first. lm_off[l] = first edge of landmark l.
    const bool grouped = grouped_a.load();
    std::vector<int> lm_off(L + 1, 0);
    if (grouped) {  // sorted keys: landmark l starts where the key first reaches l; every entry is written exactly once
        host_parallel_for(E, [&](int e0, int e1) {
            for (int e = e0; e < e1; e++) {
                const int prev = e > 0 ? p->edge_point[e - 1] : -1, cur = p->edge_point[e];
                for (int l = prev + 1; l <= cur; l++) lm_off[l] = e;
            }
        });
        for (int l = (E > 0 ? p->edge_point[E - 1] : -1) + 1; l <= L; l++) lm_off[l] = E;
    } else {
        for (int e = 0; e < E; e++) lm_off[p->edge_point[e] + 1]++;
        for (int i = 0; i < L; i++) lm_off[i + 1] += lm_off[i];
    }
    std::vector<int> perm, e_pose_v, e_point_v;
    std::vector<double> e_obs_v, e_info_v;
    if (!grouped) {
        perm.resize(E);
        std::vector<int> cur(lm_off.begin(), lm_off.end() - 1);
        for (int e = 0; e < E; e++) perm[cur[p->edge_point[e]]++] = e;
        e_pose_v.resize(E); e_point_v.resize(E); e_obs_v.resize((size_t)E * 3); e_info_v.resize(E);
    }
    std::vector<int>& e_pose = e_pose_v;
    std::vector<int>& e_point = e_point_v;
    std::vector<double>& e_obs = e_obs_v;
    std::vector<double>& e_info = e_info_v;
    if (!grouped) host_parallel_for(E, [&](int k0, int k1) {
        for (int k = k0; k < k1; k++) {
            const int e = perm[k];
            e_pose[k] = p->edge_pose[e]; e_point[k] = p->edge_point[e];
            e_obs[3 * (size_t)k] = p->edge_obs[3 * (size_t)e]; e_obs[3 * (size_t)k + 1] = p->edge_obs[3 * (size_t)e + 1];
            e_obs[3 * (size_t)k + 2] = p->edge_obs[3 * (size_t)e + 2];
            e_info[k] = p->edge_inv_sigma2[e];
        }
    });
    const int* ep = grouped ? p->edge_pose : e_pose.data();      // edge arrays in landmark order
    const int* ept = grouped ? p->edge_point : e_point.data();
    const double* eobs = grouped ? p->edge_obs : e_obs.data();
    const double* einfo = grouped ? p->edge_inv_sigma2 : e_info.data();
    lap("edges grouped by landmark");
    // The edge arrays are the bulk of the upload (40 B per observation out of pageable memory: ~5 ms at E = 1 M, during which
    // the enqueuing thread is blocked). A helper thread enqueues them now, while this thread derives the ordering and the
    // envelope; it is joined before anything that reads them is enqueued. (With an all-reduce hook the ordering itself
    // synchronises on the stream, so the copies stay where they were.)
    cudaError_t up_err = cudaSuccess;  // (declared before the thread that writes it, so it outlives the join)
    struct Joiner {
        std::thread t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } uploader;
    // The bulk of the edge arrays goes up on the arena's upload stream while the ordering is derived; the call's stream - which
    // carries the ordering's own device work and, in a sharded solve, its collectives - waits for them once, before the first
    // kernel that reads the edges. The keyframe indices of the edges (what the ordering needs) go first, on the call's stream.
    cudaStream_t up_st = H.arena ? H.arena->up_stream : nullptr;
    const bool early_upload = E > 0 && up_st != nullptr;
    if (early_upload) {
        int *q_pose, *q_point;
        double *q_obs, *q_info, *q_X;
        int rc0;
        if ((rc0 = H.alloc(&q_pose, (size_t)E)) != CORB_OK || (rc0 = H.alloc(&q_point, (size_t)E)) != CORB_OK ||
            (rc0 = H.alloc(&q_obs, (size_t)E * 3)) != CORB_OK || (rc0 = H.alloc(&q_info, (size_t)E)) != CORB_OK ||
            (rc0 = H.alloc(&q_X, (size_t)L * 3)) != CORB_OK)
            return rc0;
        d.X = q_X;  // the landmark estimates (24 B each) travel with the edge arrays
        const double* src_X = p->point_xyz;
        const size_t bytes_X = (size_t)L * 3 * sizeof(double);
        d.e_pose = q_pose; d.e_point = q_point; d.e_obs = q_obs; d.e_info = q_info;
        CORB_CUDA(cudaMemcpyAsync(q_pose, ep, (size_t)E * sizeof(int), cudaMemcpyHostToDevice, H.stream));
        cudaStream_t st = up_st;
        BaArena* stage = H.arena && H.arena->h_stage ? H.arena : nullptr;
        uploader.t = std::thread([=, &up_err] {
            struct Job { char* dst; const char* src; size_t bytes; };
            const Job jobs[4] = {{(char*)q_obs, (const char*)eobs, (size_t)E * 3 * sizeof(double)}, {(char*)q_info, (const char*)einfo, (size_t)E * sizeof(double)},
                                 {(char*)q_point, (const char*)ept, (size_t)E * sizeof(int)}, {(char*)q_X, (const char*)src_X, bytes_X}};
            // page-locked caller arrays (corb_host_alloc: what the C++ shim flattens into) are read by the DMA engines in place;
            // pageable ones go through the staging slots
            cudaError_t e0 = cudaSetDevice(device);
            std::vector<Job> chunks;
            for (const Job& j : jobs) {
                cudaPointerAttributes at;
                const bool locked = cudaPointerGetAttributes(&at, j.src) == cudaSuccess && at.type == cudaMemoryTypeHost;
                if (!locked) cudaGetLastError();
                if (locked || !stage) {
                    if (e0 == cudaSuccess) e0 = cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyHostToDevice, st);
                    continue;
                }
                for (size_t o = 0; o < j.bytes; o += BaArena::kStageBytes)
                    chunks.push_back({j.dst + o, j.src + o, std::min(BaArena::kStageBytes, j.bytes - o)});
            }
            if (chunks.empty() || e0 != cudaSuccess) {
                up_err = e0;
                return;
            }
            cudaError_t errs[BaArena::kStageSlots];
            std::vector<std::thread> th;
            for (int t = 0; t < BaArena::kStageSlots; t++)
                th.emplace_back([&, t] {
                    cudaError_t e = cudaSetDevice(device);
                    uint8_t* slot = stage->h_stage + (size_t)t * BaArena::kStageBytes;
                    bool used = false;
                    for (size_t k = t; k < chunks.size() && e == cudaSuccess; k += BaArena::kStageSlots) {
                        if (used) e = cudaEventSynchronize(stage->stage_ev[t]);  // the slot's previous DMA has read it
                        if (e != cudaSuccess) break;
                        memcpy(slot, chunks[k].src, chunks[k].bytes);
                        e = cudaMemcpyAsync(chunks[k].dst, slot, chunks[k].bytes, cudaMemcpyHostToDevice, st);
                        if (e == cudaSuccess) e = cudaEventRecord(stage->stage_ev[t], st);
                        used = true;
                    }
                    errs[t] = e;
                });
            for (auto& x : th) x.join();
            cudaError_t e = cudaSuccess;
            for (cudaError_t x : errs) if (e == cudaSuccess) e = x;
            up_err = e;
        });
    }
    // ---- ordering + envelope. nbr_min/nbr_max[j] = smallest / largest free pose sharing a free landmark with j.
    //      Long-range couplings (loop closures, map-fusion links) would stretch the envelope of every row they touch
    //      back to the far keyframe; instead the smaller of the two vertex covers of the long links is ordered last
    //      ("bordered band"): the remaining keyframes keep a narrow band and only the few border rows are long.
    int rc;
    double* d_tmp = nullptr;
    if (allreduce && Pf > 0 && (rc = H.alloc(&d_tmp, 2 * (size_t)Pf)) != CORB_OK) return rc;
    // device form of the neighbour ranges (k_ba_nbr_range): needs the edges' keyframe indices, already on their way up
    const bool dev_ranges = early_upload && Pf > 0 && L > 0;
    int *d_idx = nullptr, *d_nbr = nullptr;
    std::vector<int> h_nbr;
    if (dev_ranges) {
        int *q_lmoff, *q_lfree;
        if ((rc = H.alloc(&q_lmoff, (size_t)L + 1)) != CORB_OK || (rc = H.alloc(&q_lfree, (size_t)L)) != CORB_OK ||
            (rc = H.alloc(&d_idx, (size_t)P)) != CORB_OK || (rc = H.alloc(&d_nbr, 2 * (size_t)Pf)) != CORB_OK)
            return rc;
        CORB_CUDA(cudaMemcpyAsync(q_lmoff, lm_off.data(), ((size_t)L + 1) * sizeof(int), cudaMemcpyHostToDevice, H.stream));
        CORB_CUDA(cudaMemcpyAsync(q_lfree, lfree.data(), (size_t)L * sizeof(int), cudaMemcpyHostToDevice, H.stream));
        d.lm_off = q_lmoff;
        d.lfree = q_lfree;
        h_nbr.resize(2 * (size_t)Pf);
    }
    auto neighbour_range = [&](const std::vector<int>& idx, std::vector<double>& mn, std::vector<double>& mx) -> int {
        mn.assign(std::max(Pf, 1), 0.0);
        mx.assign(std::max(Pf, 1), 0.0);
        if (dev_ranges) {
            CORB_CUDA(cudaMemcpyAsync(d_idx, idx.data(), (size_t)P * sizeof(int), cudaMemcpyHostToDevice, H.stream));
            k_ba_nbr_init<<<(Pf + 255) / 256, 256, 0, H.stream>>>(d_nbr, Pf);
            k_ba_nbr_range<<<(L + 255) / 256, 256, 0, H.stream>>>(d.e_pose, d.lm_off, d.lfree, d_idx, L, Pf, d_nbr);
            CORB_CUDA(cudaGetLastError());
            if (allreduce) {  // all ranks must agree on the structure of the reduced system
                k_ba_int_to_double<<<(2 * Pf + 255) / 256, 256, 0, H.stream>>>(d_nbr, d_tmp, 2 * Pf, 0, nullptr);
                int r2 = H.reduce(d_tmp, Pf, 1);
                if (r2 != CORB_OK) return r2;
                r2 = H.reduce(d_tmp + Pf, Pf, 2);
                if (r2 != CORB_OK) return r2;
                CORB_CUDA(cudaMemcpyAsync(mn.data(), d_tmp, Pf * sizeof(double), cudaMemcpyDeviceToHost, H.stream));
                CORB_CUDA(cudaMemcpyAsync(mx.data(), d_tmp + Pf, Pf * sizeof(double), cudaMemcpyDeviceToHost, H.stream));
                CORB_CUDA(cudaStreamSynchronize(H.stream));
            } else {
                CORB_CUDA(cudaMemcpyAsync(h_nbr.data(), d_nbr, 2 * (size_t)Pf * sizeof(int), cudaMemcpyDeviceToHost, H.stream));
                CORB_CUDA(cudaStreamSynchronize(H.stream));
                for (int j = 0; j < Pf; j++) { mn[j] = h_nbr[j]; mx[j] = h_nbr[Pf + j]; }
            }
            return CORB_OK;
        }
        for (int j = 0; j < Pf; j++) mn[j] = mx[j] = j;
        std::mutex merge;
        host_parallel_for(L, [&](int l0, int l1) {
            std::vector<double> tmn(mn), tmx(mx);  // min / max commute: per-thread copies, merged at the end
            for (int l = l0; l < l1; l++) {
                if (lfree[l] < 0) continue;
                int lo = Pf, hi = -1;
                for (int k = lm_off[l]; k < lm_off[l + 1]; k++) {
                    const int pj = idx[ep[k]];
                    if (pj >= 0) { lo = std::min(lo, pj); hi = std::max(hi, pj); }
                }
                for (int k = lm_off[l]; k < lm_off[l + 1]; k++) {
                    const int pj = idx[ep[k]];
                    if (pj >= 0) { tmn[pj] = std::min(tmn[pj], (double)lo); tmx[pj] = std::max(tmx[pj], (double)hi); }
                }
            }
            std::lock_guard<std::mutex> lk(merge);
            for (int j = 0; j < Pf; j++) { mn[j] = std::min(mn[j], tmn[j]); mx[j] = std::max(mx[j], tmx[j]); }
        });
        if (allreduce && Pf > 0) {  // all ranks must agree on the structure of the reduced system
            CORB_CUDA(cudaMemcpyAsync(d_tmp, mn.data(), Pf * sizeof(double), cudaMemcpyHostToDevice, H.stream));
            CORB_CUDA(cudaMemcpyAsync(d_tmp + Pf, mx.data(), Pf * sizeof(double), cudaMemcpyHostToDevice, H.stream));
            int r2 = H.reduce(d_tmp, Pf, 1);
            if (r2 != CORB_OK) return r2;
            r2 = H.reduce(d_tmp + Pf, Pf, 2);
            if (r2 != CORB_OK) return r2;
            CORB_CUDA(cudaMemcpyAsync(mn.data(), d_tmp, Pf * sizeof(double), cudaMemcpyDeviceToHost, H.stream));
            CORB_CUDA(cudaMemcpyAsync(mx.data(), d_tmp + Pf, Pf * sizeof(double), cudaMemcpyDeviceToHost, H.stream));
            CORB_CUDA(cudaStreamSynchronize(H.stream));
        }
        return CORB_OK;
    };
    std::vector<double> firstd, lastd;
    if ((rc = neighbour_range(pfree, firstd, lastd)) != CORB_OK) return rc;
    H.n_band = Pf;
    {
        const int T = 64;  // links spanning more than T keyframes count as long-range
        std::vector<int> up, down;
        for (int j = 0; j < Pf; j++) {
            if (lastd[j] - j > T) up.push_back(j);
            if (j - firstd[j] > T) down.push_back(j);
        }
        const std::vector<int>& border = up.size() <= down.size() ? up : down;
        if (!border.empty() && (int)border.size() < Pf) {
            std::vector<uint8_t> is_border(Pf, 0);
            for (int j : border) is_border[j] = 1;
            std::vector<int> newidx(Pf);
            int nxt = 0;
            for (int j = 0; j < Pf; j++) if (!is_border[j]) newidx[j] = nxt++;
            for (int j = 0; j < Pf; j++) if (is_border[j]) newidx[j] = nxt++;
            for (int i = 0; i < P; i++) if (pfree[i] >= 0) pfree[i] = newidx[pfree[i]];
            if ((rc = neighbour_range(pfree, firstd, lastd)) != CORB_OK) return rc;
            res->border_poses = (int)border.size();
            H.n_band = Pf - (int)border.size();
        }
    }
    // ---- chunks. The band factorisation is a serial sweep (each block column waits for the previous one). Taking
    //      wmax consecutive keyframes out of the band every n/Q columns ("separators", wmax = band half width) leaves Q
    //      chunks that share no landmark, i.e. Q independent band factorisations, one CTA each. The separators join
    //      the border rows: ordered last (before the long-range rows), swept against the chunks by k_ba_border_rows -
    //      per (row, chunk), also concurrently - and factored in the small dense border block.
    std::vector<int> chunk_start(2, 0);
    chunk_start[1] = H.n_band;
    {
        int wmax = 0;
        for (int j = 0; j < H.n_band; j++) wmax = std::max(wmax, j - (int)firstd[j]);
        const char* qenv = getenv("CORB_BA_CHUNKS");
        // chunks of ~120 band columns: measured on B200 at P = 2000 with the tiled border solver: Q = 8 -> 2.2, 12 -> 1.8, 16 -> 1.5,
        // 24 -> 1.5, 31 -> 1.9 ms per solve (more chunks shorten the serial sweeps and grow the dense border block)
        int Q = qenv ? atoi(qenv) : std::min(24, H.n_band / 120);
        Q = std::min(Q, H.n_band / 64);
        if (wmax >= 1 && wmax <= 8 && Q >= 2 && H.n_band - (Q - 1) * wmax >= Q) {
            const int nb0 = H.n_band, nbord0 = Pf - nb0;
            std::vector<uint8_t> is_sep(nb0, 0);
            for (int q = 1; q < Q; q++) {
                const int pq = (int)((long long)q * nb0 / Q);
                for (int j = pq - wmax; j < pq; j++) is_sep[j] = 1;
            }
            std::vector<int> newidx(Pf);
            int nxt = 0;
            chunk_start.assign(1, 0);
            for (int j = 0; j < nb0; j++) {
                if (is_sep[j]) {
                    if (j + 1 == nb0 || !is_sep[j + 1]) chunk_start.push_back(nxt);  // a chunk ends in front of each separator
                    continue;
                }
                newidx[j] = nxt++;
            }
            chunk_start.push_back(nxt);
            const int nb1 = nxt;
            for (int j = 0; j < nb0; j++) if (is_sep[j]) newidx[j] = nxt++;
            for (int j = nb0; j < Pf; j++) newidx[j] = nxt++;
            for (int i = 0; i < P; i++) if (pfree[i] >= 0) pfree[i] = newidx[pfree[i]];
            if ((rc = neighbour_range(pfree, firstd, lastd)) != CORB_OK) return rc;
            H.n_band = nb1;
            res->separator_poses = nb0 - nb1;
            res->border_poses = nbord0 + res->separator_poses;
        }
    }
    H.n_chunks = (int)chunk_start.size() - 1;
    res->band_chunks = H.n_chunks;
    lap("ordering (border, chunks)");
    std::vector<int> first(Pf), rowoff(Pf + 1, 0), coloff(Pf + 1, 0);
    for (int j = 0; j < Pf; j++) {
        // a border row stores every border column to its left (the tiled dense factorisation of the border block addresses
        // all of its lower triangle); its band envelope is what the couplings say
        first[j] = j >= H.n_band ? std::min((int)firstd[j], H.n_band) : (int)firstd[j];
        rowoff[j + 1] = rowoff[j] + (j - first[j] + 1);
    }
    const long long nblocks = Pf ? rowoff[Pf] : 0;
    CORB_CHECK(nblocks < (1LL << 31) / 36 * 8, CORB_ERR_CAPACITY, "reduced camera system envelope too large (%lld blocks)", nblocks);
    for (int j = 0; j < Pf; j++)
        for (int k = first[j]; k < j; k++) coloff[k + 1]++;
    for (int k = 0; k < Pf; k++) coloff[k + 1] += coloff[k];
    std::vector<int> col_rows(std::max(1, Pf ? coloff[Pf] : 0));
    {
        std::vector<int> cur(coloff.begin(), coloff.end() - 1);
        for (int j = 0; j < Pf; j++)
            for (int k = first[j]; k < j; k++) col_rows[cur[k]++] = j;  // ascending j within a column
    }
    {
        int wmax = 1;
        for (int j = 0; j < H.n_band; j++) wmax = std::max(wmax, j - first[j]);
        H.band_wmax = wmax;
        const size_t idx = (size_t)2 * H.n_band * sizeof(int);
        H.band_idx_bytes = (3 * wmax + 1) * 36 * sizeof(double) + idx <= 180 * 1024 ? idx : 0;
        CORB_SMEM_OPT_IN(k_ba_border_rows);
        CORB_SMEM_OPT_IN(k_ba_band_backward);
        const size_t nbord = (size_t)(Pf - H.n_band);
        const size_t dense_bytes = ((nbord + 1) * 36 + nbord * 6 + nbord * 36 + (size_t)32 * kDenseU * 36) * sizeof(double);
        H.border_dense_bytes = dense_bytes;
        H.border_dense_ok = nbord > 0 && dense_bytes <= 200 * 1024;
        if (H.border_dense_ok)
            CORB_SMEM_OPT_IN(k_ba_border_dense);
        if (nbord > 0 && !getenv("CORB_BA_BORDER_OLD")) {
            const int nt = ((int)nbord + kBTB - 1) / kBTB;
            H.border_tiled_bytes = ((size_t)2 * kBT * kBLd + kBT + (size_t)nt * kBT + (size_t)nt * kBTB) * sizeof(double);
            int dev = 0, sms = 0, coop = 0, per_sm = 0;
            CORB_CUDA(cudaGetDevice(&dev));
            CORB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            CORB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
            CORB_SMEM_OPT_IN(k_ba_border_chol);
            CORB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_border_chol, kBcThreads, H.border_tiled_bytes));
            H.border_tiled = coop && per_sm > 0 && H.border_tiled_bytes <= 200 * 1024;
            H.border_grid = std::max(1, std::min(sms * per_sm, std::max(nt - 1, nt * (nt - 1) / 2)));
        }
    }
    // the same lists without the border rows in the band columns: the band sweep of the bordered solve
    std::vector<int> coloff_b(Pf + 1, 0), col_rows_b(col_rows.size());
    {
        int w = 0;
        for (int k = 0; k < Pf; k++) {
            coloff_b[k] = w;
            for (int e = coloff[k]; e < coloff[k + 1]; e++)
                if (k >= H.n_band || col_rows[e] < H.n_band) col_rows_b[w++] = col_rows[e];
        }
        if (Pf) coloff_b[Pf] = w;
    }
    res->reduced_blocks = nblocks;
    {
        int max_nact = 0;
        for (int k = 0; k < Pf; k++) max_nact = std::max(max_nact, coloff[k + 1] - coloff[k]);
        H.lact_cap = std::max(1, std::min(max_nact, 700));
        const size_t lact_bytes = (size_t)H.lact_cap * 36 * sizeof(double);
        const size_t idx_bytes = ((size_t)3 * Pf + 1 + (Pf ? coloff_b[Pf] : 0)) * sizeof(int);
        if (lact_bytes + idx_bytes <= 200 * 1024) {
            H.stage_nnz_b = Pf ? coloff_b[Pf] : 0;
            H.stage_bytes_b = idx_bytes;
        }
        CORB_SMEM_OPT_IN(k_ba_solve);
        res->max_active_rows = max_nact;
    }
    H.s_doubles = (size_t)nblocks * 36;
    lap("envelope lists");
    // ---- device buffers
#define UP(field, vec) if ((rc = H.upload(&d.field, vec)) != CORB_OK) return rc
    UP(pfree, pfree);
    if (!dev_ranges) { UP(lfree, lfree); }
    if (early_upload) {
        uploader.t.join();
        CORB_CHECK(up_err == cudaSuccess, CORB_ERR_CUDA, "uploading the edge arrays failed: %s", cudaGetErrorString(up_err));
        if (up_st != H.stream) {
            CORB_CUDA(cudaEventRecord(H.arena->ev_up, up_st));
            CORB_CUDA(cudaStreamWaitEvent(H.stream, H.arena->ev_up, 0));
        }
    } else if ((rc = H.upload_raw(&d.e_pose, ep, (size_t)E)) != CORB_OK || (rc = H.upload_raw(&d.e_point, ept, (size_t)E)) != CORB_OK ||
               (rc = H.upload_raw(&d.e_obs, eobs, (size_t)E * 3)) != CORB_OK || (rc = H.upload_raw(&d.e_info, einfo, (size_t)E)) != CORB_OK) {
        return rc;
    }
    lap("edge upload joined");
    if ((rc = H.build_pose_csr()) != CORB_OK) return rc;  // CSR by keyframe, built on the device from the uploaded e_pose
    lap("pose CSR enqueued");
    if (!dev_ranges) { UP(lm_off, lm_off); }
    UP(first, first); UP(rowoff, rowoff);
    UP(coloff, coloff); UP(col_rows, col_rows); UP(coloff_b, coloff_b); UP(col_rows_b, col_rows_b); UP(chunk_start, chunk_start);
#undef UP
    {
        std::vector<double> cam(p->pose_cam, p->pose_cam + (size_t)P * 5);
        if ((rc = H.upload(&d.cam, cam)) != CORB_OK) return rc;
    }
#define AL(field, n) if ((rc = H.alloc(&d.field, (n))) != CORB_OK) return rc
    AL(q, (size_t)P * 4); AL(t, (size_t)P * 3);
    if (!early_upload) { AL(X, (size_t)L * 3); }
    AL(q0, (size_t)P * 4); AL(t0, (size_t)P * 3); AL(X0, (size_t)L * 3);
    AL(Hpp, (size_t)Pf * 36); AL(bp, (size_t)Pf * 6); AL(Hll, (size_t)L * 9); AL(bl, (size_t)L * 3); AL(W, (size_t)E * 18);
    AL(Dinv, (size_t)L * 9); AL(db, (size_t)L * 3);
    AL(S, H.s_doubles + (size_t)Pf * 6);
    AL(xp, (size_t)Pf * 6); AL(xl, (size_t)L * 3); AL(invd, (size_t)Pf * 6 + 6);
    AL(bpart, (size_t)std::max(1, Pf - H.n_band) * H.n_chunks * 6);
    {
        const int nbs = std::max(1, Pf - H.n_band) * H.n_chunks;
        int* bstart = nullptr;
        if ((rc = H.alloc(&bstart, (size_t)nbs)) != CORB_OK) return rc;
        k_ba_fill_int<<<(nbs + 255) / 256, 256, 0, H.stream>>>(bstart, nbs, INT_MAX);
        d.bstart = bstart;
        if (Pf > H.n_band && L > 0) {
            k_ba_border_starts<<<(L + 255) / 256, 256, 0, H.stream>>>(d, H.n_band, H.n_chunks, bstart);
            if (allreduce) {  // every rank sees its own landmarks only: the structure is the union (minimum) over the ranks
                double* tmp = nullptr;
                if ((rc = H.alloc(&tmp, (size_t)nbs)) != CORB_OK) return rc;
                k_ba_int_to_double<<<(nbs + 255) / 256, 256, 0, H.stream>>>(bstart, tmp, nbs, 0, nullptr);
                if ((rc = H.reduce(tmp, (size_t)nbs, 1)) != CORB_OK) return rc;
                k_ba_int_to_double<<<(nbs + 255) / 256, 256, 0, H.stream>>>(nullptr, tmp, nbs, 1, bstart);
            }
        }
        CORB_CUDA(cudaGetLastError());
    }
    {
        const size_t nb = (size_t)(Pf - H.n_band);
        AL(syrk_part, std::max<size_t>(1, nb * (nb + 1) / 2 * H.n_chunks * 36));
    }
    const size_t npart = (size_t)std::max(std::max((E + 255) / 256, (L * 3 + 255) / 256), 2048) * 2 + 16;
    AL(partial, npart); AL(scalars, 8);
#undef AL
    {
        size_t pairs_bound = 0;
        for (int l = 0; l < L; l++) {
            const size_t k = (size_t)(lm_off[l + 1] - lm_off[l]);
            pairs_bound += k * (k + 1) / 2;
        }
        lap("buffers allocated");
        if ((rc = H.build_schur_pairs(pairs_bound, (size_t)nblocks)) != CORB_OK) return rc;
        lap("pair lists built (sync)");
        res->schur_pair_lists = H.sp_ok ? 1 : 0;
    }
    d.bs = d.S + H.s_doubles;
    CORB_CUDA(cudaMemcpyAsync(d.q, p->pose_q, (size_t)P * 4 * sizeof(double), cudaMemcpyHostToDevice, H.stream));
    CORB_CUDA(cudaMemcpyAsync(d.t, p->pose_t, (size_t)P * 3 * sizeof(double), cudaMemcpyHostToDevice, H.stream));
    if (!early_upload) CORB_CUDA(cudaMemcpyAsync(d.X, p->point_xyz, (size_t)L * 3 * sizeof(double), cudaMemcpyHostToDevice, H.stream));
    CORB_CUDA(cudaMemsetAsync(d.scalars, 0, 8 * sizeof(double), H.stream));

    lap("uploads + allocations");
    res->ms_setup = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    // ---- Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:61-164, sparse_optimizer.cpp:354-419)
    auto terminate = [&]() { return H.stop_requested(); };
    double lambda = -1, ni = 2;
    int nBad = 0, it = 0;
    bool ok = true;
    for (; it < iterations && !terminate() && ok; it++) {
        double currentChi;
        if ((rc = H.chi2(&currentChi)) != CORB_OK) return rc;
        const double iniChi = currentChi;
        if (it == 0) res->chi2_initial = currentChi;
        if (H.ar && terminate()) {  // the ranks agreed on the flag in the reduction above (first poll of a sharded solve)
            if (it == 0) res->chi2_final = currentChi;
            break;
        }
        if ((rc = H.build()) != CORB_OK) return rc;
        if (it == 0) {
            if ((rc = H.lambda_init(&lambda)) != CORB_OK) return rc;
            ni = 2; nBad = 0;
            res->lambda_initial = lambda;
        }
        double rho = 0;
        int qmax = 0;
        do {
            // push: keep the accepted estimate in the backup buffers
            CORB_CUDA(cudaMemcpyAsync(d.q0, d.q, (size_t)P * 4 * sizeof(double), cudaMemcpyDeviceToDevice, H.stream));
            CORB_CUDA(cudaMemcpyAsync(d.t0, d.t, (size_t)P * 3 * sizeof(double), cudaMemcpyDeviceToDevice, H.stream));
            CORB_CUDA(cudaMemcpyAsync(d.X0, d.X, (size_t)L * 3 * sizeof(double), cudaMemcpyDeviceToDevice, H.stream));
            bool ok2 = true;
            double scale = 0;
            if ((rc = H.trial(lambda, &ok2, &scale)) != CORB_OK) return rc;
            if (!ok2) res->solver_failures++;
            double tempChi;
            if ((rc = H.chi2(&tempChi)) != CORB_OK) return rc;
            if (!ok2) tempChi = DBL_MAX;
            rho = (currentChi - tempChi) / (scale + 1e-3);
            const bool good = rho > 0 && std::isfinite(tempChi);
            if (res->n_trials < 256) { res->trial_accepted[res->n_trials] = good; res->trial_chi2[res->n_trials] = tempChi; }
            res->n_trials++;
            if (good) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                // pop
                CORB_CUDA(cudaMemcpyAsync(d.q, d.q0, (size_t)P * 4 * sizeof(double), cudaMemcpyDeviceToDevice, H.stream));
                CORB_CUDA(cudaMemcpyAsync(d.t, d.t0, (size_t)P * 3 * sizeof(double), cudaMemcpyDeviceToDevice, H.stream));
                CORB_CUDA(cudaMemcpyAsync(d.X, d.X0, (size_t)L * 3 * sizeof(double), cudaMemcpyDeviceToDevice, H.stream));
            }
            qmax++;
        } while (rho < 0 && qmax < 10 && !terminate());
        res->chi2_final = currentChi;
        if (qmax == 10 || rho == 0) { ok = false; it++; break; }
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) { ok = false; it++; break; }
    }
    res->iterations = it;
    res->lambda_final = lambda;
    res->stopped = terminate() ? 1 : 0;
    CORB_CUDA(cudaMemcpyAsync(p->pose_q, d.q, (size_t)P * 4 * sizeof(double), cudaMemcpyDeviceToHost, H.stream));
    CORB_CUDA(cudaMemcpyAsync(p->pose_t, d.t, (size_t)P * 3 * sizeof(double), cudaMemcpyDeviceToHost, H.stream));
    CORB_CUDA(cudaMemcpyAsync(p->point_xyz, d.X, (size_t)L * 3 * sizeof(double), cudaMemcpyDeviceToHost, H.stream));
    CORB_CUDA(cudaStreamSynchronize(H.stream));
    res->ms_solve = H.ms_solve;
    res->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    return res->stopped ? CORB_ERR_STOPPED : CORB_OK;
}
