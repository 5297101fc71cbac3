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
#include <chrono>
#include <numeric>
#include <vector>

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
    double *S, *bs;                // reduced system: [S | bs] contiguous
    double *xp, *xl;
    double *partial, *scalars;     // reduction scratch; scalars[0] chi2, [1] scale part, [2] xx, [3] max diag, [4] fail flag
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
__global__ void __launch_bounds__(256) k_ba_schur_rows(BaDev d) {
    const int pi = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (pi >= d.P) return;
    const int j = d.pfree[pi];
    if (j < 0) return;
    const int lane = threadIdx.x & 31;
    const int fj = d.first[j];
    double* row = d.S + (size_t)d.rowoff[j] * 36;
    const int nrow = (j - fj + 1) * 36;
    for (int i = lane; i < nrow; i += 32) row[i] = 0.0;
    __syncwarp();
    double* diag = row + (size_t)(j - fj) * 36;
    diag[lane] = d.Hpp[(size_t)j * 36 + lane];
    if (lane < 4) diag[32 + lane] = d.Hpp[(size_t)j * 36 + 32 + lane];
    __syncwarp();
    const int r0 = lane / 6, c0 = lane - r0 * 6;  // entry `lane`
    const int c1 = 2 + lane;                      // entry 32 + lane = (5, 2 + lane) for lane < 4
    double coeff = 0.0;
    for (int k = d.pose_off[pi]; k < d.pose_off[pi + 1]; k++) {
        const int e = d.pose_edges[k];
        const int l = d.e_point[e];
        if (d.lfree[l] < 0) continue;
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
        for (int e2 = d.lm_off[l]; e2 < d.lm_off[l + 1]; e2++) {
            const int j2 = d.pfree[d.e_pose[e2]];
            if (j2 < 0 || j2 > j) continue;
            const double* W2 = d.W + (size_t)e2 * 18;
            double* blk = row + (size_t)(j2 - fj) * 36;
            blk[lane] -= bd0[0] * W2[c0 * 3] + bd0[1] * W2[c0 * 3 + 1] + bd0[2] * W2[c0 * 3 + 2];
            if (lane < 4) blk[32 + lane] -= bd1[0] * W2[c1 * 3] + bd1[1] * W2[c1 * 3 + 1] + bd1[2] * W2[c1 * 3 + 2];
        }
    }
    if (lane < 6) d.bs[(size_t)j * 6 + lane] = d.bp[(size_t)j * 6 + lane] - coeff;
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
constexpr int kSolveThreads = 1024;
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
__device__ __forceinline__ bool warp_chol6(double* D, const double* L0, double* Lout, int lane) {
    int r = 0, c = 0;
    if (lane < 21) {
        r = lane >= 15 ? 5 : lane >= 10 ? 4 : lane >= 6 ? 3 : lane >= 3 ? 2 : lane >= 1 ? 1 : 0;
        c = lane - r * (r + 1) / 2;
    }
    double a = lane < 21 ? D[r * 6 + c] : 0.0;
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
        const double dk = sqrt(akk), inv = 1.0 / dk;
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

__global__ void __launch_bounds__(kSolveThreads) k_ba_solve(BaDev d, int lact_cap, int k_begin, int k_end, int n_band, int flags) {
    // flags: 1 = first launch (load b, clear the failure flag), 2 = defer border x border updates to k_ba_border_syrk,
    //        4 = run the backward substitution after the last column
    extern __shared__ double Lact[];  // [lact_cap][36] staged L_jk of the active rows of the current column
    __shared__ double Lkk[2][36];
    __shared__ double yk[6];
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = d.Pf;
    if (tid == 0) s_fail = (flags & 1) ? 0 : (d.scalars[4] != 0.0);
    if (flags & 1)
        for (int i = tid; i < n * 6; i += kSolveThreads) d.xp[i] = d.bs[i];
    __syncthreads();
    if (warp == 0 && k_begin < k_end && !s_fail) {
        double* D = d.S + (size_t)(d.rowoff[k_begin] + k_begin - d.first[k_begin]) * 36;
        if (!warp_chol6(D, nullptr, Lkk[k_begin & 1], lane) && lane == 0) s_fail = 1;
    }
    __syncthreads();
    for (int k = k_begin; k < k_end; k++) {
        if (s_fail) break;
        const double* Lk = Lkk[k & 1];
        const int cb = d.coloff[k], nact = d.coloff[k + 1] - cb;
        const int* rows = d.col_rows + cb;
        const bool staged = nact <= lact_cap;
        // ---- phase B
        for (int it = tid; it < nact * 6; it += kSolveThreads) {
            const int a = it / 6, r = it - a * 6;
            const int j = rows[a];
            double* B = d.S + (size_t)(d.rowoff[j] + k - d.first[j]) * 36 + r * 6;
            double v[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double s = B[c];
#pragma unroll
                for (int p = 0; p < c; p++) s -= v[p] * Lk[c * 6 + p];
                v[c] = s / Lk[c * 6 + c];
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
                v[r] = s / Lk[r * 6 + r];
            }
#pragma unroll
            for (int r = 0; r < 6; r++) { d.xp[k * 6 + r] = v[r]; yk[r] = v[r]; }
        }
        __syncthreads();
        // ---- phase C
        const bool next_active = nact > 0 && rows[0] == k + 1;  // rows ascend, so k+1 can only be the first entry
        if (warp == 0) {
            if (k + 1 < k_end) {
                double* D = d.S + (size_t)(d.rowoff[k + 1] + (k + 1) - d.first[k + 1]) * 36;
                const double* L0 = !next_active ? nullptr : staged ? Lact : d.S + (size_t)(d.rowoff[k + 1] + k - d.first[k + 1]) * 36;
                if (!warp_chol6(D, L0, Lkk[(k + 1) & 1], lane) && lane == 0) s_fail = 1;
            }
        } else if (warp == 1) {  // b_j -= L_jk y_k
            for (int it = lane; it < nact * 6; it += 32) {
                const int a = it / 6, r = it - a * 6;
                const int j = rows[a];
                const double* Lr = staged ? Lact + a * 36 + r * 6 : d.S + (size_t)(d.rowoff[j] + k - d.first[j]) * 36 + r * 6;
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
                        toff[u] = (d.rowoff[j] + rows[b] - d.first[j]) * 36;
                        t0[u] = d.S[(size_t)toff[u] + lane];
                        t1[u] = lane < 4 ? d.S[(size_t)toff[u] + 32 + lane] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (toff[u] >= 0) {
                        const int ja = rows[pa[u]], jb = rows[pb[u]];
                        const double* La = staged ? Lact + pa[u] * 36 : d.S + (size_t)(d.rowoff[ja] + k - d.first[ja]) * 36;
                        const double* Lb = staged ? Lact + pb[u] * 36 : d.S + (size_t)(d.rowoff[jb] + k - d.first[jb]) * 36;
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
    }
    if (s_fail) {
        for (int i = tid; i < n * 6; i += kSolveThreads) d.xp[i] = 0.0;
        if (tid == 0) d.scalars[4] = 1.0;
        return;
    }
    if (tid == 0) d.scalars[4] = 0.0;
    if (!(flags & 4)) return;
    // ---- backward: L^T x = y (row oriented); xp already holds y from the fused forward pass
    for (int j = n - 1; j >= 0; j--) {
        const int fj = d.first[j];
        const double* rowp = d.S + (size_t)d.rowoff[j] * 36;
        const double* D = rowp + (size_t)(j - fj) * 36;
        if (tid == 0) {
            double v[6];
#pragma unroll
            for (int r = 5; r >= 0; r--) {
                double s = d.xp[j * 6 + r];
#pragma unroll
                for (int p = r + 1; p < 6; p++) s -= D[p * 6 + r] * v[p];
                v[r] = s / D[r * 6 + r];
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
}

// ---- K12b: deferred update of the border block (keyframes with long-range links, ordered last):
//      S(j, i) -= sum_k L_jk L_ik^T over the band columns k shared by the envelopes of border rows j >= i.
//      One warp per (j, i) pair; lane = entry of the 6x6 block; fixed k order => deterministic.
__global__ void __launch_bounds__(256) k_ba_border_syrk(BaDev d, int n_band) {
    const int nbord = d.Pf - n_band;
    const int p = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (p >= nbord * (nbord + 1) / 2) return;
    const int lane = threadIdx.x & 31;
    int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > p) a--;
    while ((a + 1) * (a + 2) / 2 <= p) a++;
    const int b = p - a * (a + 1) / 2;
    const int j = n_band + a, i = n_band + b;
    const int k0 = max(d.first[j], d.first[i]);
    const double* Lj = d.S + (size_t)(d.rowoff[j] - d.first[j]) * 36;
    const double* Li = d.S + (size_t)(d.rowoff[i] - d.first[i]) * 36;
    const int r0 = lane / 6, c0 = lane - r0 * 6, c1 = 2 + lane;
    double s0 = 0, s1 = 0;
    for (int k = k0; k < n_band; k++) {
        const double* A = Lj + (size_t)k * 36;
        const double* B = Li + (size_t)k * 36;
#pragma unroll
        for (int q = 0; q < 6; q++) s0 += A[r0 * 6 + q] * B[c0 * 6 + q];
        if (lane < 4) {
#pragma unroll
            for (int q = 0; q < 6; q++) s1 += A[30 + q] * B[c1 * 6 + q];
        }
    }
    double* T = d.S + (size_t)(d.rowoff[j] + i - d.first[j]) * 36;
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

using namespace corb;

namespace {

struct BaHost {
    int device;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    BaDev d;
    size_t s_doubles = 0;  // S blocks * 36
    int lact_cap = 0;      // active rows of one column that fit the solve kernel's shared memory
    int n_band = 0;        // free keyframes before the border block (== Pf when there is no border)
    corb_allreduce_fn ar = nullptr;
    void* ar_user = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms_solve = 0;
    double* h_scalars = nullptr;  // pinned [8]

    ~BaHost() {
        for (void* p : allocs) cudaFree(p);
        if (h_scalars) cudaFreeHost(h_scalars);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
    template <typename T>
    int alloc(T** p, size_t n) {
        void* q = nullptr;
        CORB_CUDA(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T) + 256));
        allocs.push_back(q);
        *p = (T*)q;
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
        k_reduce_final<false><<<1, 256, 0, stream>>>(d.partial, nb, d.scalars);
        int rc = reduce(d.scalars, 1, 0);
        if (rc != CORB_OK) return rc;
        rc = read_scalars(1);
        *out = h_scalars[0];
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
        if (d.P > 0) k_ba_schur_rows<<<(d.P * 32 + 255) / 256, 256, 0, stream>>>(d);
        int rc = reduce(d.S, s_doubles + (size_t)d.Pf * 6, 0);
        if (rc != CORB_OK) return rc;
        if (d.Pf > 0) k_ba_add_lambda<<<(d.Pf * 6 + 255) / 256, 256, 0, stream>>>(d, lambda);
        cudaEventRecord(ev0, stream);
        {   // band columns (border x border updates deferred) -> border SYRK on the whole GPU -> border columns + backward
            const size_t sm = (size_t)lact_cap * 36 * sizeof(double);
            const int nbord = d.Pf - n_band;
            if (nbord > 0) {
                k_ba_solve<<<1, kSolveThreads, sm, stream>>>(d, lact_cap, 0, n_band, n_band, 1 | 2);
                const int npairs = nbord * (nbord + 1) / 2;
                k_ba_border_syrk<<<(npairs * 32 + 255) / 256, 256, 0, stream>>>(d, n_band);
                k_ba_solve<<<1, kSolveThreads, sm, stream>>>(d, lact_cap, n_band, d.Pf, n_band, 4);
            } else {
                k_ba_solve<<<1, kSolveThreads, sm, stream>>>(d, lact_cap, 0, d.Pf, n_band, 1 | 4);
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
    for (int e = 0; e < E; e++)
        CORB_CHECK(p->edge_pose[e] >= 0 && p->edge_pose[e] < P && p->edge_point[e] >= 0 && p->edge_point[e] < L, CORB_ERR_INVALID,
                   "edge %d references pose %d / point %d out of range", e, p->edge_pose[e], p->edge_point[e]);
    int ndev = 0;
    CORB_CUDA(cudaGetDeviceCount(&ndev));
    CORB_CHECK(device >= 0 && device < ndev, CORB_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    CORB_CUDA(cudaSetDevice(device));
    BaHost H;
    H.device = device;
    H.ar = allreduce;
    H.ar_user = allreduce_user;
    CORB_CUDA(cudaStreamCreateWithFlags(&H.stream, cudaStreamNonBlocking));
    CORB_CUDA(cudaEventCreate(&H.ev0));
    CORB_CUDA(cudaEventCreate(&H.ev1));
    CORB_CUDA(cudaMallocHost(&H.h_scalars, 8 * sizeof(double)));
    BaDev& d = H.d;
    memset(&d, 0, sizeof(d));
    d.P = P; d.L = L; d.E = E; d.robust = robust != 0;
    d.delta2d = (double)(float)sqrt(5.99);   // thHuber2D / thHuber3D are floats in the reference (Optimizer.cc:101-102)
    d.delta3d = (double)(float)sqrt(7.815);
    // ---- host-side structure (the analogue of g2o's buildStructure, block_solver.hpp:143-295)
    std::vector<int> pfree(P), lfree(L);
    int Pf = 0, Lf = 0;
    for (int i = 0; i < P; i++) pfree[i] = p->pose_fixed[i] ? -1 : Pf++;
    for (int i = 0; i < L; i++) lfree[i] = p->point_fixed[i] ? -1 : Lf++;
    d.Pf = Pf;
    std::vector<int> lm_off(L + 1, 0), perm(E);
    for (int e = 0; e < E; e++) lm_off[p->edge_point[e] + 1]++;
    for (int i = 0; i < L; i++) lm_off[i + 1] += lm_off[i];
    {
        std::vector<int> cur(lm_off.begin(), lm_off.end() - 1);
        for (int e = 0; e < E; e++) perm[cur[p->edge_point[e]]++] = e;
    }
    std::vector<int> e_pose(E), e_point(E);
    std::vector<double> e_obs((size_t)E * 3), e_info(E);
    for (int k = 0; k < E; k++) {
        const int e = perm[k];
        e_pose[k] = p->edge_pose[e]; e_point[k] = p->edge_point[e];
        e_obs[3 * (size_t)k] = p->edge_obs[3 * (size_t)e]; e_obs[3 * (size_t)k + 1] = p->edge_obs[3 * (size_t)e + 1];
        e_obs[3 * (size_t)k + 2] = p->edge_obs[3 * (size_t)e + 2];
        e_info[k] = p->edge_inv_sigma2[e];
    }
    std::vector<int> pose_off(P + 1, 0), pose_edges(E);
    for (int k = 0; k < E; k++) pose_off[e_pose[k] + 1]++;
    for (int i = 0; i < P; i++) pose_off[i + 1] += pose_off[i];
    {
        std::vector<int> cur(pose_off.begin(), pose_off.end() - 1);
        for (int k = 0; k < E; k++) pose_edges[cur[e_pose[k]]++] = k;
    }
    // ---- ordering + envelope. nbr_min/nbr_max[j] = smallest / largest free pose sharing a free landmark with j.
    //      Long-range couplings (loop closures, map-fusion links) would stretch the envelope of every row they touch
    //      back to the far keyframe; instead the smaller of the two vertex covers of the long links is ordered last
    //      ("bordered band"): the remaining keyframes keep a narrow band and only the few border rows are long.
    int rc;
    double* d_tmp = nullptr;
    if (allreduce && Pf > 0 && (rc = H.alloc(&d_tmp, 2 * (size_t)Pf)) != CORB_OK) return rc;
    auto neighbour_range = [&](const std::vector<int>& idx, std::vector<double>& mn, std::vector<double>& mx) -> int {
        mn.assign(std::max(Pf, 1), 0.0);
        mx.assign(std::max(Pf, 1), 0.0);
        for (int j = 0; j < Pf; j++) mn[j] = mx[j] = j;
        for (int l = 0; l < L; l++) {
            if (lfree[l] < 0) continue;
            int lo = Pf, hi = -1;
            for (int k = lm_off[l]; k < lm_off[l + 1]; k++) {
                const int pj = idx[e_pose[k]];
                if (pj >= 0) { lo = std::min(lo, pj); hi = std::max(hi, pj); }
            }
            for (int k = lm_off[l]; k < lm_off[l + 1]; k++) {
                const int pj = idx[e_pose[k]];
                if (pj >= 0) { mn[pj] = std::min(mn[pj], (double)lo); mx[pj] = std::max(mx[pj], (double)hi); }
            }
        }
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
    std::vector<int> first(Pf), rowoff(Pf + 1, 0), coloff(Pf + 1, 0);
    for (int j = 0; j < Pf; j++) {
        first[j] = (int)firstd[j];
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
    res->reduced_blocks = nblocks;
    {
        int max_nact = 0;
        for (int k = 0; k < Pf; k++) max_nact = std::max(max_nact, coloff[k + 1] - coloff[k]);
        H.lact_cap = std::max(1, std::min(max_nact, 700));
        CORB_CUDA(cudaFuncSetAttribute(k_ba_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, H.lact_cap * 36 * (int)sizeof(double)));
        res->max_active_rows = max_nact;
    }
    H.s_doubles = (size_t)nblocks * 36;
    // ---- device buffers
#define UP(field, vec) if ((rc = H.upload(&d.field, vec)) != CORB_OK) return rc
    UP(pfree, pfree); UP(lfree, lfree); UP(e_pose, e_pose); UP(e_point, e_point); UP(e_obs, e_obs); UP(e_info, e_info);
    UP(lm_off, lm_off); UP(pose_off, pose_off); UP(pose_edges, pose_edges); UP(first, first); UP(rowoff, rowoff);
    UP(coloff, coloff); UP(col_rows, col_rows);
#undef UP
    {
        std::vector<double> cam(p->pose_cam, p->pose_cam + (size_t)P * 5);
        if ((rc = H.upload(&d.cam, cam)) != CORB_OK) return rc;
    }
#define AL(field, n) if ((rc = H.alloc(&d.field, (n))) != CORB_OK) return rc
    AL(q, (size_t)P * 4); AL(t, (size_t)P * 3); AL(X, (size_t)L * 3); AL(q0, (size_t)P * 4); AL(t0, (size_t)P * 3); AL(X0, (size_t)L * 3);
    AL(Hpp, (size_t)Pf * 36); AL(bp, (size_t)Pf * 6); AL(Hll, (size_t)L * 9); AL(bl, (size_t)L * 3); AL(W, (size_t)E * 18);
    AL(Dinv, (size_t)L * 9); AL(db, (size_t)L * 3);
    AL(S, H.s_doubles + (size_t)Pf * 6);
    AL(xp, (size_t)Pf * 6); AL(xl, (size_t)L * 3);
    const size_t npart = (size_t)std::max(std::max((E + 255) / 256, (L * 3 + 255) / 256), 2048) * 2 + 16;
    AL(partial, npart); AL(scalars, 8);
#undef AL
    d.bs = d.S + H.s_doubles;
    CORB_CUDA(cudaMemcpyAsync(d.q, p->pose_q, (size_t)P * 4 * sizeof(double), cudaMemcpyHostToDevice, H.stream));
    CORB_CUDA(cudaMemcpyAsync(d.t, p->pose_t, (size_t)P * 3 * sizeof(double), cudaMemcpyHostToDevice, H.stream));
    CORB_CUDA(cudaMemcpyAsync(d.X, p->point_xyz, (size_t)L * 3 * sizeof(double), cudaMemcpyHostToDevice, H.stream));
    CORB_CUDA(cudaMemsetAsync(d.scalars, 0, 8 * sizeof(double), H.stream));

    // ---- Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:61-164, sparse_optimizer.cpp:354-419)
    auto terminate = [&]() { return stop && *stop; };
    double lambda = -1, ni = 2;
    int nBad = 0, it = 0;
    bool ok = true;
    for (; it < iterations && !terminate() && ok; it++) {
        double currentChi;
        if ((rc = H.chi2(&currentChi)) != CORB_OK) return rc;
        const double iniChi = currentChi;
        if (it == 0) res->chi2_initial = currentChi;
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
