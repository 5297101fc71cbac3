/* CPU ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Plain C++ restatement (no Eigen, no g2o) of the reference's global bundle adjustment:
 *   corbslam_client/src/Optimizer.cc:54-270                      graph construction / write-back (flattened by the caller)
 *   Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp}        EdgeSE3ProjectXYZ, EdgeStereoSE3ProjectXYZ, VertexSE3Expmap
 *   Thirdparty/g2o/g2o/types/se3quat.h:108-113,217-257           SE3Quat product / exp
 *   Thirdparty/g2o/g2o/types/types_sba.h:52-56                   VertexSBAPointXYZ::oplusImpl
 *   Thirdparty/g2o/g2o/core/base_binary_edge.hpp:55-120          constructQuadraticForm (+ Huber, robust_kernel_impl.cpp:78-91)
 *   Thirdparty/g2o/g2o/core/block_solver.hpp:354-486,502-604     Schur complement, back-substitution, lambda on the diagonals
 *   Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189   LM control flow
 *   Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:100-114,354-435 chi2, optimize loop, update
 *
 * The reduced camera system is solved with a block-skyline Cholesky (the reference uses Eigen's SimplicialLDLT with AMD
 * ordering, linear_solver_eigen.h:94-124 — Eigen is not vendored): same direct solution up to fp64 round-off.
 *
 * PARITY STATUS: parity unpinned (the only path left as a port: g2o and Optimizer.cc need Eigen3, which is not in the image,
 * so they cannot be compiled into oracle/_ref) — the reference ships no BA fixtures. Pinned by algebra instead: analytic
 * Jacobians vs central differences, chi2 monotone over accepted steps, LM constants from the source, and the converged minimum
 * against scipy.optimize.least_squares on an independently written residual (tests/test_oracle_cpu.py).
 *
 * Landmarks may be sharded over ranks: every rank holds all poses and a subset of landmarks with their edges; the
 * reduced system is summed through the caller's all-reduce callback.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

extern "C" {
typedef struct {
    int32_t n_poses, n_points, n_edges;
    double* pose_q;              /* [n_poses*4] x,y,z,w (world -> camera) */
    double* pose_t;              /* [n_poses*3] */
    const uint8_t* pose_fixed;   /* [n_poses] */
    const double* pose_cam;      /* [n_poses*5] fx fy cx cy bf */
    double* point_xyz;           /* [n_points*3] */
    const uint8_t* point_fixed;  /* [n_points] */
    const int32_t* edge_pose;    /* [n_edges] */
    const int32_t* edge_point;   /* [n_edges] */
    const double* edge_obs;      /* [n_edges*3] u v ur ; ur < 0 => monocular edge */
    const double* edge_inv_sigma2; /* [n_edges] */
} oracle_ba_problem;

typedef struct {
    int32_t iterations, n_trials, stopped, solver_failures;
    double chi2_initial, chi2_final, lambda_initial, lambda_final;
    uint8_t trial_accepted[256];
    double trial_chi2[256];
} oracle_ba_result;

/* op: 0 sum, 1 min, 2 max — in place over all ranks */
typedef int (*oracle_allreduce_fn)(void* user, double* buf, int n, int op);
}

namespace {

struct Mat3 { double m[9]; };

inline void quat_to_R(const double* q, double* R) { /* Eigen Quaternion::toRotationMatrix */
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
                 tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
inline void R_to_quat(const double* R, double* q) { /* Eigen quaternion from rotation matrix (Shepperd) */
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
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
}
inline void quat_normalize(double* q) { /* SE3Quat::normalizeRotation */
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
inline void quat_mul(const double* a, const double* b, double* o) { /* a * b, (x,y,z,w) */
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
/* T <- exp(delta) * T   (VertexSE3Expmap::oplusImpl, se3quat.h:217-257 and :108-113); delta = (omega, upsilon) */
void pose_oplus(double* q, double* t, const double* d) {
    const double om[3] = {d[0], d[1], d[2]}, up[3] = {d[3], d[4], d[5]};
    const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double O2[9], R[9], V[9];
    mat3_mul(O, O, O2);
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i];
        memcpy(V, R, sizeof(R));
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3);
        for (int i = 0; i < 9; i++) {
            R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
            V[i] = (i % 4 == 0 ? 1.0 : 0.0) + b * O[i] + c * O2[i];
        }
    }
    double qd[4], td[3];
    R_to_quat(R, qd);
    quat_normalize(qd);
    for (int i = 0; i < 3; i++) td[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
    /* result._t = td + qd * t ; result._r = qd * q ; normalize */
    double Rd[9];
    quat_to_R(qd, Rd);
    double nt[3];
    for (int i = 0; i < 3; i++) nt[i] = td[i] + Rd[i * 3] * t[0] + Rd[i * 3 + 1] * t[1] + Rd[i * 3 + 2] * t[2];
    double nq[4];
    quat_mul(qd, q, nq);
    quat_normalize(nq);
    memcpy(q, nq, sizeof(nq));
    memcpy(t, nt, sizeof(nt));
}

struct EdgeLin {
    int D;          /* 2 mono, 3 stereo */
    double e[3];    /* error = obs - projection */
    double A[9];    /* D x 3  d e / d point  (_jacobianOplusXi) */
    double B[18];   /* D x 6  d e / d pose   (_jacobianOplusXj) */
};
/* computeError (+ linearizeOplus if lin) of one edge */
void edge_eval(const double* q, const double* t, const double* cam, const double* X, const double* obs, bool lin, EdgeLin* o) {
    double R[9];
    quat_to_R(q, R);
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], bf = cam[4];
    const bool stereo = !(obs[2] < 0);
    o->D = stereo ? 3 : 2;
    if (!stereo) {
        o->e[0] = obs[0] - ((x / z) * fx + cx);
        o->e[1] = obs[1] - ((y / z) * fy + cy);
        o->e[2] = 0;
    } else { /* invz is rounded to float32 (types_six_dof_expmap.cpp:150-157) */
        const float invz = (float)(1.0f / z);
        const double u = x * invz * fx + cx;
        o->e[0] = obs[0] - u;
        o->e[1] = obs[1] - (y * invz * fy + cy);
        o->e[2] = obs[2] - (u - bf * invz);
    }
    if (!lin) return;
    const double z_2 = z * z;
    double* A = o->A;
    double* B = o->B;
    if (!stereo) { /* :103-139 : Xi = -1/z * tmp * R */
        const double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++)
                A[r * 3 + c] = (-1. / z) * (tmp[r * 3] * R[c] + tmp[r * 3 + 1] * R[3 + c] + tmp[r * 3 + 2] * R[6 + c]);
    } else { /* :188-234 */
        for (int c = 0; c < 3; c++) {
            A[c] = -fx * R[c] / z + fx * x * R[6 + c] / z_2;
            A[3 + c] = -fy * R[3 + c] / z + fy * y * R[6 + c] / z_2;
            A[6 + c] = A[c] - bf * R[6 + c] / z_2;
        }
    }
    B[0] = x * y / z_2 * fx; B[1] = -(1 + (x * x / z_2)) * fx; B[2] = y / z * fx; B[3] = -1. / z * fx; B[4] = 0; B[5] = x / z_2 * fx;
    B[6] = (1 + y * y / z_2) * fy; B[7] = -x * y / z_2 * fy; B[8] = -x / z * fy; B[9] = 0; B[10] = -1. / z * fy; B[11] = y / z_2 * fy;
    if (stereo) {
        B[12] = B[0] - bf * y / z_2; B[13] = B[1] + bf * x / z_2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf / z_2;
    }
}
/* Huber (robust_kernel_impl.cpp:78-91): rho(e), rho'(e) */
inline void huber(double e, double delta, double* rho0, double* rho1) {
    const double dsqr = delta * delta;
    if (e <= dsqr) { *rho0 = e; *rho1 = 1.; }
    else { const double s = sqrt(e); *rho0 = 2 * s * delta - dsqr; *rho1 = delta / s; }
}

inline bool inv3(const double* a, double* o) { /* Eigen 3x3 inverse (cofactors) */
    const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
    const double id = 1.0 / det;
    o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return det != 0;
}

/* Block-skyline (6x6 blocks, lower envelope) symmetric matrix + Cholesky. Row j stores columns first[j]..j. */
struct Skyline {
    int n = 0;
    std::vector<int> first, rowoff, last; /* last[k] = max row whose envelope contains column k */
    std::vector<double> a;
    void init(const std::vector<int>& f) {
        n = (int)f.size();
        first = f;
        rowoff.assign(n + 1, 0);
        for (int j = 0; j < n; j++) rowoff[j + 1] = rowoff[j] + (j - first[j] + 1);
        a.assign((size_t)rowoff[n] * 36, 0.0);
        last.assign(n, 0);
        for (int k = 0; k < n; k++) last[k] = k;
        for (int j = 0; j < n; j++) last[first[j]] = std::max(last[first[j]], j);
        for (int k = 1; k < n; k++) last[k] = std::max(last[k], last[k - 1]);
    }
    double* blk(int j, int i) { return &a[(size_t)(rowoff[j] + i - first[j]) * 36]; }
    bool factor() { /* right-looking, in place: A = L L^T */
        for (int k = 0; k < n; k++) {
            double* D = blk(k, k);
            for (int c = 0; c < 6; c++) { /* dense 6x6 Cholesky of the diagonal block (lower) */
                double s = D[c * 6 + c];
                for (int p = 0; p < c; p++) s -= D[c * 6 + p] * D[c * 6 + p];
                if (!(s > 0.0)) return false;
                const double d = sqrt(s);
                D[c * 6 + c] = d;
                for (int r = c + 1; r < 6; r++) {
                    double v = D[r * 6 + c];
                    for (int p = 0; p < c; p++) v -= D[r * 6 + p] * D[c * 6 + p];
                    D[r * 6 + c] = v / d;
                }
                for (int r = 0; r < c; r++) D[r * 6 + c] = 0.0;
            }
            for (int j = k + 1; j <= last[k]; j++) { /* L_jk = A_jk L_kk^-T */
                if (first[j] > k) continue;
                double* B = blk(j, k);
                for (int r = 0; r < 6; r++)
                    for (int c = 0; c < 6; c++) {
                        double v = B[r * 6 + c];
                        for (int p = 0; p < c; p++) v -= B[r * 6 + p] * D[c * 6 + p];
                        B[r * 6 + c] = v / D[c * 6 + c];
                    }
            }
            for (int j = k + 1; j <= last[k]; j++) { /* trailing update A_ji -= L_jk L_ik^T */
                if (first[j] > k) continue;
                const double* Lj = blk(j, k);
                for (int i = k + 1; i <= j; i++) {
                    if (first[i] > k) continue;
                    const double* Li = blk(i, k);
                    double* T = blk(j, i);
                    for (int r = 0; r < 6; r++)
                        for (int c = 0; c < 6; c++) {
                            double s = 0;
                            for (int p = 0; p < 6; p++) s += Lj[r * 6 + p] * Li[c * 6 + p];
                            T[r * 6 + c] -= s;
                        }
                }
            }
        }
        return true;
    }
    void solve(double* x) { /* in place: L L^T x = b */
        for (int j = 0; j < n; j++) {
            double* xj = x + 6 * j;
            for (int i = first[j]; i < j; i++) {
                const double* L = blk(j, i);
                const double* xi = x + 6 * i;
                for (int r = 0; r < 6; r++)
                    for (int c = 0; c < 6; c++) xj[r] -= L[r * 6 + c] * xi[c];
            }
            const double* D = blk(j, j);
            for (int r = 0; r < 6; r++) {
                double v = xj[r];
                for (int p = 0; p < r; p++) v -= D[r * 6 + p] * xj[p];
                xj[r] = v / D[r * 6 + r];
            }
        }
        for (int j = n - 1; j >= 0; j--) {
            double* xj = x + 6 * j;
            const double* D = blk(j, j);
            for (int r = 5; r >= 0; r--) {
                double v = xj[r];
                for (int p = r + 1; p < 6; p++) v -= D[p * 6 + r] * xj[p];
                xj[r] = v / D[r * 6 + r];
            }
            for (int i = first[j]; i < j; i++) {
                const double* L = blk(j, i);
                double* xi = x + 6 * i;
                for (int r = 0; r < 6; r++)
                    for (int c = 0; c < 6; c++) xi[c] -= L[r * 6 + c] * xj[r];
            }
        }
    }
};

struct Solver {
    const oracle_ba_problem* p;
    int robust;
    oracle_allreduce_fn ar;
    void* ar_user;
    int P, L, E, Pf, Lf;
    std::vector<int> pfree, lfree;        /* index among free poses / points, -1 if fixed */
    std::vector<int> lm_off, lm_edges;    /* edges grouped by landmark, original order kept inside a landmark */
    std::vector<uint8_t> active;          /* edge has at least one free vertex */
    std::vector<double> q, t, X;          /* working estimates */
    std::vector<double> Hpp, bp, Hll, bl, W; /* Hpp [Pf*36], bp [Pf*6], Hll [L*9], bl [L*3], W [E*18] = B^T Omega A (6x3) */
    Skyline S;
    std::vector<double> bs, xp, xl;
    double delta2d, delta3d;

    int reduce(double* buf, int n, int op) { return ar ? ar(ar_user, buf, n, op) : 0; }

    void setup() {
        P = p->n_poses; L = p->n_points; E = p->n_edges;
        pfree.assign(P, -1); lfree.assign(L, -1);
        Pf = Lf = 0;
        for (int i = 0; i < P; i++) if (!p->pose_fixed[i]) pfree[i] = Pf++;
        for (int i = 0; i < L; i++) if (!p->point_fixed[i]) lfree[i] = Lf++;
        lm_off.assign(L + 1, 0);
        for (int e = 0; e < E; e++) lm_off[p->edge_point[e] + 1]++;
        for (int i = 0; i < L; i++) lm_off[i + 1] += lm_off[i];
        lm_edges.resize(E);
        std::vector<int> cur(lm_off.begin(), lm_off.end() - 1);
        for (int e = 0; e < E; e++) lm_edges[cur[p->edge_point[e]]++] = e;
        active.resize(E);
        for (int e = 0; e < E; e++) active[e] = pfree[p->edge_pose[e]] >= 0 || lfree[p->edge_point[e]] >= 0;
        q.assign(p->pose_q, p->pose_q + 4 * P);
        t.assign(p->pose_t, p->pose_t + 3 * P);
        X.assign(p->point_xyz, p->point_xyz + 3 * L);
        delta2d = (float)sqrt(5.99); delta3d = (float)sqrt(7.815); /* Optimizer.cc:101-102 (float) */
        /* envelope of the reduced system: pose j couples with every free pose that shares a free landmark with it */
        std::vector<double> f(std::max(Pf, 1));
        for (int j = 0; j < Pf; j++) f[j] = j;
        for (int l = 0; l < L; l++) {
            if (lfree[l] < 0) continue;
            int mn = Pf;
            for (int k = lm_off[l]; k < lm_off[l + 1]; k++) {
                const int pj = pfree[p->edge_pose[lm_edges[k]]];
                if (pj >= 0) mn = std::min(mn, pj);
            }
            for (int k = lm_off[l]; k < lm_off[l + 1]; k++) {
                const int pj = pfree[p->edge_pose[lm_edges[k]]];
                if (pj >= 0) f[pj] = std::min(f[pj], (double)mn);
            }
        }
        reduce(f.data(), Pf, 1);
        std::vector<int> fi(Pf);
        for (int j = 0; j < Pf; j++) fi[j] = (int)f[j];
        S.init(fi);
        Hpp.resize((size_t)Pf * 36); bp.resize((size_t)Pf * 6); Hll.resize((size_t)L * 9); bl.resize((size_t)L * 3);
        W.resize((size_t)E * 18);
        bs.resize((size_t)Pf * 6); xp.resize((size_t)Pf * 6); xl.resize((size_t)L * 3);
    }

    double chi2() { /* computeActiveErrors + activeRobustChi2 (sparse_optimizer.cpp:100-114) */
        double chi = 0;
        EdgeLin el;
        for (int e = 0; e < E; e++) {
            if (!active[e]) continue;
            const int pi = p->edge_pose[e], li = p->edge_point[e];
            edge_eval(&q[4 * pi], &t[3 * pi], p->pose_cam + 5 * pi, &X[3 * li], p->edge_obs + 3 * e, false, &el);
            double c = 0;
            for (int r = 0; r < el.D; r++) c += el.e[r] * p->edge_inv_sigma2[e] * el.e[r];
            if (robust) { double r0, r1; huber(c, el.D == 2 ? delta2d : delta3d, &r0, &r1); chi += r0; }
            else chi += c;
        }
        double buf = chi;
        reduce(&buf, 1, 0);
        return buf;
    }

    void build() { /* buildSystem (block_solver.hpp:502-560) */
        std::fill(Hpp.begin(), Hpp.end(), 0.0); std::fill(bp.begin(), bp.end(), 0.0);
        std::fill(Hll.begin(), Hll.end(), 0.0); std::fill(bl.begin(), bl.end(), 0.0);
        std::fill(W.begin(), W.end(), 0.0);
        EdgeLin el;
        for (int e = 0; e < E; e++) {
            if (!active[e]) continue;
            const int pi = p->edge_pose[e], li = p->edge_point[e];
            const int pf = pfree[pi], lf = lfree[li];
            edge_eval(&q[4 * pi], &t[3 * pi], p->pose_cam + 5 * pi, &X[3 * li], p->edge_obs + 3 * e, true, &el);
            const int D = el.D;
            double om = p->edge_inv_sigma2[e];
            double omega_r[3];
            for (int r = 0; r < D; r++) omega_r[r] = -om * el.e[r];
            if (robust) {
                double c = 0;
                for (int r = 0; r < D; r++) c += el.e[r] * om * el.e[r];
                double r0, r1;
                huber(c, D == 2 ? delta2d : delta3d, &r0, &r1);
                for (int r = 0; r < D; r++) omega_r[r] *= r1;
                om *= r1; /* robustInformation = rho' * Omega (base_edge.h:96-102) */
            }
            if (lf >= 0) {
                double* H = &Hll[(size_t)li * 9];
                double* b = &bl[(size_t)li * 3];
                for (int a = 0; a < 3; a++) {
                    for (int c = 0; c < 3; c++) {
                        double s = 0;
                        for (int r = 0; r < D; r++) s += el.A[r * 3 + a] * om * el.A[r * 3 + c];
                        H[a * 3 + c] += s;
                    }
                    double s = 0;
                    for (int r = 0; r < D; r++) s += el.A[r * 3 + a] * omega_r[r];
                    b[a] += s;
                }
            }
            if (pf >= 0) {
                double* H = &Hpp[(size_t)pf * 36];
                double* b = &bp[(size_t)pf * 6];
                for (int a = 0; a < 6; a++) {
                    for (int c = 0; c < 6; c++) {
                        double s = 0;
                        for (int r = 0; r < D; r++) s += el.B[r * 6 + a] * om * el.B[r * 6 + c];
                        H[a * 6 + c] += s;
                    }
                    double s = 0;
                    for (int r = 0; r < D; r++) s += el.B[r * 6 + a] * omega_r[r];
                    b[a] += s;
                }
            }
            if (pf >= 0 && lf >= 0) {
                double* w = &W[(size_t)e * 18];
                for (int a = 0; a < 6; a++)
                    for (int c = 0; c < 3; c++) {
                        double s = 0;
                        for (int r = 0; r < D; r++) s += el.B[r * 6 + a] * om * el.A[r * 3 + c];
                        w[a * 3 + c] = s;
                    }
            }
        }
    }

    double lambda_init() { /* computeLambdaInit (optimization_algorithm_levenberg.cpp:166-180) */
        std::vector<double> d((size_t)Pf * 6 + 1, 0.0);
        for (int j = 0; j < Pf; j++)
            for (int a = 0; a < 6; a++) d[(size_t)j * 6 + a] = Hpp[(size_t)j * 36 + a * 7];
        reduce(d.data(), Pf * 6, 0);
        double mx = 0;
        for (int i = 0; i < Pf * 6; i++) mx = std::max(fabs(d[i]), mx);
        for (int l = 0; l < L; l++)
            if (lfree[l] >= 0)
                for (int a = 0; a < 3; a++) mx = std::max(fabs(Hll[(size_t)l * 9 + a * 4]), mx);
        reduce(&mx, 1, 2);
        return 1e-5 * mx;
    }

    bool solve(double lambda, double* scale_out) { /* setLambda + BlockSolver::solve + computeScale */
        std::fill(S.a.begin(), S.a.end(), 0.0);
        for (int j = 0; j < Pf; j++) {
            double* D = S.blk(j, j);
            memcpy(D, &Hpp[(size_t)j * 36], 36 * sizeof(double));
        }
        std::vector<double> coeff((size_t)Pf * 6, 0.0);
        std::vector<double> Dinv((size_t)L * 9, 0.0);
        for (int l = 0; l < L; l++) {
            if (lfree[l] < 0) continue;
            double D[9];
            memcpy(D, &Hll[(size_t)l * 9], sizeof(D));
            D[0] += lambda; D[4] += lambda; D[8] += lambda;
            double* Di = &Dinv[(size_t)l * 9];
            inv3(D, Di);
            double db[3];
            for (int a = 0; a < 3; a++) db[a] = Di[a * 3] * bl[(size_t)l * 3] + Di[a * 3 + 1] * bl[(size_t)l * 3 + 1] + Di[a * 3 + 2] * bl[(size_t)l * 3 + 2];
            for (int k1 = lm_off[l]; k1 < lm_off[l + 1]; k1++) {
                const int e1 = lm_edges[k1];
                const int p1 = pfree[p->edge_pose[e1]];
                if (p1 < 0) continue;
                const double* W1 = &W[(size_t)e1 * 18];
                double BD[18];
                for (int a = 0; a < 6; a++)
                    for (int c = 0; c < 3; c++) BD[a * 3 + c] = W1[a * 3] * Di[c] + W1[a * 3 + 1] * Di[3 + c] + W1[a * 3 + 2] * Di[6 + c];
                for (int a = 0; a < 6; a++) coeff[(size_t)p1 * 6 + a] += W1[a * 3] * db[0] + W1[a * 3 + 1] * db[1] + W1[a * 3 + 2] * db[2];
                for (int k2 = lm_off[l]; k2 < lm_off[l + 1]; k2++) {
                    const int e2 = lm_edges[k2];
                    const int p2 = pfree[p->edge_pose[e2]];
                    if (p2 < 0 || p2 < p1) continue; /* lower block (p2,p1), p2 >= p1 */
                    const double* W2 = &W[(size_t)e2 * 18];
                    double* T = S.blk(p2, p1);
                    for (int a = 0; a < 6; a++)      /* row of p2 */
                        for (int c = 0; c < 6; c++)  /* col of p1 */
                            T[a * 6 + c] -= W2[a * 3] * BD[c * 3] + W2[a * 3 + 1] * BD[c * 3 + 1] + W2[a * 3 + 2] * BD[c * 3 + 2];
                }
            }
        }
        for (int i = 0; i < Pf * 6; i++) bs[i] = bp[i] - coeff[i];
        /* sum the partial reduced systems of all ranks, then damp the pose diagonal once */
        reduce(S.a.data(), (int)S.a.size(), 0);
        reduce(bs.data(), Pf * 6, 0);
        for (int j = 0; j < Pf; j++) {
            double* D = S.blk(j, j);
            for (int a = 0; a < 6; a++) D[a * 7] += lambda;
            for (int a = 0; a < 6; a++) /* keep the diagonal block symmetric: only its lower part is factored */
                for (int c = a + 1; c < 6; c++) D[a * 6 + c] = D[c * 6 + a];
        }
        xp = bs;
        bool ok = S.factor();
        if (ok) S.solve(xp.data());
        else std::fill(xp.begin(), xp.end(), 0.0);
        /* xl = Dinv (bl - Hpl^T xp) */
        std::fill(xl.begin(), xl.end(), 0.0);
        for (int l = 0; l < L && ok; l++) {
            if (lfree[l] < 0) continue;
            double c[3] = {bl[(size_t)l * 3], bl[(size_t)l * 3 + 1], bl[(size_t)l * 3 + 2]};
            for (int k = lm_off[l]; k < lm_off[l + 1]; k++) {
                const int e = lm_edges[k];
                const int pj = pfree[p->edge_pose[e]];
                if (pj < 0) continue;
                const double* w = &W[(size_t)e * 18];
                for (int a = 0; a < 6; a++)
                    for (int cc = 0; cc < 3; cc++) c[cc] -= w[a * 3 + cc] * xp[(size_t)pj * 6 + a];
            }
            const double* Di = &Dinv[(size_t)l * 9];
            for (int a = 0; a < 3; a++) xl[(size_t)l * 3 + a] = Di[a * 3] * c[0] + Di[a * 3 + 1] * c[1] + Di[a * 3 + 2] * c[2];
        }
        /* computeScale: sum_j x_j (lambda x_j + b_j)  (optimization_algorithm_levenberg.cpp:182-189) */
        double part = 0, xx = 0;
        for (int i = 0; i < Pf * 6; i++) { part += xp[i] * bp[i]; xx += xp[i] * xp[i]; }
        for (int l = 0; l < L; l++)
            if (lfree[l] >= 0)
                for (int a = 0; a < 3; a++) part += xl[(size_t)l * 3 + a] * (lambda * xl[(size_t)l * 3 + a] + bl[(size_t)l * 3 + a]);
        reduce(&part, 1, 0);
        *scale_out = part + lambda * xx;
        return ok;
    }

    void update() { /* SparseOptimizer::update */
        for (int i = 0; i < P; i++)
            if (pfree[i] >= 0) pose_oplus(&q[4 * i], &t[3 * i], &xp[(size_t)pfree[i] * 6]);
        for (int l = 0; l < L; l++)
            if (lfree[l] >= 0)
                for (int a = 0; a < 3; a++) X[(size_t)l * 3 + a] += xl[(size_t)l * 3 + a];
    }
};

} // namespace

extern "C" {

int oracle_ba_solve(oracle_ba_problem* prob, int iterations, const volatile uint8_t* stop, int robust, oracle_ba_result* res,
                    oracle_allreduce_fn ar, void* ar_user) {
    Solver s;
    s.p = prob; s.robust = robust; s.ar = ar; s.ar_user = ar_user;
    s.setup();
    memset(res, 0, sizeof(*res));
    double lambda = -1, ni = 2;
    int nBad = 0;
    bool ok = true;
    auto terminate = [&]() { return stop && *stop; };
    int it = 0;
    for (; it < iterations && !terminate() && ok; it++) {
        double currentChi = s.chi2();
        const double iniChi = currentChi;
        if (it == 0) res->chi2_initial = currentChi;
        s.build();
        if (it == 0) { lambda = s.lambda_init(); ni = 2; nBad = 0; res->lambda_initial = lambda; }
        double rho = 0;
        int qmax = 0;
        do {
            std::vector<double> q0 = s.q, t0 = s.t, X0 = s.X; /* push */
            double scale;
            const bool ok2 = s.solve(lambda, &scale);
            if (!ok2) res->solver_failures++;
            s.update();
            double tempChi = s.chi2();
            if (!ok2) tempChi = DBL_MAX;
            rho = (currentChi - tempChi) / (scale + 1e-3);
            const bool good = rho > 0 && std::isfinite(tempChi);
            if (res->n_trials < 256) { res->trial_accepted[res->n_trials] = good; res->trial_chi2[res->n_trials] = tempChi; }
            res->n_trials++;
            if (good) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                const double sf = std::max(1. / 3., alpha);
                lambda *= sf;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                s.q = q0; s.t = t0; s.X = X0; /* pop */
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
    memcpy(prob->pose_q, s.q.data(), sizeof(double) * 4 * s.P);
    memcpy(prob->pose_t, s.t.data(), sizeof(double) * 3 * s.P);
    memcpy(prob->point_xyz, s.X.data(), sizeof(double) * 3 * s.L);
    return 0;
}

/* exposed for algebraic pinning: error and Jacobians of one edge, and the manifold update */
void oracle_ba_edge(const double* q, const double* t, const double* cam, const double* X, const double* obs, int* D, double* e,
                    double* A, double* B) {
    EdgeLin el;
    memset(&el, 0, sizeof(el));
    edge_eval(q, t, cam, X, obs, true, &el);
    *D = el.D;
    memcpy(e, el.e, sizeof(el.e)); memcpy(A, el.A, sizeof(el.A)); memcpy(B, el.B, sizeof(el.B));
}
void oracle_ba_pose_oplus(double* q, double* t, const double* delta) { pose_oplus(q, t, delta); }
/* total chi2 (no robust kernel) and RMS reprojection error over active edges at the problem's current estimate */
double oracle_ba_chi2(const oracle_ba_problem* prob, double* rms_px) {
    double chi = 0, ss = 0;
    long n = 0;
    EdgeLin el;
    for (int e = 0; e < prob->n_edges; e++) {
        const int pi = prob->edge_pose[e], li = prob->edge_point[e];
        if (prob->pose_fixed[pi] && prob->point_fixed[li]) continue;
        edge_eval(prob->pose_q + 4 * pi, prob->pose_t + 3 * pi, prob->pose_cam + 5 * pi, prob->point_xyz + 3 * li,
                  prob->edge_obs + 3 * e, false, &el);
        for (int r = 0; r < el.D; r++) { chi += el.e[r] * prob->edge_inv_sigma2[e] * el.e[r]; ss += el.e[r] * el.e[r]; n++; }
    }
    if (rms_px) *rms_px = n ? sqrt(ss / n) : 0;
    return chi;
}

} // extern "C"
