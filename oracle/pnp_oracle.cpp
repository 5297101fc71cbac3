/* CPU ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Plain C++ restatement of the reference's EPnP-RANSAC, sequential exactly like the reference:
 *   corbslam_client/src/PnPsolver.cc:66-151   PnPsolver::PnPsolver (flattening of Frame / KeyFrame + MapPoint matches)
 *   corbslam_client/src/PnPsolver.cc:163-198  SetRansacParameters
 *   corbslam_client/src/PnPsolver.cc:206-300  iterate (minimal sets of 4, best-so-far bookkeeping, early exit on Refine)
 *   corbslam_client/src/PnPsolver.cc:302-346  Refine (EPnP over mvbBestInliers - the best set, not the current one)
 *   corbslam_client/src/PnPsolver.cc:349-383  CheckInliers (float / double mix kept expression by expression)
 *   corbslam_client/src/PnPsolver.cc:420-962  EPnP (Lepetit / Moreno-Noguer / Fua): control points, barycentric
 *                                             coordinates, M, L_6x10, rho, three beta approximations, Gauss-Newton with
 *                                             the Householder qr_solve, Procrustes R|t, reprojection error
 *   Thirdparty/DBoW2/DUtils/Random.cpp:47-50  RandomInt (the draws themselves are an INPUT here, see below)
 *
 * Third-party arithmetic that is not under /root/reference: OpenCV (>= 2.4.3, unpinned; cl/CMakeLists.txt:43) -
 * cvSVD, cvSolve(CV_SVD), cvInvert(CV_SVD), cvMulTransposed, called at PnPsolver.cc:446-447,468,518-519,626,718,749,782.
 * Restated from OpenCV's published algorithm (modules/core/src/lapack.cpp: JacobiSVDImpl_ = one-sided Hestenes Jacobi on
 * the rows of A^T with the relative threshold 10*DBL_EPSILON, at most max(m,30) sweeps, singular values sorted
 * descending, U = rows scaled by 1/w; SVBkSbImpl_ = sum over singular values above 2*DBL_EPSILON*sum(w) of
 * v_i (u_i.b / w_i)). Two deliberate differences, both documented in DESIGN.md: hypot(p, beta) is evaluated as
 * sqrt(p*p + beta*beta) (only + - * / sqrt remain, which are correctly rounded on the host and on the GPU, so the GPU
 * path can be compared bit for bit), and a singular value <= DBL_MIN leaves a zero left vector instead of OpenCV's
 * random completion.
 *
 * PARITY STATUS: PINNED against the reference's own PnPsolver.cc compiled unmodified into oracle/_ref/libref.so (cvSVD /
 * cvSolve / cvInvert of the stub OpenCV are the restatements below): with the draws its rand() stream produces, status,
 * bNoMore, inlier sets, iteration counts and the float32 Tcw are equal bit for bit over resumed calls
 * (tests/test_ref_cpu.py). The RANSAC of the reference consumes the process-global rand() stream, so its outcome is not a
 * function of its arguments; here the draws are. Also pinned: (1) the SVD / solve / invert restatements against cv2.SVDecomp / cv2.solve / cv2.invert
 * 4.13.0 (LAPACK-backed in this wheel, so to 1e-9, not bit for bit) and (2) compute_pose against cv2.solvePnP
 * (SOLVEPNP_EPNP) - OpenCV's own copy of the same EPnP code - on 6..200 points (tests/golden/pnp_cv2.npz,
 * tools/gen_golden_pnp.py). With a minimal set of 4 points M^T M has a 4-dimensional null space whose basis is decided
 * by rounding noise in ANY SVD, so per-hypothesis poses are only reproducible between implementations that perform the
 * same operations in the same order; that is what the GPU path does and what the parity tests check.
 *
 * The random draws are an input: draws[4*it + k] is the value RandomInt(0, size-1) returned for pick k of RANSAC
 * iteration `it` (counted from the solver's first iteration), so the outcome is a pure function of the arguments.
 * Built with -ffp-contract=off.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

/* ---- OpenCV JacobiSVDImpl_<double>: At = A^T (n rows of length m, m >= n). On return row i of At is the left
 * singular vector u_i, W the singular values (descending), row i of Vt the right singular vector v_i. */
void jacobi_svd(double* At, double* W, double* Vt, int m, int n) {
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    const int max_iter = std::max(m, 30);
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { const double t = At[i * m + k]; sd += t * t; }
        W[i] = sd;
        for (int k = 0; k < n; k++) Vt[i * n + k] = 0;
        Vt[i * n + i] = 1;
    }
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double *Ai = At + i * m, *Aj = At + j * m;
                double a = W[i], p = 0, b = W[j];
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
                for (int k = 0; k < m; k++) {
                    const double t0 = c * Ai[k] + s * Aj[k];
                    const double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                double *Vi = Vt + i * n, *Vj = Vt + j * n;
                for (int k = 0; k < n; k++) {
                    const double t0 = c * Vi[k] + s * Vj[k];
                    const double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
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
            std::swap(W[i], W[j]);
            for (int k = 0; k < m; k++) std::swap(At[i * m + k], At[j * m + k]);
            for (int k = 0; k < n; k++) std::swap(Vt[i * n + k], Vt[j * n + k]);
        }
    }
    for (int i = 0; i < n; i++) {
        const double sd = W[i];
        const double s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < m; k++) At[i * m + k] *= s;
    }
}

/* SVD of a row-major m x n matrix (m >= n): Ut (n x m, rows = left vectors), W, Vt (n x n). cv::SVD::compute transposes
 * the source into the work buffer before JacobiSVD. */
void svd(const double* A, int m, int n, double* Ut, double* W, double* Vt) {
    for (int i = 0; i < n; i++)
        for (int k = 0; k < m; k++) Ut[i * m + k] = A[k * n + i];
    jacobi_svd(Ut, W, Vt, m, n);
}

/* OpenCV SVBkSbImpl_ with one right-hand side: x = sum_i v_i * ((u_i . b) * (1 / w_i)) over w_i > 2 eps sum(w). */
void svd_solve(const double* A, int m, int n, const double* b, double* x) {
    double Ut[6 * 6], W[6], Vt[6 * 6];
    svd(A, m, n, Ut, W, Vt);
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

/* cvInvert(CV_SVD) of a 3 x 3 matrix: inv[j][k] = sum_i v_i[j] * (u_i[k] * (1 / w_i)). */
void svd_invert3(const double* A, double* inv) {
    double Ut[9], W[3], Vt[9];
    svd(A, 3, 3, Ut, W, Vt);
    double threshold = 0;
    for (int i = 0; i < 3; i++) threshold += W[i];
    threshold *= DBL_EPSILON * 2;
    for (int i = 0; i < 9; i++) inv[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = W[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double buffer[3];
        for (int k = 0; k < 3; k++) buffer[k] = Ut[i * 3 + k] * wi;
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) inv[j * 3 + k] = inv[j * 3 + k] + Vt[i * 3 + j] * buffer[k];
    }
}

/* cvMulTransposed(src, dst, 1): dst = src^T src, every element summed over the rows in order. */
void mul_transposed(const double* src, int rows, int cols, double* dst) {
    for (int i = 0; i < cols; i++)
        for (int j = i; j < cols; j++) {
            double s = 0;
            for (int k = 0; k < rows; k++) s += src[k * cols + i] * src[k * cols + j];
            dst[i * cols + j] = s;
            dst[j * cols + i] = s;
        }
}

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; } /* :563-566 */
inline double dist2(const double* p1, const double* p2) {                                                /* :555-561 */
    return (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) + (p1[2] - p2[2]) * (p1[2] - p2[2]);
}

struct Epnp {
    double uc, vc, fu, fv;
    std::vector<double> pws, us, alphas, pcs;
    int n = 0;
    double cws[4][3], ccs[4][3];

    void reset() { n = 0; pws.clear(); us.clear(); }
    void add(double X, double Y, double Z, double u, double v) { /* :407-418 */
        pws.push_back(X); pws.push_back(Y); pws.push_back(Z);
        us.push_back(u); us.push_back(v);
        n++;
    }

    void choose_control_points() { /* :420-455 */
        cws[0][0] = cws[0][1] = cws[0][2] = 0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < 3; j++) cws[0][j] += pws[3 * i + j];
        for (int j = 0; j < 3; j++) cws[0][j] /= n;
        std::vector<double> pw0(3 * n);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < 3; j++) pw0[3 * i + j] = pws[3 * i + j] - cws[0][j];
        double pw0tpw0[9], dc[3], uct[9], vt[9];
        mul_transposed(pw0.data(), n, 3, pw0tpw0);
        svd(pw0tpw0, 3, 3, uct, dc, vt);
        for (int i = 1; i < 4; i++) {
            const double k = sqrt(dc[i - 1] / n);
            for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + k * uct[3 * (i - 1) + j];
        }
    }

    void compute_barycentric_coordinates() { /* :457-482 */
        double cc[9], ci[9];
        for (int i = 0; i < 3; i++)
            for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
        svd_invert3(cc, ci);
        alphas.resize(4 * n);
        for (int i = 0; i < n; i++) {
            const double* pi = &pws[3 * i];
            double* a = &alphas[4 * i];
            for (int j = 0; j < 3; j++)
                a[1 + j] = ci[3 * j] * (pi[0] - cws[0][0]) + ci[3 * j + 1] * (pi[1] - cws[0][1]) + ci[3 * j + 2] * (pi[2] - cws[0][2]);
            a[0] = 1.0f - a[1] - a[2] - a[3];
        }
    }

    void fill_M(double* M, int row, const double* as, double u, double v) { /* :484-500 */
        double* M1 = M + row * 12;
        double* M2 = M1 + 12;
        for (int i = 0; i < 4; i++) {
            M1[3 * i] = as[i] * fu; M1[3 * i + 1] = 0.0; M1[3 * i + 2] = as[i] * (uc - u);
            M2[3 * i] = 0.0; M2[3 * i + 1] = as[i] * fv; M2[3 * i + 2] = as[i] * (vc - v);
        }
    }

    void compute_ccs(const double* betas, const double* ut) { /* :502-514 */
        for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0f;
        for (int i = 0; i < 4; i++) {
            const double* v = ut + 12 * (11 - i);
            for (int j = 0; j < 4; j++)
                for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
        }
    }

    void compute_pcs() { /* :516-525 */
        pcs.resize(3 * n);
        for (int i = 0; i < n; i++) {
            const double* a = &alphas[4 * i];
            double* pc = &pcs[3 * i];
            for (int j = 0; j < 3; j++) pc[j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
        }
    }

    void solve_for_sign() { /* :660-674 */
        if (pcs[2] < 0.0) {
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 3; j++) ccs[i][j] = -ccs[i][j];
            for (int i = 0; i < n; i++) { pcs[3 * i] = -pcs[3 * i]; pcs[3 * i + 1] = -pcs[3 * i + 1]; pcs[3 * i + 2] = -pcs[3 * i + 2]; }
        }
    }

    void estimate_R_and_t(double R[3][3], double t[3]) { /* :587-651 */
        double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
        for (int i = 0; i < n; i++)
            for (int j = 0; j < 3; j++) { pc0[j] += pcs[3 * i + j]; pw0[j] += pws[3 * i + j]; }
        for (int j = 0; j < 3; j++) { pc0[j] /= n; pw0[j] /= n; }
        double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, abt_d[3], abt_ut[9], abt_vt[9];
        for (int i = 0; i < n; i++) {
            const double* pc = &pcs[3 * i];
            const double* pw = &pws[3 * i];
            for (int j = 0; j < 3; j++) {
                abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
                abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
                abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
            }
        }
        svd(abt, 3, 3, abt_ut, abt_d, abt_vt);
        /* cvSVD without the _T flags returns U and V: U[i][k] = abt_ut[k][i], V[j][k] = abt_vt[k][j];
         * R[i][j] = dot(row i of U, row j of V) */
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                R[i][j] = abt_ut[0 * 3 + i] * abt_vt[0 * 3 + j] + abt_ut[1 * 3 + i] * abt_vt[1 * 3 + j] + abt_ut[2 * 3 + i] * abt_vt[2 * 3 + j];
        const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                           R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
        if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
        t[0] = pc0[0] - dot3(R[0], pw0);
        t[1] = pc0[1] - dot3(R[1], pw0);
        t[2] = pc0[2] - dot3(R[2], pw0);
    }

    double reprojection_error(const double R[3][3], const double t[3]) { /* :568-585 */
        double sum2 = 0.0;
        for (int i = 0; i < n; i++) {
            const double* pw = &pws[3 * i];
            const double Xc = dot3(R[0], pw) + t[0];
            const double Yc = dot3(R[1], pw) + t[1];
            const double inv_Zc = 1.0 / (dot3(R[2], pw) + t[2]);
            const double ue = uc + fu * Xc * inv_Zc;
            const double ve = vc + fv * Yc * inv_Zc;
            const double u = us[2 * i], v = us[2 * i + 1];
            sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
        }
        return sum2 / n;
    }

    double compute_R_and_t(const double* ut, const double* betas, double R[3][3], double t[3]) { /* :676-687 */
        compute_ccs(betas, ut);
        compute_pcs();
        solve_for_sign();
        estimate_R_and_t(R, t);
        return reprojection_error(R, t);
    }

    static void find_betas_approx_1(const double* L, const double* rho, double* betas) { /* :692-719 */
        double l_6x4[24], b4[4];
        for (int i = 0; i < 6; i++) {
            l_6x4[4 * i] = L[10 * i]; l_6x4[4 * i + 1] = L[10 * i + 1]; l_6x4[4 * i + 2] = L[10 * i + 3]; l_6x4[4 * i + 3] = L[10 * i + 6];
        }
        svd_solve(l_6x4, 6, 4, rho, b4);
        if (b4[0] < 0) {
            betas[0] = sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0];
        } else {
            betas[0] = sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0];
        }
    }
    static void find_betas_approx_2(const double* L, const double* rho, double* betas) { /* :724-752 */
        double l_6x3[18], b3[3];
        for (int i = 0; i < 6; i++) { l_6x3[3 * i] = L[10 * i]; l_6x3[3 * i + 1] = L[10 * i + 1]; l_6x3[3 * i + 2] = L[10 * i + 2]; }
        svd_solve(l_6x3, 6, 3, rho, b3);
        if (b3[0] < 0) {
            betas[0] = sqrt(-b3[0]);
            betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0;
        } else {
            betas[0] = sqrt(b3[0]);
            betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0;
        }
        if (b3[1] < 0) betas[0] = -betas[0];
        betas[2] = 0.0;
        betas[3] = 0.0;
    }
    static void find_betas_approx_3(const double* L, const double* rho, double* betas) { /* :757-785 */
        double l_6x5[30], b5[5];
        for (int i = 0; i < 6; i++)
            for (int k = 0; k < 5; k++) l_6x5[5 * i + k] = L[10 * i + k];
        svd_solve(l_6x5, 6, 5, rho, b5);
        if (b5[0] < 0) {
            betas[0] = sqrt(-b5[0]);
            betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0;
        } else {
            betas[0] = sqrt(b5[0]);
            betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0;
        }
        if (b5[1] < 0) betas[0] = -betas[0];
        betas[2] = b5[3] / betas[0];
        betas[3] = 0.0;
    }

    static void compute_L_6x10(const double* ut, double* l_6x10) { /* :787-829 */
        const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
        double dv[4][6][3];
        for (int i = 0; i < 4; i++) {
            int a = 0, b = 1;
            for (int j = 0; j < 6; j++) {
                dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
                dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
                dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
                b++;
                if (b > 3) { a++; b = a + 1; }
            }
        }
        for (int i = 0; i < 6; i++) {
            double* row = l_6x10 + 10 * i;
            row[0] = dot3(dv[0][i], dv[0][i]);
            row[1] = 2.0f * dot3(dv[0][i], dv[1][i]);
            row[2] = dot3(dv[1][i], dv[1][i]);
            row[3] = 2.0f * dot3(dv[0][i], dv[2][i]);
            row[4] = 2.0f * dot3(dv[1][i], dv[2][i]);
            row[5] = dot3(dv[2][i], dv[2][i]);
            row[6] = 2.0f * dot3(dv[0][i], dv[3][i]);
            row[7] = 2.0f * dot3(dv[1][i], dv[3][i]);
            row[8] = 2.0f * dot3(dv[2][i], dv[3][i]);
            row[9] = dot3(dv[3][i], dv[3][i]);
        }
    }

    void compute_rho(double* rho) { /* :831-839 */
        rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);
        rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
    }

    static void compute_A_and_b_gauss_newton(const double* l_6x10, const double* rho, const double betas[4], double* A, double* b) { /* :841-868 */
        for (int i = 0; i < 6; i++) {
            const double* rowL = l_6x10 + i * 10;
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
    }

    /* Householder QR least squares, 6 x 4 (:888-977). A singular column returns without touching X (the reference
     * prints a message and returns; X keeps the previous iteration's step - here the caller's array, like there). */
    static void qr_solve(double* pA, double* pb, double* pX) {
        const int nr = 6, nc = 4;
        double A1[6], A2[6];
        double* ppAkk = pA;
        for (int k = 0; k < nc; k++) {
            double* ppAik = ppAkk;
            double eta = fabs(*ppAik);
            for (int i = k + 1; i < nr; i++) {
                const double elt = fabs(*ppAik);
                if (eta < elt) eta = elt;
                ppAik += nc;
            }
            if (eta == 0) {
                A1[k] = A2[k] = 0.0;
                return;
            } else {
                double *p = ppAkk, sum = 0.0;
                const double inv_eta = 1. / eta;
                for (int i = k; i < nr; i++) {
                    *p *= inv_eta;
                    sum += *p * *p;
                    p += nc;
                }
                double sigma = sqrt(sum);
                if (*ppAkk < 0) sigma = -sigma;
                *ppAkk += sigma;
                A1[k] = sigma * *ppAkk;
                A2[k] = -eta * sigma;
                for (int j = k + 1; j < nc; j++) {
                    double *q = ppAkk, s = 0;
                    for (int i = k; i < nr; i++) { s += *q * q[j - k]; q += nc; }
                    const double tau = s / A1[k];
                    q = ppAkk;
                    for (int i = k; i < nr; i++) { q[j - k] -= tau * *q; q += nc; }
                }
            }
            ppAkk += nc + 1;
        }
        double* ppAjj = pA;
        for (int j = 0; j < nc; j++) {
            double *ppAij = ppAjj, tau = 0;
            for (int i = j; i < nr; i++) { tau += *ppAij * pb[i]; ppAij += nc; }
            tau /= A1[j];
            ppAij = ppAjj;
            for (int i = j; i < nr; i++) { pb[i] -= tau * *ppAij; ppAij += nc; }
            ppAjj += nc + 1;
        }
        pX[nc - 1] = pb[nc - 1] / A2[nc - 1];
        for (int i = nc - 2; i >= 0; i--) {
            double *ppAij = pA + i * nc + (i + 1), sum = 0;
            for (int j = i + 1; j < nc; j++) { sum += *ppAij * pX[j]; ppAij++; }
            pX[i] = (pb[i] - sum) / A2[i];
        }
    }

    static void gauss_newton(const double* L, const double* rho, double betas[4]) { /* :870-886 */
        double a[24], b[6], x[4] = {0, 0, 0, 0};
        for (int k = 0; k < 5; k++) {
            compute_A_and_b_gauss_newton(L, rho, betas, a, b);
            qr_solve(a, b, x);
            for (int i = 0; i < 4; i++) betas[i] += x[i];
        }
    }

    double compute_pose(double R[3][3], double t[3]) { /* :527-574 */
        choose_control_points();
        compute_barycentric_coordinates();
        std::vector<double> M(2 * n * 12);
        for (int i = 0; i < n; i++) fill_M(M.data(), 2 * i, &alphas[4 * i], us[2 * i], us[2 * i + 1]);
        double mtm[144], d[12], ut[144], vt[144];
        mul_transposed(M.data(), 2 * n, 12, mtm);
        svd(mtm, 12, 12, ut, d, vt);
        double l_6x10[60], rho[6];
        compute_L_6x10(ut, l_6x10);
        compute_rho(rho);
        double Betas[4][4], rep_errors[4], Rs[4][3][3], ts[4][3];
        find_betas_approx_1(l_6x10, rho, Betas[1]);
        gauss_newton(l_6x10, rho, Betas[1]);
        rep_errors[1] = compute_R_and_t(ut, Betas[1], Rs[1], ts[1]);
        find_betas_approx_2(l_6x10, rho, Betas[2]);
        gauss_newton(l_6x10, rho, Betas[2]);
        rep_errors[2] = compute_R_and_t(ut, Betas[2], Rs[2], ts[2]);
        find_betas_approx_3(l_6x10, rho, Betas[3]);
        gauss_newton(l_6x10, rho, Betas[3]);
        rep_errors[3] = compute_R_and_t(ut, Betas[3], Rs[3], ts[3]);
        int N = 1;
        if (rep_errors[2] < rep_errors[1]) N = 2;
        if (rep_errors[3] < rep_errors[N]) N = 3;
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) R[i][j] = Rs[N][i][j];
            t[i] = ts[N][i];
        }
        return rep_errors[N];
    }
};

/* The RANSAC wrapper (PnPsolver.cc:66-383), state kept between iterate() calls like the reference object. */
struct PnpSolver {
    Epnp e;
    int N = 0;
    std::vector<float> p2d, p3d, max_err;
    int min_inliers = 0, max_its = 0;
    int mnIterations = 0, mnBestInliers = 0, mnInliersi = 0, mnRefinedInliers = 0;
    std::vector<uint8_t> inliersi, best_inliers, refined_inliers;
    double mRi[3][3], mti[3];
    float best_Tcw[16], refined_Tcw[16];
    const int32_t* draws = nullptr;
    int n_draw_iters = 0;

    void check_inliers() { /* :349-383 */
        mnInliersi = 0;
        for (int i = 0; i < N; i++) {
            const float Px = p3d[3 * i], Py = p3d[3 * i + 1], Pz = p3d[3 * i + 2];
            const float Xc = mRi[0][0] * Px + mRi[0][1] * Py + mRi[0][2] * Pz + mti[0];
            const float Yc = mRi[1][0] * Px + mRi[1][1] * Py + mRi[1][2] * Pz + mti[1];
            const float invZc = 1 / (mRi[2][0] * Px + mRi[2][1] * Py + mRi[2][2] * Pz + mti[2]);
            const double ue = e.uc + e.fu * Xc * invZc;
            const double ve = e.vc + e.fv * Yc * invZc;
            const float distX = p2d[2 * i] - ue;
            const float distY = p2d[2 * i + 1] - ve;
            const float error2 = distX * distX + distY * distY;
            if (error2 < max_err[i]) { inliersi[i] = 1; mnInliersi++; } else inliersi[i] = 0;
        }
    }

    static void to_Tcw(const double R[3][3], const double t[3], float* T) { /* :254-260 */
        for (int i = 0; i < 16; i++) T[i] = 0.f;
        T[15] = 1.f;
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) T[4 * i + j] = (float)R[i][j];
            T[4 * i + 3] = (float)t[i];
        }
    }

    bool refine() { /* :302-346 */
        e.reset();
        for (int i = 0; i < N; i++)
            if (best_inliers[i]) e.add(p3d[3 * i], p3d[3 * i + 1], p3d[3 * i + 2], p2d[2 * i], p2d[2 * i + 1]);
        e.compute_pose(mRi, mti);
        check_inliers();
        mnRefinedInliers = mnInliersi;
        refined_inliers = inliersi;
        if (mnInliersi > min_inliers) {
            to_Tcw(mRi, mti, refined_Tcw);
            return true;
        }
        return false;
    }

    /* returns 1 = refined pose, 2 = best pose at the end (bNoMore), 0 = empty; *no_more as the reference sets it */
    int iterate(int nIterations, int* no_more, uint8_t* inliers, int* n_inliers, float* Tcw) { /* :206-300 */
        *no_more = 0;
        *n_inliers = 0;
        memset(inliers, 0, N);
        if (N < min_inliers) { *no_more = 1; return 0; }
        int nCurrentIterations = 0;
        while (mnIterations < max_its || nCurrentIterations < nIterations) {
            if (mnIterations >= n_draw_iters) return -1; /* the caller did not provide enough draws */
            nCurrentIterations++;
            mnIterations++;
            e.reset();
            /* vAvailableIndices = mvAllIndices (identity); pick, replace by the back, pop (:231-242) */
            std::vector<int> avail(N);
            for (int i = 0; i < N; i++) avail[i] = i;
            for (int k = 0; k < 4; k++) {
                const int randi = draws[4 * (mnIterations - 1) + k];
                const int idx = avail[randi];
                e.add(p3d[3 * idx], p3d[3 * idx + 1], p3d[3 * idx + 2], p2d[2 * idx], p2d[2 * idx + 1]);
                avail[randi] = avail.back();
                avail.pop_back();
            }
            e.compute_pose(mRi, mti);
            check_inliers();
            if (mnInliersi >= min_inliers) {
                if (mnInliersi > mnBestInliers) {
                    best_inliers = inliersi;
                    mnBestInliers = mnInliersi;
                    to_Tcw(mRi, mti, best_Tcw);
                }
                if (refine()) {
                    *n_inliers = mnRefinedInliers;
                    memcpy(inliers, refined_inliers.data(), N);
                    memcpy(Tcw, refined_Tcw, sizeof(refined_Tcw));
                    return 1;
                }
            }
        }
        if (mnIterations >= max_its) {
            *no_more = 1;
            if (mnBestInliers >= min_inliers) {
                *n_inliers = mnBestInliers;
                memcpy(inliers, best_inliers.data(), N);
                memcpy(Tcw, best_Tcw, sizeof(best_Tcw));
                return 2;
            }
        }
        return 0;
    }
};

} // namespace

extern "C" {

/* PnPsolver::SetRansacParameters (PnPsolver.cc:163-198): the adjusted minimum inlier count and iteration cap. */
void oracle_pnp_ransac_params(int N, double probability, int minInliers, int maxIterations, int minSet, float epsilon,
                              int* out_min_inliers, int* out_max_its) {
    float mRansacEpsilon = epsilon;
    int nMinInliers = N * mRansacEpsilon;
    if (nMinInliers < minInliers) nMinInliers = minInliers;
    if (nMinInliers < minSet) nMinInliers = minSet;
    const int mRansacMinInliers = nMinInliers;
    if (mRansacEpsilon < (float)mRansacMinInliers / N) mRansacEpsilon = (float)mRansacMinInliers / N;
    int nIterations;
    if (mRansacMinInliers == N)
        nIterations = 1;
    else
        nIterations = ceil(log(1 - probability) / log(1 - pow(mRansacEpsilon, 3)));
    *out_min_inliers = mRansacMinInliers;
    *out_max_its = std::max(1, std::min(nIterations, maxIterations));
}

void* oracle_pnp_create(int N, const float* p2d, const float* p3d, const float* max_err, float fx, float fy, float cx, float cy,
                        int min_inliers, int max_its) {
    PnpSolver* s = new PnpSolver;
    s->N = N;
    s->p2d.assign(p2d, p2d + 2 * N);
    s->p3d.assign(p3d, p3d + 3 * N);
    s->max_err.assign(max_err, max_err + N);
    s->e.fu = fx; s->e.fv = fy; s->e.uc = cx; s->e.vc = cy; /* :107-110 (float members of Frame widened to double) */
    s->min_inliers = min_inliers;
    s->max_its = max_its;
    s->inliersi.assign(N, 0);
    s->best_inliers.assign(N, 0);
    return s;
}
void oracle_pnp_destroy(void* h) { delete (PnpSolver*)h; }
int oracle_pnp_iterations(void* h) { return ((PnpSolver*)h)->mnIterations; }

int oracle_pnp_iterate(void* h, int nIterations, const int32_t* draws, int n_draw_iters, int* no_more, uint8_t* inliers, int* n_inliers,
                       float* Tcw) {
    PnpSolver* s = (PnpSolver*)h;
    s->draws = draws;
    s->n_draw_iters = n_draw_iters;
    return s->iterate(nIterations, no_more, inliers, n_inliers, Tcw);
}

/* EPnP alone (compute_pose) on n correspondences, for pinning against cv2.solvePnP(SOLVEPNP_EPNP). */
double oracle_epnp_pose(int n, const double* pws, const double* us, double fu, double fv, double uc, double vc, double* R9, double* t3) {
    Epnp e;
    e.fu = fu; e.fv = fv; e.uc = uc; e.vc = vc;
    e.reset();
    for (int i = 0; i < n; i++) e.add(pws[3 * i], pws[3 * i + 1], pws[3 * i + 2], us[2 * i], us[2 * i + 1]);
    double R[3][3], t[3];
    const double err = e.compute_pose(R, t);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) R9[3 * i + j] = R[i][j];
        t3[i] = t[i];
    }
    return err;
}

/* The OpenCV primitive restatements, for pinning against cv2.SVDecomp / cv2.solve / cv2.invert. m >= n, n <= 12. */
void oracle_svd(const double* A, int m, int n, double* Ut, double* W, double* Vt) { svd(A, m, n, Ut, W, Vt); }
void oracle_svd_solve(const double* A, int m, int n, const double* b, double* x) { svd_solve(A, m, n, b, x); }
void oracle_svd_invert3(const double* A, double* inv) { svd_invert3(A, inv); }
void oracle_mul_transposed(const double* src, int rows, int cols, double* dst) { mul_transposed(src, rows, cols, dst); }

} // extern "C"
