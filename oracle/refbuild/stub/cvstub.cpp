/* STUB OpenCV — TEST INFRASTRUCTURE ONLY (see opencv2/core/core.hpp).
 *
 * Containers / expression rewriting restated from OpenCV's published behaviour (modules/core/src/matop.cpp,
 * matmul.cpp, stat.cpp, copy.cpp); observable arithmetic delegated to the cv2-pinned models of oracle/*.cpp:
 *   cv::FAST -> oracle_fast_detect, cv::resize -> oracle_resize_linear_u8, cv::GaussianBlur -> oracle_gaussian7_u8,
 *   cv::fastAtan2 -> oracle_fast_atan2, cvSVD / cvSolve / cvInvert / cvMulTransposed -> oracle_svd*,
 *   cv::gemm (CV_32F) -> the two kernels pinned by tests/golden/opencv_primitives.npz (small-matrix float dot product
 *   when flags == 0 and the inner dimension is 2..4, double accumulation otherwise; oracle_gemm3 is the 3x3 case). */
#include "opencv2/opencv.hpp"

extern "C" {
void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void oracle_gaussian7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
int oracle_fast_detect(const uint8_t* src, int w, int h, int sstride, int th, int32_t* out, int cap);
float oracle_fast_atan2(float y, float x);
void oracle_svd(const double* A, int m, int n, double* Ut, double* W, double* Vt);
void oracle_svd_solve(const double* A, int m, int n, const double* b, double* x);
void oracle_svd_invert3(const double* A, double* inv);
void oracle_mul_transposed(const double* src, int rows, int cols, double* dst);
}

#define STUB_FAIL(msg) do { std::fprintf(stderr, "opencv stub: %s (%s:%d)\n", msg, __FILE__, __LINE__); std::abort(); } while (0)

namespace cv {

/* ------------------------------------------------------------------ element access by depth */
static inline double get_d(const Mat& m, int y, int x) {
    switch (m.depth()) {
    case CV_8U: return m.at<uchar>(y, x);
    case CV_32S: return m.at<int>(y, x);
    case CV_32F: return m.at<float>(y, x);
    case CV_64F: return m.at<double>(y, x);
    }
    STUB_FAIL("unsupported depth");
}
static inline void set_d(Mat& m, int y, int x, double v) {
    switch (m.depth()) {
    case CV_8U: m.at<uchar>(y, x) = saturate_cast<uchar>(v); return;
    case CV_32S: m.at<int>(y, x) = saturate_cast<int>(v); return;
    case CV_32F: m.at<float>(y, x) = (float)v; return;
    case CV_64F: m.at<double>(y, x) = v; return;
    }
    STUB_FAIL("unsupported depth");
}

Mat& Mat::setTo(const Scalar& s) {
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) set_d(*this, y, x, s[0]);
    return *this;
}

void Mat::convertTo(Mat& m, int rtype, double alpha, double beta) const {
    if (channels() != 1) STUB_FAIL("convertTo: single channel only");
    rtype = CV_MAT_DEPTH(rtype < 0 ? flags : rtype);
    Mat src = *this; /* keeps the pixels alive when m is this */
    Mat dst;
    if (m.data && m.data != src.data && m.rows == rows && m.cols == cols && m.flags == rtype) dst = m;
    else dst.create(rows, cols, rtype);
    const bool plain = alpha == 1 && beta == 0;
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            if (plain) set_d(dst, y, x, get_d(src, y, x));
            else if (rtype == CV_32F && src.depth() != CV_64F) /* cvtScale with float work type */
                dst.at<float>(y, x) = (float)get_d(src, y, x) * (float)alpha + (float)beta;
            else set_d(dst, y, x, get_d(src, y, x) * alpha + beta);
        }
    m = dst;
}

Mat Mat::reshape(int, int) const { STUB_FAIL("Mat::reshape is only reached with lens distortion (Frame.cc:425)"); }

double Mat::dot(const Mat& m) const { /* dotProd_: double accumulation */
    double s = 0;
    if (rows * cols != m.rows * m.cols) STUB_FAIL("dot: size");
    for (int i = 0; i < rows * cols; i++) s += get_d(*this, i / cols, i % cols) * get_d(m, i / m.cols, i % m.cols);
    return s;
}

/* ------------------------------------------------------------------ norm (stat.cpp: double accumulators, element order) */
double norm(InputArray _a, int normType) {
    Mat a = _a.getMat();
    double s = 0;
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) {
            const double v = get_d(a, y, x);
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(v);
            else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
double norm(InputArray _a, InputArray _b, int normType) {
    Mat a = _a.getMat(), b = _b.getMat();
    if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type()) STUB_FAIL("norm: size/type");
    double s = 0;
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) {
            /* the difference is formed in the element type (float for CV_32F), accumulated in double */
            double v = a.depth() == CV_32F ? (double)(a.at<float>(y, x) - b.at<float>(y, x)) : get_d(a, y, x) - get_d(b, y, x);
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(v);
            else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}

void transpose(InputArray _a, OutputArray _d) {
    Mat a = _a.getMat();
    Mat d(a.cols, a.rows, a.type());
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) std::memcpy(d.data + d.step * x + y * a.elemSize(), a.data + a.step * y + x * a.elemSize(), a.elemSize());
    *_d.m = d;
}

/* ------------------------------------------------------------------ gemm (matmul.cpp) */
void gemm(InputArray _a, InputArray _b, double alpha, InputArray _c, double beta, OutputArray _d, int flags) {
    Mat A = _a.getMat(), B = _b.getMat(), C = beta != 0 ? _c.getMat() : Mat();
    const int depth = A.depth();
    if (depth != CV_32F && depth != CV_64F) STUB_FAIL("gemm: float types only");
    if (B.depth() != depth || (C.data && C.depth() != depth)) STUB_FAIL("gemm: mixed types");
    const bool aT = flags & GEMM_1_T, bT = flags & GEMM_2_T, cT = flags & GEMM_3_T;
    const int M = aT ? A.cols : A.rows, K = aT ? A.rows : A.cols, N = bT ? B.rows : B.cols;
    if ((bT ? B.cols : B.rows) != K) STUB_FAIL("gemm: inner dimension");
    if (C.data && ((cT ? C.cols : C.rows) != M || (cT ? C.rows : C.cols) != N)) STUB_FAIL("gemm: C size");
    Mat D(M, N, depth); /* fresh: D may alias an operand */
    /* small-matrix special case: flags == 0, inner dimension 2..4 equal to a side of D -> the dot product is evaluated
     * in the element type, left to right; alpha / beta are applied in double. */
    const bool small = flags == 0 && K >= 2 && K <= 4 && (K == N || K == M);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double s;
            if (depth == CV_32F) {
                if (small) {
                    float t = A.at<float>(i, 0) * B.at<float>(0, j);
                    for (int k = 1; k < K; k++) t = t + A.at<float>(i, k) * B.at<float>(k, j);
                    s = (double)t;
                } else {
                    s = 0;
                    for (int k = 0; k < K; k++)
                        s += (double)(aT ? A.at<float>(k, i) : A.at<float>(i, k)) * (double)(bT ? B.at<float>(j, k) : B.at<float>(k, j));
                }
                s *= alpha;
                if (C.data) s += beta * (double)(cT ? C.at<float>(j, i) : C.at<float>(i, j));
                D.at<float>(i, j) = (float)s;
            } else {
                s = 0;
                if (small) {
                    s = A.at<double>(i, 0) * B.at<double>(0, j);
                    for (int k = 1; k < K; k++) s = s + A.at<double>(i, k) * B.at<double>(k, j);
                } else
                    for (int k = 0; k < K; k++)
                        s += (aT ? A.at<double>(k, i) : A.at<double>(i, k)) * (bT ? B.at<double>(j, k) : B.at<double>(k, j));
                s *= alpha;
                if (C.data) s += beta * (cT ? C.at<double>(j, i) : C.at<double>(i, j));
                D.at<double>(i, j) = s;
            }
        }
    Mat& dst = *_d.m;
    if (dst.data && dst.rows == M && dst.cols == N && dst.flags == D.flags) D.copyTo(dst);
    else dst = D;
}

/* ------------------------------------------------------------------ MatExpr (matop.cpp) */
static MatExpr mk_addex(const Mat& a, const Mat& b, double alpha, double beta, double s = 0) {
    MatExpr e; e.op = MatExpr::ADDEX; e.a = a; e.b = b; e.alpha = alpha; e.beta = beta; e.s = s; return e;
}
static MatExpr mk_t(const Mat& a, double alpha) { MatExpr e; e.op = MatExpr::T; e.a = a; e.alpha = alpha; return e; }
static MatExpr mk_gemm(int flags, const Mat& a, const Mat& b, double alpha, const Mat& c = Mat(), double beta = 1) {
    MatExpr e; e.op = MatExpr::GEMM; e.flags = flags; e.a = a; e.b = b; e.alpha = alpha; e.c = c; e.beta = beta; return e;
}
static MatExpr mk_init(int kind, int r, int c, int type) {
    MatExpr e; e.op = MatExpr::INIT; e.flags = kind; e.irows = r; e.icols = c; e.itype = type; e.alpha = 1; return e;
}
static inline bool isIdentity(const MatExpr& e) { return e.op == MatExpr::IDENT; }
static inline bool isAddEx(const MatExpr& e) { return e.op == MatExpr::ADDEX; }
static inline bool isScaled(const MatExpr& e) { return isAddEx(e) && (!e.b.data || e.beta == 0) && e.s == 0; }
static inline bool isT(const MatExpr& e) { return e.op == MatExpr::T; }
static inline bool isMatProd(const MatExpr& e) { return e.op == MatExpr::GEMM && (!e.c.data || e.beta == 0); }

void MatExpr::assign(Mat& m) const {
    switch (op) {
    case IDENT: m = a; return;
    case INIT: {
        m.create(irows, icols, itype); /* keeps the destination's pixels when size and type already match */
        for (int y = 0; y < irows; y++)
            for (int x = 0; x < icols; x++) set_d(m, y, x, flags == '0' ? 0. : flags == '1' ? alpha : (x == y ? alpha : 0.));
        return;
    }
    case T: {
        Mat d;
        cv::transpose(a, d);
        if (alpha != 1) d.convertTo(d, d.type(), alpha);
        m = d;
        return;
    }
    case GEMM: cv::gemm(a, b, alpha, c, beta, m, flags); return;
    case ADDEX: {
        if (!b.data) {
            if (s == 0) { a.convertTo(m, a.type(), alpha); return; }
            Mat d(a.rows, a.cols, a.type());
            for (int y = 0; y < a.rows; y++)
                for (int x = 0; x < a.cols; x++) set_d(d, y, x, get_d(a, y, x) * alpha + s);
            m = d;
            return;
        }
        if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type()) STUB_FAIL("a + b: size/type");
        Mat d(a.rows, a.cols, a.type());
        for (int y = 0; y < a.rows; y++)
            for (int x = 0; x < a.cols; x++) {
                if (a.depth() == CV_32F) { /* add / subtract / scaleAdd / addWeighted work in float */
                    const float fa = a.at<float>(y, x), fb = b.at<float>(y, x);
                    float r;
                    if (alpha == 1 && beta == 1) r = fa + fb;
                    else if (alpha == 1 && beta == -1) r = fa - fb;
                    else if (alpha == -1 && beta == 1) r = fb - fa;
                    else if (alpha == 1) r = fb * (float)beta + fa;
                    else if (beta == 1) r = fa * (float)alpha + fb;
                    else r = fa * (float)alpha + fb * (float)beta;
                    d.at<float>(y, x) = r + (float)s;
                } else
                    set_d(d, y, x, get_d(a, y, x) * alpha + get_d(b, y, x) * beta + s);
            }
        m = d;
        return;
    }
    case MUL: {
        Mat d(a.rows, a.cols, a.type());
        for (int y = 0; y < a.rows; y++)
            for (int x = 0; x < a.cols; x++) set_d(d, y, x, get_d(a, y, x) * get_d(b, y, x) * alpha);
        m = d;
        return;
    }
    case INV: STUB_FAIL("Mat::inv is not on the hot path");
    }
}
Size MatExpr::size() const {
    switch (op) {
    case INIT: return Size(icols, irows);
    case T: return Size(a.rows, a.cols);
    case GEMM: return Size((flags & GEMM_2_T) ? b.rows : b.cols, (flags & GEMM_1_T) ? a.cols : a.rows);
    default: return a.size();
    }
}
int MatExpr::type() const { return op == INIT ? itype : a.type(); }

MatExpr Mat::t() const { return mk_t(*this, 1); }
MatExpr Mat::inv(int method) const { MatExpr e; e.op = MatExpr::INV; e.a = *this; e.flags = method; return e; }
MatExpr Mat::mul(const Mat& m, double scale) const { MatExpr e; e.op = MatExpr::MUL; e.a = *this; e.b = m; e.alpha = scale; return e; }
MatExpr Mat::zeros(int r, int c, int type) { return mk_init('0', r, c, type); }
MatExpr Mat::ones(int r, int c, int type) { return mk_init('1', r, c, type); }
MatExpr Mat::eye(int r, int c, int type) { return mk_init('I', r, c, type); }
MatExpr Mat::zeros(Size s, int type) { return mk_init('0', s.height, s.width, type); }
MatExpr Mat::ones(Size s, int type) { return mk_init('1', s.height, s.width, type); }

MatExpr MatExpr::t() const {
    if (op == T) return alpha == 1 ? MatExpr(a) : mk_addex(a, Mat(), alpha, 0);
    if (isScaled(*this)) return mk_t(a, alpha);
    return mk_t(Mat(*this), 1);
}

/* e * s : MatOp::multiply(e, s) and its overrides */
static MatExpr scale_expr(const MatExpr& e, double s) {
    MatExpr r = e;
    switch (e.op) {
    case MatExpr::ADDEX: r.alpha *= s; r.beta *= s; r.s *= s; return r;
    case MatExpr::T: r.alpha *= s; return r;
    case MatExpr::GEMM: r.alpha *= s; r.beta *= s; return r;
    case MatExpr::INIT: r.alpha *= s; return r;
    default: return mk_addex(Mat(e), Mat(), s, 0);
    }
}
/* e1 + sign * e2 : MatOp_GEMM::add / subtract first, then the generic MatOp::add / subtract */
static MatExpr add_expr(const MatExpr& e1, const MatExpr& e2, double sign) {
    const bool i1 = isIdentity(e1), i2 = isIdentity(e2);
    const double alpha1 = i1 ? 1 : e1.alpha, alpha2 = (i2 ? 1 : e2.alpha) * sign;
    if (isMatProd(e1) && (i2 || isScaled(e2) || isT(e2)))
        return mk_gemm(e1.flags | (isT(e2) ? GEMM_3_T : 0), e1.a, e1.b, alpha1, e2.a, alpha2);
    if (isMatProd(e2) && (i1 || isScaled(e1) || isT(e1)))
        return mk_gemm(e2.flags | (isT(e1) ? GEMM_3_T : 0), e2.a, e2.b, alpha2, e1.a, alpha1);
    double alpha = 1, beta = sign, s = 0;
    Mat m1, m2;
    if (isAddEx(e1) && (!e1.b.data || e1.beta == 0)) { m1 = e1.a; alpha = e1.alpha; s = e1.s; } else e1.assign(m1);
    if (isAddEx(e2) && (!e2.b.data || e2.beta == 0)) { m2 = e2.a; beta = e2.alpha * sign; s += e2.s * sign; } else e2.assign(m2);
    return mk_addex(m1, m2, alpha, beta, s);
}
/* e1 * e2 : MatOp::matmul */
static MatExpr matmul_expr(const MatExpr& e1, const MatExpr& e2) {
    double scale = 1;
    int flags = 0;
    Mat m1, m2;
    if (isT(e1)) { flags = GEMM_1_T; scale = e1.alpha; m1 = e1.a; }
    else if (isScaled(e1)) { scale = e1.alpha; m1 = e1.a; }
    else e1.assign(m1);
    if (isT(e2)) { flags |= GEMM_2_T; scale *= e2.alpha; m2 = e2.a; }
    else if (isScaled(e2)) { scale *= e2.alpha; m2 = e2.a; }
    else e2.assign(m2);
    return mk_gemm(flags, m1, m2, scale);
}

MatExpr operator+(const Mat& a, const Mat& b) { return mk_addex(a, b, 1, 1); }
MatExpr operator+(const Mat& a, const MatExpr& e) { return add_expr(e, MatExpr(a), 1); }
MatExpr operator+(const MatExpr& e, const Mat& b) { return add_expr(e, MatExpr(b), 1); }
MatExpr operator+(const MatExpr& e1, const MatExpr& e2) { return add_expr(e1, e2, 1); }
MatExpr operator+(const Mat& a, const Scalar& s) { return mk_addex(a, Mat(), 1, 0, s[0]); }
MatExpr operator-(const Mat& a, const Mat& b) { return mk_addex(a, b, 1, -1); }
MatExpr operator-(const Mat& a, const MatExpr& e) { return add_expr(MatExpr(a), e, -1); }
MatExpr operator-(const MatExpr& e, const Mat& b) { return add_expr(e, MatExpr(b), -1); }
MatExpr operator-(const MatExpr& e1, const MatExpr& e2) { return add_expr(e1, e2, -1); }
MatExpr operator-(const Mat& m) { return mk_addex(m, Mat(), -1, 0); }
/* MatOp::subtract(Scalar(0), e): only AddEx negates in place - every other node (a transpose, a product) is
 * materialised first, so `-R.t()*t` multiplies by the *stored* transpose with flags == 0. */
MatExpr operator-(const MatExpr& e) {
    if (isAddEx(e)) { MatExpr r = e; r.alpha = -r.alpha; r.beta = -r.beta; r.s = -r.s; return r; }
    return mk_addex(Mat(e), Mat(), -1, 0);
}
MatExpr operator*(const Mat& a, const Mat& b) { return mk_gemm(0, a, b, 1); }
MatExpr operator*(const Mat& a, const MatExpr& e) { return matmul_expr(MatExpr(a), e); }
MatExpr operator*(const MatExpr& e, const Mat& b) { return matmul_expr(e, MatExpr(b)); }
MatExpr operator*(const MatExpr& e1, const MatExpr& e2) { return matmul_expr(e1, e2); }
MatExpr operator*(const Mat& a, double s) { return mk_addex(a, Mat(), s, 0); }
MatExpr operator*(double s, const Mat& a) { return mk_addex(a, Mat(), s, 0); }
MatExpr operator*(const MatExpr& e, double s) { return scale_expr(e, s); }
MatExpr operator*(double s, const MatExpr& e) { return scale_expr(e, s); }
MatExpr operator/(const Mat& a, double s) { return mk_addex(a, Mat(), 1. / s, 0); }
MatExpr operator/(const MatExpr& e, double s) { return scale_expr(e, 1. / s); }

float fastAtan2(float y, float x) { return oracle_fast_atan2(y, x); }

/* ------------------------------------------------------------------ imgproc */
void resize(InputArray _src, OutputArray _dst, Size dsize, double, double, int interpolation) {
    Mat src = _src.getMat();
    if (src.type() != CV_8UC1 || interpolation != INTER_LINEAR || dsize.area() == 0) STUB_FAIL("resize: u8 INTER_LINEAR with dsize only");
    _dst.create(dsize, src.type()); /* an ROI header of the right size keeps its pixels (ORBextractor.cc:1120) */
    Mat dst = _dst.getMat();
    oracle_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}
void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType, const Scalar&) {
    Mat src = _src.getMat();
    if (src.type() != CV_8UC1) STUB_FAIL("copyMakeBorder: u8 only");
    if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) STUB_FAIL("copyMakeBorder: REFLECT_101 only");
    if (!(borderType & BORDER_ISOLATED) && src.datastart && src.step) {
        /* a submatrix is extended with the real pixels around it first (copy.cpp: locateROI) */
        const size_t ofs = src.data - src.datastart;
        const int oy = (int)(ofs / src.step), ox = (int)(ofs % src.step);
        const int wrows = (int)((src.dataend - src.datastart) / src.step), wcols = (int)src.step;
        const int dtop = std::min(oy, top), dleft = std::min(ox, left);
        const int dbottom = std::min(wrows - src.rows - oy, bottom), dright = std::min(wcols - src.cols - ox, right);
        if (dtop > 0 || dleft > 0 || dbottom > 0 || dright > 0) {
            Mat ext = src;
            ext.data -= (size_t)std::max(dtop, 0) * src.step + std::max(dleft, 0);
            ext.rows += std::max(dtop, 0) + std::max(dbottom, 0);
            ext.cols += std::max(dleft, 0) + std::max(dright, 0);
            top -= std::max(dtop, 0); left -= std::max(dleft, 0); bottom -= std::max(dbottom, 0); right -= std::max(dright, 0);
            src = ext;
        }
    }
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    /* interior first (a no-op when src is the ROI of dst, ORBextractor.cc:1122), then the frame from interior pixels */
    for (int y = 0; y < src.rows; y++) {
        uchar* d = dst.data + (size_t)(y + top) * dst.step + left;
        const uchar* s = src.data + (size_t)y * src.step;
        if (d != s) std::memmove(d, s, src.cols);
    }
    for (int y = 0; y < dst.rows; y++) {
        const bool frame_row = y < top || y >= top + src.rows;
        const uchar* s = src.data + (size_t)reflect101(y - top, src.rows) * src.step;
        uchar* d = dst.data + (size_t)y * dst.step;
        for (int x = 0; x < dst.cols; x++) {
            if (!frame_row && x >= left && x < left + src.cols) { x = left + src.cols - 1; continue; }
            d[x] = s[reflect101(x - left, src.cols)];
        }
    }
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sx, double sy, int borderType) {
    Mat src = _src.getMat();
    if (src.type() != CV_8UC1 || ksize.width != 7 || ksize.height != 7 || sx != 2 || sy != 2 || borderType != BORDER_REFLECT_101)
        STUB_FAIL("GaussianBlur: u8 7x7 sigma 2 REFLECT_101 only (ORBextractor.cc:1086)");
    _dst.create(src.size(), src.type());
    Mat dst = _dst.getMat();
    oracle_gaussian7_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step); /* reads everything first */
}

void undistortPoints(InputArray, OutputArray, InputArray, InputArray, InputArray, InputArray) {
    STUB_FAIL("undistortPoints is only reached with lens distortion (Frame.cc:426)");
}

/* ------------------------------------------------------------------ features2d */
void FAST(InputArray _image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    Mat img = _image.getMat();
    if (img.type() != CV_8UC1 || !nonmaxSuppression) STUB_FAIL("FAST: u8 with NMS only");
    keypoints.clear();
    if (img.rows < 7 || img.cols < 7) return;
    std::vector<int32_t> out((size_t)img.rows * img.cols * 3 / 2 + 3);
    const int n = oracle_fast_detect(img.data, img.cols, img.rows, (int)img.step, threshold, out.data(), (int)(out.size() / 3));
    keypoints.reserve(n);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)out[3 * i], (float)out[3 * i + 1], 7.f, -1, (float)out[3 * i + 2]));
}
void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) { STUB_FAIL("retainBest: ComputeKeyPointsOld is dead code"); }

} // namespace cv

/* ------------------------------------------------------------------ C API (PnPsolver.cc) */
CvMat* cvCreateMat(int rows, int cols, int type) {
    if (CV_MAT_DEPTH(type) != CV_64F) STUB_FAIL("cvCreateMat: CV_64F only");
    CvMat* m = new CvMat(cvMat(rows, cols, type, new double[(size_t)rows * cols]));
    return m;
}
void cvReleaseMat(CvMat** m) {
    if (m && *m) { delete[] (*m)->data.db; delete *m; *m = 0; }
}
void cvSetZero(CvMat* m) { std::memset(m->data.ptr, 0, (size_t)m->step * m->rows); }
void cvMulTransposed(const CvMat* src, CvMat* dst, int order) {
    if (order != 1) STUB_FAIL("cvMulTransposed: order 1 only");
    oracle_mul_transposed(src->data.db, src->rows, src->cols, dst->data.db);
}
void cvSVD(CvMat* A, CvMat* W, CvMat* U, CvMat* V, int flags) {
    const int m = A->rows, n = A->cols;
    if (m < n) STUB_FAIL("cvSVD: m >= n only");
    std::vector<double> Ut((size_t)n * m), Wv(n), Vt((size_t)n * n);
    oracle_svd(A->data.db, m, n, Ut.data(), Wv.data(), Vt.data());
    for (int i = 0; i < n; i++) W->data.db[i] = Wv[i];
    if (U) {
        if (m != n) STUB_FAIL("cvSVD: U only for square matrices");
        for (int i = 0; i < n; i++)
            for (int k = 0; k < n; k++) U->data.db[(flags & CV_SVD_U_T) ? i * n + k : k * n + i] = Ut[i * m + k];
    }
    if (V)
        for (int i = 0; i < n; i++)
            for (int k = 0; k < n; k++) V->data.db[(flags & CV_SVD_V_T) ? i * n + k : k * n + i] = Vt[i * n + k];
}
int cvSolve(const CvMat* A, const CvMat* b, CvMat* x, int method) {
    if (method != CV_SVD || A->cols > 6 || A->rows > 6) STUB_FAIL("cvSolve: CV_SVD up to 6x6");
    oracle_svd_solve(A->data.db, A->rows, A->cols, b->data.db, x->data.db);
    return 1;
}
double cvInvert(const CvMat* A, CvMat* inv, int method) {
    if (method != CV_SVD || A->rows != 3 || A->cols != 3) STUB_FAIL("cvInvert: CV_SVD 3x3");
    oracle_svd_invert3(A->data.db, inv->data.db);
    return 1;
}
