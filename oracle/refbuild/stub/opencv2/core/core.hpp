/* STUB OpenCV — TEST INFRASTRUCTURE ONLY (oracle/refbuild).
 *
 * A minimal, single-channel stand-in for the OpenCV 2.4 C++ API surface that the reference's hot-path
 * sources use, so that they compile UNMODIFIED from /root/reference with g++ alone (no OpenCV in the image):
 *   corbslam_client/src/ORBextractor.cc, ORBmatcher.cc, Frame.cc, PnPsolver.cc and Thirdparty/DBoW2.
 * Containers and expression templates are restated here; every piece of OpenCV *arithmetic* whose rounding is
 * observable (FAST, resize, GaussianBlur, fastAtan2, gemm on CV_32F, SVD / solve / invert) is delegated to the
 * models in oracle/*.cpp that are pinned bit-exactly against the real OpenCV 4.13 (tests/golden/*.npz,
 * tools/gen_golden*.py). Nothing under corb_slam_b200/ includes this.
 *
 * cv::MatExpr follows the structure of OpenCV's matop.cpp (lazy AddEx / T / GEMM / Initializer nodes): which
 * gemm call an expression such as `-Rcw.t()*tcw` or `Rcw*x+tcw` becomes decides the float rounding of the result.
 */
#ifndef CORB_REFSTUB_OPENCV_CORE_HPP
#define CORB_REFSTUB_OPENCV_CORE_HPP

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> 3) & 511) + 1)
#define CV_MAKETYPE(d, cn) (CV_MAT_DEPTH(d) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_GEMM_A_T 1
#define CV_GEMM_B_T 2
#define CV_GEMM_C_T 4

/* cvRound: SSE2 cvtsd2si = round-half-even in the default rounding mode */
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

using std::string;
using std::vector;

template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
};
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <> template <> inline Point_<float>::Point_(const Point_<int>& p) : x((float)p.x), y((float)p.y) {}
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size2i;
typedef Size2i Size;
inline bool operator==(const Size& a, const Size& b) { return a.width == b.width && a.height == b.height; }
inline bool operator!=(const Size& a, const Size& b) { return !(a == b); }

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1 = 0, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar_<T> all(T v) { return Scalar_<T>(v, v, v, v); }
    T operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT_MIN, INT_MAX); }
};

/* field order and size (28 B) of cv::KeyPoint */
class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1 };
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };

class Mat;
class MatExpr;
class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef InputArray InputArrayOfArrays;

class Mat {
public:
    enum { AUTO_STEP = 0 };
    Mat() : flags(0), dims(0), rows(0), cols(0), step(0), data(0), datastart(0), dataend(0) {}
    Mat(int r, int c, int type) : flags(0), dims(0), rows(0), cols(0), step(0), data(0), datastart(0), dataend(0) { create(r, c, type); }
    Mat(Size sz, int type) : flags(0), dims(0), rows(0), cols(0), step(0), data(0), datastart(0), dataend(0) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : flags(0), dims(0), rows(0), cols(0), step(0), data(0), datastart(0), dataend(0) {
        create(r, c, type);
        setTo(s);
    }
    /* user-allocated data: no ownership */
    Mat(int r, int c, int type, void* d, size_t st = AUTO_STEP) : flags(type), dims(2), rows(r), cols(c), data((uchar*)d) {
        step = st ? st : (size_t)c * elemSize();
        datastart = data;
        dataend = data + step * r;
    }
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Scalar& s) { return setTo(s); }

    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && flags == type) return;
        flags = type;
        dims = 2;
        rows = r;
        cols = c;
        step = (size_t)c * elemSize();
        size_t bytes = step * (size_t)r;
        buf.reset(new uchar[bytes + 64], std::default_delete<uchar[]>());
        data = buf.get();
        datastart = data;
        dataend = data + bytes;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf.reset(); data = 0; datastart = dataend = 0; rows = cols = 0; step = 0; dims = 0; }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& m) const {
        m.create(rows, cols, flags);
        if (m.data == data && m.step == step) return;
        size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; y++) std::memmove(m.data + y * m.step, data + y * step, rb);
    }
    void copyTo(OutputArray dst) const;
    void convertTo(Mat& m, int rtype, double alpha = 1, double beta = 0) const;
    void convertTo(OutputArray dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat& setTo(const Scalar& s);

    Mat rowRange(int r0, int r1) const { Mat m(*this); m.data = data + step * r0; m.rows = r1 - r0; return m; }
    Mat colRange(int c0, int c1) const { Mat m(*this); m.data = data + elemSize() * c0; m.cols = c1 - c0; return m; }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Mat operator()(const Range& rr, const Range& cr) const {
        Mat m(*this);
        if (rr.start != INT_MIN) m = m.rowRange(rr.start, rr.end);
        if (cr.start != INT_MIN) m = m.colRange(cr.start, cr.end);
        return m;
    }
    Mat reshape(int cn, int nrows = 0) const;

    template <typename T> T& at(int i, int j) { return *(T*)(data + step * i + sizeof(T) * j); }
    template <typename T> const T& at(int i, int j) const { return *(const T*)(data + step * i + sizeof(T) * j); }
    /* single index: element i of a row or column vector (row-major walk otherwise) */
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    template <typename T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }
    uchar* ptr(int y = 0) { return data + step * y; }
    const uchar* ptr(int y = 0) const { return data + step * y; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + step * y); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + step * y); }

    bool empty() const { return data == 0 || rows * cols == 0; }
    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 0}; return sz[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows == 1 || step == (size_t)cols * elemSize(); }
    Mat getMat() const { return *this; }

    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const;
    MatExpr mul(const Mat& m, double scale = 1) const;
    double dot(const Mat& m) const;
    static MatExpr zeros(int r, int c, int type);
    static MatExpr ones(int r, int c, int type);
    static MatExpr eye(int r, int c, int type);
    static MatExpr zeros(Size s, int type);
    static MatExpr ones(Size s, int type);

    int flags, dims, rows, cols;
    size_t step;
    uchar* data;
    const uchar *datastart, *dataend; /* the whole allocation (an ROI keeps them) */
    std::shared_ptr<uchar> buf;       /* reference count of the allocation */
};

/* OpenCV passes arrays through proxy classes so that temporaries (ROI headers) can be outputs */
class _InputArray {
public:
    _InputArray() : m(0) {}
    _InputArray(const Mat& _m) : m(const_cast<Mat*>(&_m)) {}
    _InputArray(const MatExpr& e);
    Mat getMat() const { return m ? *m : Mat(); }
    bool empty() const { return !m || m->empty(); }
    Size size() const { return m ? m->size() : Size(); }
    int type() const { return m ? m->type() : 0; }
    Mat* m;
    mutable std::shared_ptr<Mat> hold;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& _m) : _InputArray(_m) {}
    _OutputArray(const Mat& _m) : _InputArray(_m) {} /* a temporary header that shares the caller's pixels */
    void create(int r, int c, int type) const { m->create(r, c, type); }
    void create(Size sz, int type) const { m->create(sz, type); }
    void release() const { m->release(); }
    bool needed() const { return m != 0; }
    Mat& getMatRef() const { return *m; }
};
inline InputArray noArray() { static _InputArray none; return none; }

/* ---- lazy expressions (structure of OpenCV matop.cpp) ---- */
class MatExpr {
public:
    enum Op { IDENT, ADDEX, T, GEMM, INIT, INV, MUL };
    MatExpr() : op(IDENT), flags(0), alpha(1), beta(0), s(0) {}
    MatExpr(const Mat& m) : op(IDENT), flags(0), a(m), alpha(1), beta(0), s(0) {}
    operator Mat() const { Mat m; assign(m); return m; }
    void assign(Mat& m) const;
    Size size() const;
    int type() const;
    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const { return Mat(*this).inv(method); }
    MatExpr mul(const Mat& m, double scale = 1) const { return Mat(*this).mul(m, scale); }
    double dot(const Mat& m) const { return Mat(*this).dot(m); }
    Mat row(int y) const { return Mat(*this).row(y); }
    Mat col(int x) const { return Mat(*this).col(x); }
    template <typename Tp> Tp at(int i, int j = 0) const { Mat m(*this); return j ? m.at<Tp>(i, j) : m.at<Tp>(i); }

    Op op;
    int flags;     /* GEMM: transposition flags; INIT: '0','1','I'; INV: method */
    Mat a, b, c;
    double alpha, beta, s;
    int irows, icols, itype; /* INIT */
};

MatExpr operator+(const Mat& a, const Mat& b);
MatExpr operator+(const Mat& a, const MatExpr& e);
MatExpr operator+(const MatExpr& e, const Mat& b);
MatExpr operator+(const MatExpr& e1, const MatExpr& e2);
MatExpr operator+(const Mat& a, const Scalar& s);
MatExpr operator-(const Mat& a, const Mat& b);
MatExpr operator-(const Mat& a, const MatExpr& e);
MatExpr operator-(const MatExpr& e, const Mat& b);
MatExpr operator-(const MatExpr& e1, const MatExpr& e2);
MatExpr operator-(const Mat& m);
MatExpr operator-(const MatExpr& e);
MatExpr operator*(const Mat& a, const Mat& b);
MatExpr operator*(const Mat& a, const MatExpr& e);
MatExpr operator*(const MatExpr& e, const Mat& b);
MatExpr operator*(const MatExpr& e1, const MatExpr& e2);
MatExpr operator*(const Mat& a, double s);
MatExpr operator*(double s, const Mat& a);
MatExpr operator*(const MatExpr& e, double s);
MatExpr operator*(double s, const MatExpr& e);
MatExpr operator/(const Mat& a, double s);
MatExpr operator/(const MatExpr& e, double s);

inline Mat::Mat(const MatExpr& e) : flags(0), dims(0), rows(0), cols(0), step(0), data(0), datastart(0), dataend(0) { e.assign(*this); }
inline Mat& Mat::operator=(const MatExpr& e) { e.assign(*this); return *this; }
inline _InputArray::_InputArray(const MatExpr& e) : hold(new Mat(e)) { m = hold.get(); }
inline void Mat::copyTo(OutputArray dst) const { copyTo(*dst.m); }
inline void Mat::convertTo(OutputArray dst, int rtype, double al, double be) const { convertTo(*dst.m, rtype, al, be); }

template <typename Tp> class MatCommaInitializer_;
template <typename Tp> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c);
    Mat_(const Mat& m) : Mat(m) {}
    Tp& operator()(int i, int j) { return this->template at<Tp>(i, j); }
    const Tp& operator()(int i, int j) const { return this->template at<Tp>(i, j); }
};
template <typename Tp> struct DepthOf;
template <> struct DepthOf<uchar> { enum { v = CV_8U }; };
template <> struct DepthOf<int> { enum { v = CV_32S }; };
template <> struct DepthOf<float> { enum { v = CV_32F }; };
template <> struct DepthOf<double> { enum { v = CV_64F }; };
template <typename Tp> inline Mat_<Tp>::Mat_(int r, int c) : Mat(r, c, DepthOf<Tp>::v) {}
template <typename Tp> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(Mat_<Tp>* _m) : m(_m), i(0) {}
    MatCommaInitializer_<Tp>& operator,(Tp v) { m->template at<Tp>((int)i) = v; i++; return *this; }
    operator Mat_<Tp>() const { return *m; }
    operator Mat() const { return *m; }
    Mat_<Tp>* m;
    size_t i;
};
template <typename Tp> inline MatCommaInitializer_<Tp> operator<<(const Mat_<Tp>& m, Tp v) {
    MatCommaInitializer_<Tp> ci(const_cast<Mat_<Tp>*>(&m));
    return (ci, v);
}

/* ---- free functions of core ---- */
double norm(InputArray a, int normType = NORM_L2);
double norm(InputArray a, InputArray b, int normType = NORM_L2);
void gemm(InputArray a, InputArray b, double alpha, InputArray c, double beta, OutputArray d, int flags = 0);
void transpose(InputArray a, OutputArray d);
float fastAtan2(float y, float x);

/* persistence: only what DBoW2's (unused here) YAML save/load needs to compile; every use aborts */
class FileNode {
public:
    FileNode operator[](const string&) const { std::abort(); }
    FileNode operator[](const char*) const { std::abort(); }
    FileNode operator[](int) const { std::abort(); }
    size_t size() const { std::abort(); }
    operator int() const { std::abort(); }
    operator float() const { std::abort(); }
    operator double() const { std::abort(); }
    operator string() const { std::abort(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const string&) const { std::abort(); }
    FileNode operator[](const char*) const { std::abort(); }
    void release() {}
};
template <typename Tp> inline FileStorage& operator<<(FileStorage& fs, const Tp&) { std::abort(); return fs; }

} // namespace cv

/* ---- the C API PnPsolver.cc uses (CvMat on caller memory, CV_64F only) ---- */
#define CV_SVD 1
#define CV_SVD_MODIFY_A 1
#define CV_SVD_U_T 2
#define CV_SVD_V_T 4
struct CvMat {
    int type, step, rows, cols;
    union { uchar* ptr; double* db; float* fl; } data;
};
inline CvMat cvMat(int rows, int cols, int type, void* data = 0) {
    CvMat m;
    m.type = type;
    m.rows = rows;
    m.cols = cols;
    m.step = cols * (CV_MAT_DEPTH(type) == CV_64F ? 8 : 4);
    m.data.ptr = (uchar*)data;
    return m;
}
inline double cvmGet(const CvMat* m, int r, int c) { return m->data.db[(size_t)r * (m->step / 8) + c]; }
inline void cvmSet(CvMat* m, int r, int c, double v) { m->data.db[(size_t)r * (m->step / 8) + c] = v; }
CvMat* cvCreateMat(int rows, int cols, int type);
void cvReleaseMat(CvMat** m);
void cvSetZero(CvMat* m);
void cvMulTransposed(const CvMat* src, CvMat* dst, int order);
void cvSVD(CvMat* A, CvMat* W, CvMat* U = 0, CvMat* V = 0, int flags = 0);
int cvSolve(const CvMat* A, const CvMat* b, CvMat* x, int method);
double cvInvert(const CvMat* A, CvMat* inv, int method);

#endif
