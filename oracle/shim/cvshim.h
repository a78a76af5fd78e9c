// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for the slice of the OpenCV C++ API that
// /root/reference/src/ORBextractor.cc uses, so that the reference file can be compiled UNMODIFIED
// in a container that has no OpenCV C++ (see oracle/build_ref.sh). The image-processing
// primitives are the cv2-pinned restatements of oracle/cvprim.h; everything else here is plain
// container plumbing (ref-counted Mat with ROI views, KeyPoint, Point, InputArray...).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <iterator>
#include <sstream>
#include <memory>
#include <string>
#include <vector>

#include "../cvprim.h"
#include "../cvprim_mat.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

static inline int cvRound(double v) { return cvprim::round_half_even(v); }
static inline int cvRound(float v) { return cvprim::round_half_even(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return cvprim::ifloor(v); }
static inline int cvCeil(double v) { return cvprim::iceil(v); }

namespace cv {

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <typename S> Point_& operator*=(S s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

// Matrix expressions. OpenCV evaluates cv::Mat arithmetic lazily (MatExpr): which fused primitive runs - and so how
// the floats round - depends on the expression's shape (Rcw*P+tcw is ONE gemm with C; -R.t()*t is a gemm with the
// transpose flag and alpha = -1; M/s is a convertTo by the float of 1./s ...). The forms the reference uses are modelled
// here after modules/core/src/matop.cpp; the arithmetic is cvprim_mat.h's (pinned to cv2 4.13).
// Mat::zeros assigned to a Mat of the right size zero-fills IN PLACE (computeDescriptors at src/ORBextractor.cc:1037
// relies on it to write through the rowRange view of the output descriptor matrix).
class Mat;
struct MatExpr;

static inline size_t cvshim_elem_size(int type) { return type == 5 ? 4 : (type == 6 ? 8 : 1); }

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    int mtype = CV_8U;  // CV_8U or CV_32F
    int cn = 1;         // channels: only reshape(2) / reshape(1) around cv::undistortPoints change it
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() {}
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type) { create(r, c, type); }
    // external data, not owned
    Mat(int r, int c, int type, void* ext, size_t stp = 0)
        : rows(r), cols(c), step(stp ? stp : (size_t)c * cvshim_elem_size(type)), data((uchar*)ext), mtype(type) {}
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);

    size_t elemSize() const { return cvshim_elem_size(mtype) * cn; }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == mtype && cn == 1) return;
        const size_t es = cvshim_elem_size(type);
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c * es);
        rows = r; cols = c; step = (size_t)c * es; data = buf->data(); mtype = type; cn = 1;
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    static MatExpr zeros(int r, int c, int type);
    static MatExpr ones(int r, int c, int type);
    static MatExpr eye(int r, int c, int type);
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return mtype; }
    int channels() const { return cn; }
    size_t total() const { return (size_t)rows * cols; }
    size_t step1() const { return step / cvshim_elem_size(mtype); }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows == 1; }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols * cn, mtype);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        m.cols = cols; m.cn = cn;
        return m;
    }
    // copyTo(OutputArray): create() keeps a destination of the right size and type, so copying into a temporary
    // ROI view writes through (src/KeyFrame.cc:95-96)
    void copyTo(const Mat& dst_) const {
        Mat& dst = const_cast<Mat&>(dst_);
        if (empty()) { dst.release(); return; }
        dst.create(rows, cols, mtype);
        for (int y = 0; y < rows; ++y) std::memmove(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
    }
    Mat view(int x, int y, int w, int h) const {
        Mat m; m.rows = h; m.cols = w; m.step = step; m.data = data + (size_t)y * step + (size_t)x * elemSize();
        m.buf = buf; m.mtype = mtype; m.cn = cn; return m;
    }
    Mat operator()(const Rect& r) const { return view(r.x, r.y, r.width, r.height); }
    Mat rowRange(int a, int b) const { return view(0, a, cols, b - a); }
    Mat colRange(int a, int b) const { return view(a, 0, b - a, rows); }
    Mat row(int r) const { return view(0, r, cols, 1); }
    Mat col(int c) const { return view(c, 0, 1, rows); }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + c * sizeof(T)); }
    // single index: element i of a row or column vector
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    // reshape(cn): same data, the row's scalars regrouped into cn-channel elements
    Mat reshape(int ncn) const { Mat m = *this; m.cols = cols * cn / ncn; m.cn = ncn; return m; }
    // convertTo: 8U -> 32F (exact) and 32F -> 32F scaling, x*(float)alpha + (float)beta. A destination of another
    // type gets a new buffer (also when it is the source itself, src/Frame.cc:564).
    void convertTo(Mat& dst, int type, double alpha = 1, double beta = 0) const {
        if (type < 0) type = mtype;
        Mat out(rows, cols, type);
        const float a = (float)alpha, b = (float)beta;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                const float v = mtype == CV_32F ? at<float>(y, x) : (float)at<uchar>(y, x);
                const float r = (alpha == 1 && beta == 0) ? v : v * a + b;
                if (type == CV_32F) out.at<float>(y, x) = r;
                else out.at<uchar>(y, x) = (uchar)std::min(255, std::max(0, cvprim::round_half_even(r)));
            }
        if (dst.data && dst.rows == rows && dst.cols == cols && dst.mtype == type && dst.data != data) out.copyTo(dst);
        else dst = out;
    }
    MatExpr t() const;
    double dot(const Mat& m) const {
        assert(mtype == CV_32F && m.mtype == CV_32F && rows == m.rows && cols == m.cols);
        return cvprim::dot_32f((const float*)data, step / 4, (const float*)m.data, m.step / 4, rows, cols);
    }
};

struct MatExpr {
    enum Kind { INIT, TRANSPOSE, ADDEX, GEMM };
    Kind kind = INIT;
    int init = 0;            // INIT: 0 zeros, 1 ones, 2 eye; scaled by alpha
    int rows = 0, cols = 0, type = 0;
    Mat a, b, c;             // TRANSPOSE: alpha*a^T; ADDEX: alpha*a + beta*b (b may be empty); GEMM: alpha*op(a)*op(b) + beta*c
    double alpha = 1, beta = 0;
    int flags = 0;

    bool scaled() const { return kind == ADDEX && (b.empty() || beta == 0); }
    Mat eval() const {
        Mat m;
        if (kind == INIT) {
            m.create(rows, cols, type);
            for (int y = 0; y < rows; ++y)
                for (int x = 0; x < cols; ++x) {
                    const double v = init == 0 ? 0.0 : (init == 1 || x == y ? alpha : 0.0);
                    if (type == CV_32F) m.at<float>(y, x) = (float)v; else m.at<uchar>(y, x) = (uchar)v;
                }
        } else if (kind == TRANSPOSE) {
            assert(a.mtype == CV_32F);
            m.create(a.cols, a.rows, CV_32F);
            const float s = (float)alpha;
            for (int y = 0; y < a.rows; ++y)
                for (int x = 0; x < a.cols; ++x) m.at<float>(x, y) = alpha == 1 ? a.at<float>(y, x) : a.at<float>(y, x) * s;
        } else if (kind == ADDEX) {
            assert(a.mtype == CV_32F);
            m.create(a.rows, a.cols, CV_32F);
            for (int y = 0; y < a.rows; ++y)
                for (int x = 0; x < a.cols; ++x) {
                    const float p = a.at<float>(y, x);
                    float r;
                    if (b.empty() || beta == 0) r = alpha == 1 ? p : p * (float)alpha;          // copy / convertTo
                    else {
                        const float q = b.at<float>(y, x);
                        if (alpha == 1 && beta == 1) r = p + q;                                  // cv::add
                        else if (alpha == 1 && beta == -1) r = p - q;                            // cv::subtract
                        else if (alpha == 1) r = q * (float)beta + p;                            // cv::scaleAdd(b, beta, a)
                        else if (beta == 1 && alpha == -1) r = q - p;
                        else if (beta == 1) r = p * (float)alpha + q;                            // cv::scaleAdd(a, alpha, b)
                        else r = p * (float)alpha + q * (float)beta;                             // cv::addWeighted
                    }
                    m.at<float>(y, x) = r;
                }
        } else {
            // a transposed SECOND operand (A*B.t()) takes another OpenCV kernel that is not modelled: no compiled reference
            // file uses it, so it is refused rather than guessed
            if (a.mtype != CV_32F || b.mtype != CV_32F || (flags & cvprim::GEMM_B_T)) { std::cerr << "cvshim: unsupported gemm form\n"; std::abort(); }
            const int mr = (flags & cvprim::GEMM_A_T) ? a.cols : a.rows, nc = (flags & cvprim::GEMM_B_T) ? b.rows : b.cols;
            m.create(mr, nc, CV_32F);
            cvprim::gemm32f((const float*)a.data, a.rows, a.cols, a.step / 4, (const float*)b.data, b.rows, b.cols, b.step / 4, alpha,
                            c.empty() ? nullptr : (const float*)c.data, c.step / 4, beta, (float*)m.data, m.step / 4, flags);
        }
        return m;
    }
    MatExpr t() const { Mat m = eval(); return m.t(); }
};

inline Mat::Mat(const MatExpr& e) { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e) {
    Mat v = e.eval();
    if (data && rows == v.rows && cols == v.cols && mtype == v.mtype) v.copyTo(*this);  // in place, through views
    else { Mat& self = *this; const Mat& src = v; self = src; }
    return *this;
}
inline MatExpr Mat::zeros(int r, int c, int type) { MatExpr e; e.kind = MatExpr::INIT; e.init = 0; e.rows = r; e.cols = c; e.type = type; return e; }
inline MatExpr Mat::ones(int r, int c, int type) { MatExpr e = zeros(r, c, type); e.init = 1; return e; }
inline MatExpr Mat::eye(int r, int c, int type) { MatExpr e = zeros(r, c, type); e.init = 2; return e; }
inline MatExpr Mat::t() const { MatExpr e; e.kind = MatExpr::TRANSPOSE; e.a = *this; return e; }

namespace shim_detail {
static inline MatExpr as_expr(const Mat& m) { MatExpr e; e.kind = MatExpr::ADDEX; e.a = m; return e; }
// operand of a matrix product: (matrix, transpose flag, scale) - MatOp::matmul
static inline void prod_operand(const MatExpr& e, Mat& m, bool& tr, double& s) {
    tr = false; s = 1;
    if (e.kind == MatExpr::TRANSPOSE) { m = e.a; tr = true; s = e.alpha; }
    else if (e.scaled()) { m = e.a; s = e.alpha; }
    else m = e.eval();
}
// operand of a sum: (matrix, scale) - MatOp::add / subtract
static inline void sum_operand(const MatExpr& e, Mat& m, double& s) {
    s = 1;
    if (e.scaled()) { m = e.a; s = e.alpha; }
    else m = e.eval();
}
static inline MatExpr matmul(const MatExpr& x, const MatExpr& y) {
    MatExpr r; r.kind = MatExpr::GEMM;
    bool ta, tb; double sa, sb;
    prod_operand(x, r.a, ta, sa); prod_operand(y, r.b, tb, sb);
    r.flags = (ta ? cvprim::GEMM_A_T : 0) | (tb ? cvprim::GEMM_B_T : 0);
    r.alpha = sa * sb;
    return r;
}
static inline MatExpr addsub(const MatExpr& x, const MatExpr& y, double sign) {
    // MatOp_GEMM::add: a product without C absorbs a plain / scaled matrix as its C term
    if (x.kind == MatExpr::GEMM && x.c.empty() && y.scaled()) { MatExpr r = x; r.c = y.a; r.beta = y.alpha * sign; return r; }
    if (y.kind == MatExpr::GEMM && y.c.empty() && x.scaled() && sign == 1) { MatExpr r = y; r.c = x.a; r.beta = x.alpha; return r; }
    MatExpr r; r.kind = MatExpr::ADDEX;
    double sb;
    sum_operand(x, r.a, r.alpha); sum_operand(y, r.b, sb);
    r.beta = sb * sign;
    return r;
}
static inline MatExpr scale(const MatExpr& e, double s) {
    MatExpr r = e;
    r.alpha *= s;
    if (e.kind == MatExpr::ADDEX || e.kind == MatExpr::GEMM) r.beta *= s;
    return r;
}
}  // namespace shim_detail

static inline MatExpr operator*(const Mat& a, const Mat& b) { return shim_detail::matmul(shim_detail::as_expr(a), shim_detail::as_expr(b)); }
static inline MatExpr operator*(const MatExpr& a, const Mat& b) { return shim_detail::matmul(a, shim_detail::as_expr(b)); }
static inline MatExpr operator*(const Mat& a, const MatExpr& b) { return shim_detail::matmul(shim_detail::as_expr(a), b); }
static inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return shim_detail::matmul(a, b); }
static inline MatExpr operator*(const Mat& a, double s) { return shim_detail::scale(shim_detail::as_expr(a), s); }
static inline MatExpr operator*(double s, const Mat& a) { return shim_detail::scale(shim_detail::as_expr(a), s); }
static inline MatExpr operator*(const MatExpr& a, double s) { return shim_detail::scale(a, s); }
static inline MatExpr operator*(double s, const MatExpr& a) { return shim_detail::scale(a, s); }
static inline MatExpr operator/(const Mat& a, double s) { return shim_detail::scale(shim_detail::as_expr(a), 1. / s); }
static inline MatExpr operator/(const MatExpr& a, double s) { return shim_detail::scale(a, 1. / s); }
static inline MatExpr operator-(const Mat& a) { return shim_detail::scale(shim_detail::as_expr(a), -1); }
static inline MatExpr operator-(const MatExpr& a) { return shim_detail::scale(a, -1); }
static inline MatExpr operator+(const Mat& a, const Mat& b) { return shim_detail::addsub(shim_detail::as_expr(a), shim_detail::as_expr(b), 1); }
static inline MatExpr operator+(const MatExpr& a, const Mat& b) { return shim_detail::addsub(a, shim_detail::as_expr(b), 1); }
static inline MatExpr operator+(const Mat& a, const MatExpr& b) { return shim_detail::addsub(shim_detail::as_expr(a), b, 1); }
static inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return shim_detail::addsub(a, b, 1); }
static inline MatExpr operator-(const Mat& a, const Mat& b) { return shim_detail::addsub(shim_detail::as_expr(a), shim_detail::as_expr(b), -1); }
static inline MatExpr operator-(const MatExpr& a, const Mat& b) { return shim_detail::addsub(a, shim_detail::as_expr(b), -1); }
static inline MatExpr operator-(const Mat& a, const MatExpr& b) { return shim_detail::addsub(shim_detail::as_expr(a), b, -1); }
static inline MatExpr operator-(const MatExpr& a, const MatExpr& b) { return shim_detail::addsub(a, b, -1); }

enum { NORM_L1 = 2, NORM_L2 = 4 };
static inline double norm(const Mat& m, int kind = NORM_L2) {
    assert(m.mtype == CV_32F && kind == NORM_L2); (void)kind;
    return cvprim::norm_l2_32f((const float*)m.data, m.rows, m.cols, m.step / 4);
}
static inline double norm(const MatExpr& e) { return norm(e.eval()); }
static inline double norm(const Mat& a, const Mat& b, int kind) {
    assert(a.mtype == CV_32F && b.mtype == CV_32F && kind == NORM_L1 && a.rows == b.rows && a.cols == b.cols); (void)kind;
    return cvprim::norm_l1_diff_32f((const float*)a.data, a.step / 4, (const float*)b.data, b.step / 4, a.rows, a.cols);
}
// cv::undistortPoints(src, dst, K, dist, R = noArray(), P): N x 1 CV_32FC2 (src/Frame.cc:421-423, 447-449)
static inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat& R, const Mat& P) {
    assert(src.mtype == CV_32F && src.cn == 2 && src.cols == 1 && R.empty() && P.data == K.data); (void)R; (void)P;
    Mat out(src.rows, 2, CV_32F);
    std::vector<float> in((size_t)src.rows * 2), d;
    for (int i = 0; i < src.rows; ++i) { in[2 * i] = src.ptr<float>(i)[0]; in[2 * i + 1] = src.ptr<float>(i)[1]; }
    for (int i = 0; i < (int)dist.total(); ++i) d.push_back(dist.at<float>(i));
    cvprim::undistort_points_32f(in.data(), (float*)out.data, src.rows, (const float*)K.data, K.step / 4, d.data(), (int)d.size());
    out.cols = 1; out.cn = 2;
    dst = out;
}

template <typename T> struct DataType_;
template <> struct DataType_<float> { enum { type = CV_32F }; };
template <> struct DataType_<uchar> { enum { type = CV_8U }; };
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, DataType_<T>::type) {}
};
template <typename T> struct MatCommaInitializer_ {
    Mat m; int idx;
    template <typename V> MatCommaInitializer_& operator,(V v) { ((T*)m.data)[idx++] = (T)v; return *this; }
    operator Mat() const { return m; }
};
template <typename T, typename V> static inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) {
    MatCommaInitializer_<T> ci{m, 0};
    return (ci, v);
}

// cv::FileStorage / cv::FileNode: only so that the (virtual, hence always instantiated) YAML save/load
// members of DBoW2::TemplatedVocabulary compile; the oracle loads vocabularies from text files only.
class FileNode {
public:
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator float() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    template <typename T> FileStorage& operator<<(const T&) { return *this; }
};

class _InputArray {
public:
    const Mat* m;
    _InputArray(const Mat& mm) : m(&mm) {}
    bool empty() const { return m->empty(); }
    Mat getMat() const { return *m; }
};
class _OutputArray {
public:
    Mat* m;
    _OutputArray(Mat& mm) : m(&mm) {}
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void release() const { m->release(); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

static inline float fastAtan2(float y, float x) { return cvprim::fast_atan2_deg(y, x); }

static inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int) {
    dst.create(sz.height, sz.width, 0);  // keeps an existing buffer (ROI) of the right size
    cvprim::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}
static inline void copyMakeBorder(const Mat& src, Mat& dst, int t, int b, int l, int r, int) {
    dst.create(src.rows + t + b, src.cols + l + r, 0);
    cvprim::copy_make_border_reflect101(src.data, src.cols, src.rows, src.step, dst.data, dst.step, t, b, l, r);
}
static inline void GaussianBlur(const Mat& src, Mat& dst, Size k, double sx, double sy, int) {
    assert(k.width == 7 && k.height == 7 && sx == 2 && sy == 2);
    (void)k; (void)sx; (void)sy;
    dst.create(src.rows, src.cols, 0);
    cvprim::gaussian7x7_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}
static inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmax) {
    assert(nonmax); (void)nonmax;
    std::vector<cvprim::FastKp> f;
    cvprim::fast9_nms(img.data, img.cols, img.rows, img.step, threshold, f);
    kps.clear();
    for (const auto& k : f) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1.f, (float)k.score));
}
struct KeyPointsFilter {
    // only referenced from the dead ComputeKeyPointsOld path (src/ORBextractor.cc:1006,1024)
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n >= 0 && (int)k.size() > n) {
            std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            k.resize(n);
        }
    }
};

}  // namespace cv
