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

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

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

// Mat::zeros yields an expression; assigning it to a Mat of the right size zero-fills IN PLACE
// (OpenCV's MatExpr semantics - computeDescriptors at src/ORBextractor.cc:1037 relies on it to
// write through the rowRange view of the output descriptor matrix).
struct MatExpr { int rows, cols; };

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    int mtype = CV_8U;  // CV_8U or CV_32F (the latter only so that DBoW2's FORB::toMat32F compiles)
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() {}
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type) { create(r, c, type); }
    // external data, not owned
    Mat(int r, int c, int /*type*/, void* ext, size_t stp) : rows(r), cols(c), step(stp), data((uchar*)ext) {}

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == mtype) return;
        const size_t es = type == CV_32F ? 4 : 1;
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c * es);
        rows = r; cols = c; step = (size_t)c * es; data = buf->data(); mtype = type;
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    static MatExpr zeros(int r, int c, int /*type*/) { return MatExpr{r, c}; }
    Mat(const MatExpr& e) { *this = e; }
    Mat& operator=(const MatExpr& e) {
        create(e.rows, e.cols, 0);
        for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step, 0, cols);
        return *this;
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return mtype; }
    size_t step1() const { return step; }
    Mat clone() const {
        Mat m(rows, cols, 0);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    Mat view(int x, int y, int w, int h) const {
        Mat m; m.rows = h; m.cols = w; m.step = step; m.data = data + (size_t)y * step + x; m.buf = buf; return m;
    }
    Mat operator()(const Rect& r) const { return view(r.x, r.y, r.width, r.height); }
    Mat rowRange(int a, int b) const { return view(0, a, cols, b - a); }
    Mat colRange(int a, int b) const { return view(a, 0, b - a, rows); }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + c * sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    Mat row(int r) const { return view(0, r, cols, 1); }
};

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
