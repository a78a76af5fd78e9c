// Minimal cv:: types for building the facade where OpenCV's C++ headers are not installed. With real
// OpenCV present (the normal case for an ORB-SLAM2 build) this header is not used: ORBextractor.h /
// ORBmatcher.h include <opencv2/core/core.hpp> instead. Only what the two facade classes touch.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#endif

namespace cv {

typedef unsigned char uchar;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t stp) : rows(r), cols(c), step(stp), data((uchar*)ext), type_(type) {}
    void create(int r, int c, int type) {  // CV_8U or CV_32F, one channel
        if (data && r == rows && c == cols && type == type_) return;
        const size_t esz = type == CV_32F ? 4 : 1;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * esz);
        rows = r; cols = c; step = (size_t)c * esz; data = buf_->data(); type_ = type;
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return type_; }
    Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.buf_ = buf_; m.type_ = type_; return m; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + c * sizeof(T)); }
private:
    std::shared_ptr<std::vector<uchar>> buf_;
    int type_ = CV_8UC1;
};

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    void release() const { m_->release(); }
    Mat getMat() const { return *m_; }
private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
