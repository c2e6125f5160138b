// cvlite.h -- the handful of OpenCV value types that appear in the reference's ACMMP class surface
// (cv::Mat_<float>, cv::Mat_<cv::Vec3f>, cv::Point, cv::Rect; reference ACMMP.h:57-81, main.h:62-65),
// for images without OpenCV C++ (this one has none).  Define ACMMP_HAVE_OPENCV before including
// acmmp_host.h to use the real ones instead.  Storage is a dense row-major std::vector, copy = deep copy.
#pragma once
#include <cstddef>
#include <vector>

namespace cv {

struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};

struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
    bool contains(const Point &p) const { return p.x >= x && p.x < x + width && p.y >= y && p.y < y + height; }
};

struct Vec3f {
    float v[3] = {0.f, 0.f, 0.f};
    Vec3f() {}
    Vec3f(float a, float b, float c) { v[0] = a; v[1] = b; v[2] = c; }
    float &operator[](int i) { return v[i]; }
    const float &operator[](int i) const { return v[i]; }
};

struct Vec3b {
    unsigned char v[3] = {0, 0, 0};
    Vec3b() {}
    Vec3b(unsigned char a, unsigned char b, unsigned char c) { v[0] = a; v[1] = b; v[2] = c; }
    unsigned char &operator[](int i) { return v[i]; }
    const unsigned char &operator[](int i) const { return v[i]; }
};

template <typename T>
class Mat_ {
public:
    int rows = 0, cols = 0;
    Mat_() {}
    Mat_(int r, int c) : rows(r), cols(c), store_((size_t)r * c) {}
    Mat_(int r, int c, const T &fill) : rows(r), cols(c), store_((size_t)r * c, fill) {}
    static Mat_ zeros(int r, int c) { return Mat_(r, c, T()); }
    bool empty() const { return store_.empty(); }
    Mat_ clone() const { return *this; }
    T &operator()(int r, int c) { return store_[(size_t)r * cols + c]; }
    const T &operator()(int r, int c) const { return store_[(size_t)r * cols + c]; }
    T *ptr() { return store_.data(); }
    const T *ptr() const { return store_.data(); }
    size_t total() const { return store_.size(); }

private:
    std::vector<T> store_;
};

} // namespace cv
