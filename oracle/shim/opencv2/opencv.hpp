// oracle/shim/opencv2/opencv.hpp
//
// TEST INFRASTRUCTURE ONLY.  A minimal, self-written stand-in for the handful of
// OpenCV types the reference sources mention (reference main.h:4-9 includes the
// opencv2 headers; there is no OpenCV C++ in this image).  It exists so that the
// UNMODIFIED reference translation units (/root/reference/ACMMP.cu, ACMMP.cpp)
// compile and link into oracle/_ref/libacmmp_ref.so.  It is not OpenCV, carries
// no OpenCV code, and is never part of the product path.
//
// Functional: Mat / Mat_<T> storage (zeros, clone, ptr, step, operator(), at),
// Vec, Point, Rect, Size, Scalar.
// Aborting stubs: imread, imwrite, resize, cvtColor, merge, line, Subdiv2D,
// SVD::solveZ -- the harness never reaches them (images arrive as raw floats).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)

namespace cv {

[[noreturn]] inline void shim_unavailable(const char *what)
{
    std::fprintf(stderr, "oracle cv shim: %s is not available in this image\n", what);
    std::abort();
}

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N >= 3, "Vec"); val[0] = a; val[1] = b; val[2] = c; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 6> Vec6f;
typedef Vec<unsigned char, 3> Vec3b;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
    bool contains(const Point &p) const
    {
        return p.x >= x && p.x < x + width && p.y >= y && p.y < y + height;
    }
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

inline int shim_elem_size(int type)
{
    const int depth = type & 7;
    const int cn = (type >> 3) + 1;
    return (depth == CV_32F ? 4 : 1) * cn;
}

template <typename T> struct shim_type;
template <> struct shim_type<float> { enum { value = CV_32FC1 }; };
template <> struct shim_type<unsigned char> { enum { value = CV_8UC1 }; };
template <> struct shim_type<Vec3f> { enum { value = CV_32FC3 }; };
template <> struct shim_type<Vec3b> { enum { value = CV_8UC3 }; };

class Mat {
public:
    int rows, cols;
    unsigned char *data;
    size_t step[2];

    Mat() : rows(0), cols(0), data(nullptr), type_(CV_8UC1) { step[0] = step[1] = 0; }
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }

    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step[1] = (size_t)shim_elem_size(type);
        step[0] = step[1] * (size_t)c;
        const size_t bytes = step[0] * (size_t)r;
        buf_.reset(new unsigned char[bytes ? bytes : 1], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }

    static Mat zeros(int r, int c, int type)
    {
        Mat m(r, c, type);
        std::memset(m.data, 0, m.step[0] * (size_t)r);
        return m;
    }

    Mat clone() const
    {
        Mat m;
        if (!data) return m;
        m.create(rows, cols, type_);
        std::memcpy(m.data, data, step[0] * (size_t)rows);
        return m;
    }

    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }

    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + step[0] * (size_t)r); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + step[0] * (size_t)r); }
    template <typename T> T &at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T &at(int r, int c) const { return ptr<T>(r)[c]; }

    void convertTo(Mat &dst, int rtype, double alpha = 1.0) const
    {
        const int sdepth = type_ & 7, ddepth = rtype & 7;
        const int cn = channels();
        Mat out(rows, cols, CV_MAKETYPE(ddepth, cn));
        if (ddepth != CV_32F) shim_unavailable("Mat::convertTo to a non-float type");
        for (int r = 0; r < rows; ++r) {
            float *d = out.ptr<float>(r);
            for (int c = 0; c < cols * cn; ++c) {
                const double v = (sdepth == CV_32F) ? (double)ptr<float>(r)[c] : (double)ptr<unsigned char>(r)[c];
                d[c] = (float)(v * alpha);
            }
        }
        dst = out;
    }

protected:
    int type_;
    std::shared_ptr<unsigned char> buf_;
};

template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() { type_ = shim_type<T>::value; }
    Mat_(int r, int c) : Mat(r, c, shim_type<T>::value) {}
    Mat_(const Mat &m) : Mat() { assign(m); }
    Mat_ &operator=(const Mat &m) { assign(m); return *this; }

    T &operator()(int r, int c) { return this->template ptr<T>(r)[c]; }
    const T &operator()(int r, int c) const { return this->template ptr<T>(r)[c]; }
    Mat_ clone() const { return Mat_(Mat::clone()); }

private:
    void assign(const Mat &m)
    {
        if (!m.empty() && shim_elem_size(m.type()) != (int)sizeof(T))
            shim_unavailable("Mat_ conversion between element types");
        Mat::operator=(m);
        type_ = shim_type<T>::value;
    }
};

enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { INTER_LINEAR = 1 };
enum { COLOR_BGR2RGBA = 2, COLOR_RGB2RGBA = 0 };

inline Mat imread(const std::string &, int = IMREAD_COLOR) { shim_unavailable("cv::imread"); }
inline bool imwrite(const std::string &, const Mat &) { shim_unavailable("cv::imwrite"); }
inline void resize(const Mat &, Mat &, Size, double = 0, double = 0, int = INTER_LINEAR) { shim_unavailable("cv::resize"); }
inline void cvtColor(const Mat &, Mat &, int) { shim_unavailable("cv::cvtColor"); }
inline void merge(const std::vector<Mat> &, Mat &) { shim_unavailable("cv::merge"); }
inline void line(Mat &, Point, Point, const Scalar &) { shim_unavailable("cv::line"); }

class Subdiv2D {
public:
    explicit Subdiv2D(Rect) {}
    int insert(Point2f) { shim_unavailable("cv::Subdiv2D"); }
    void getTriangleList(std::vector<Vec6f> &) const { shim_unavailable("cv::Subdiv2D"); }
};

struct SVD {
    static void solveZ(const Mat &, Mat &) { shim_unavailable("cv::SVD::solveZ"); }
};

} // namespace cv
