// host_capi.cpp -- C entry points into the host-side helpers, for the CPU tests (ctypes); none touches the GPU.
#include <algorithm>
#include <cstring>
#include "acmmp_host.h"

extern "C" {

int acmmp_host_read_camera(const char *path, acmmp_camera *out)
{
    if (!path || !out) return -1;
    *out = ReadCamera(path);
    return 0;
}

int acmmp_host_write_depth_dmb(const char *path, const float *data, int w, int h)
{
    cv::Mat_<float> m(h, w);
    std::memcpy(m.ptr(), data, sizeof(float) * (size_t)w * h);
    return writeDepthDmb(path, m);
}

int acmmp_host_read_depth_dmb(const char *path, float *data, int cap, int *w, int *h)
{
    cv::Mat_<float> m;
    if (readDepthDmb(path, m)) return -1;
    *w = m.cols; *h = m.rows;
    if ((size_t)cap < m.total()) return -2;
    std::memcpy(data, m.ptr(), sizeof(float) * m.total());
    return 0;
}

int acmmp_host_write_normal_dmb(const char *path, const float *data, int w, int h)
{
    cv::Mat_<cv::Vec3f> m(h, w);
    std::memcpy(static_cast<void *>(m.ptr()), data, sizeof(float) * 3 * (size_t)w * h);
    return writeNormalDmb(path, m);
}

int acmmp_host_read_normal_dmb(const char *path, float *data, int cap, int *w, int *h)
{
    cv::Mat_<cv::Vec3f> m;
    if (readNormalDmb(path, m)) return -1;
    *w = m.cols; *h = m.rows;
    if ((size_t)cap < 3 * m.total()) return -2;
    std::memcpy(data, m.ptr(), sizeof(float) * 3 * m.total());
    return 0;
}

int acmmp_host_resize_linear(const float *src, int w, int h, float *dst, int nw, int nh)
{
    cv::Mat_<float> s(h, w), d;
    std::memcpy(s.ptr(), src, sizeof(float) * (size_t)w * h);
    ResizeLinear(s, d, nw, nh);
    std::memcpy(dst, d.ptr(), sizeof(float) * (size_t)nw * nh);
    return 0;
}

// points: n x (x, y) int32; out: up to cap index triples; returns the number of triangles.
// acmmp_host_delaunay_rect: with the enclosing triangle cv::Subdiv2D builds for Rect(0, 0, w, h) (see delaunay.cpp).
// cv::resize INTER_LINEAR on an 8-bit B, G, R image (tests compare it with cv2.resize)
int acmmp_host_resize_linear_bgr(const unsigned char *src, int w, int h, unsigned char *dst, int nw, int nh)
{
    if (!src || !dst || w <= 0 || h <= 0 || nw <= 0 || nh <= 0) return -1;
    cv::Mat_<cv::Vec3b> a(h, w), b;
    std::memcpy(a.ptr(), src, (size_t)w * h * 3);
    ResizeLinearBgr(a, b, nw, nh);
    std::memcpy(dst, b.ptr(), (size_t)nw * nh * 3);
    return 0;
}

// LoadColourImage: the B, G, R image of a view of a dense folder (w*h*3 bytes)
int acmmp_host_load_colour(const char *dense_folder, int id, unsigned char *dst, int cap, int *w, int *h)
{
    cv::Mat_<cv::Vec3b> img;
    if (!LoadColourImage(dense_folder, id, img)) return -1;
    *w = img.cols; *h = img.rows;
    if ((size_t)cap < img.total() * 3) return -2;
    std::memcpy(dst, img.ptr(), img.total() * 3);
    return 0;
}

int acmmp_host_delaunay_rect(const int32_t *points, int n, int w, int h, int32_t *out, int cap)
{
    std::vector<cv::Point> pts(n);
    for (int i = 0; i < n; ++i) pts[i] = cv::Point(points[2 * i], points[2 * i + 1]);
    const std::vector<int> idx = DelaunayIndices(pts, w, h);
    const int nt = (int)(idx.size() / 3);
    for (int i = 0; i < std::min(nt, cap) * 3; ++i) out[i] = idx[i];
    return nt;
}

int acmmp_host_delaunay(const int32_t *points, int n, int32_t *out, int cap)
{
    std::vector<cv::Point> pts(n);
    for (int i = 0; i < n; ++i) pts[i] = cv::Point(points[2 * i], points[2 * i + 1]);
    const std::vector<int> idx = DelaunayIndices(pts);
    const int nt = (int)(idx.size() / 3);
    for (int i = 0; i < std::min(nt, cap) * 3; ++i) out[i] = idx[i];
    return nt;
}

// The CPU planar-prior stage on host arrays: masks (w*h floats, 1-based triangle ids) and up to cap float4 plane
// parameters; returns the number of triangles (planes).
int acmmp_host_planar_prior(const acmmp_camera *cam, int w, int h, const float *depths, const float *costs, float depth_min,
                            float depth_max, float *masks, float *params4, int cap)
{
    cv::Mat_<float> d(h, w), m;
    std::memcpy(d.ptr(), depths, sizeof(float) * (size_t)w * h);
    std::vector<float4> pp;
    Camera c = *cam;
    c.width = w;
    c.height = h;
    PlanarPriorCpu(c, d, costs, depth_min, depth_max, m, pp);
    std::memcpy(masks, m.ptr(), sizeof(float) * (size_t)w * h);
    for (int i = 0; i < std::min((int)pp.size(), cap); ++i) {
        params4[4 * i] = pp[i].x; params4[4 * i + 1] = pp[i].y; params4[4 * i + 2] = pp[i].z; params4[4 * i + 3] = pp[i].w;
    }
    return (int)pp.size();
}

int acmmp_host_load_grey(const char *dense_folder, int id, float *dst, int cap, int *w, int *h)
{
    cv::Mat_<float> img;
    if (!LoadGreyImage(dense_folder, id, img)) return -1;
    *w = img.cols; *h = img.rows;
    if ((size_t)cap < img.total()) return -2;
    std::memcpy(dst, img.ptr(), sizeof(float) * img.total());
    return 0;
}

// ImageSize: the size of a view's image from the file header (images/%08d.pgm or .jpg)
int acmmp_host_image_size(const char *dense_folder, int id, int *w, int *h)
{
    int cols = 0, rows = 0;
    if (!ImageSize(dense_folder, id, cols, rows)) return -1;
    *w = cols; *h = rows;
    return 0;
}

int acmmp_host_pair_count(const char *dense_folder)
{
    std::vector<Problem> problems;
    GenerateSampleList(dense_folder, problems);
    int total = 0;
    for (const auto &p : problems) total += 1000 + (int)p.src_image_ids.size();
    return total;      // 1000 * views + kept source views
}

// points: n x 9 floats (PointList); writes the PLY the reference's StoreColorPlyFileBinaryPointCloud writes
int acmmp_host_write_ply(const char *path, const float *points9, int n)
{
    std::vector<PointList> pc((size_t)n);
    std::memcpy(static_cast<void *>(pc.data()), points9, sizeof(PointList) * (size_t)n);
    try {
        StoreColorPlyFileBinaryPointCloud(path, pc);
    } catch (...) {
        return -1;
    }
    return 0;
}

} // extern "C"
