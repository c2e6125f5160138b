// delaunay.cpp -- Delaunay triangulation of integer pixel positions; stands in for cv::Subdiv2D in the CPU
// planar-prior stage (reference ACMMP.cpp:932-954; OpenCV C++ is not available in this image).
// Incremental Bowyer-Watson with triangle adjacency, point location by walking from the last new triangle
// (the support points arrive in scan order, so walks are a few steps), EXACT predicates in integer arithmetic
// (orientation in int64, in-circle in __int128), co-circular points count as outside.  On co-circular quadruples
// the diagonal may differ from OpenCV's -- both are Delaunay triangulations (on a 3200x2130 view's 272 640 near-grid
// support points the two triangle sets come out identical, tests/test_cpu_cpp_host.py).
#include <cstdint>
#include <vector>

#include "acmmp_host.h"

namespace {

struct P2 {
    int64_t x, y;
};

struct Tri {
    int v[3];
    int n[3];      // n[i]: neighbour across the edge opposite v[i], -1 = none
    bool alive;
};

inline int64_t orient(const P2 &a, const P2 &b, const P2 &c)
{
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}

// > 0: d strictly inside the circumcircle of the counter-clockwise triangle a b c
inline bool in_circle(const P2 &a, const P2 &b, const P2 &c, const P2 &d)
{
    const __int128 ax = a.x - d.x, ay = a.y - d.y, bx = b.x - d.x, by = b.y - d.y, cx = c.x - d.x, cy = c.y - d.y;
    const __int128 a2 = ax * ax + ay * ay, b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
    const __int128 det = ax * (by * c2 - b2 * cy) - ay * (bx * c2 - b2 * cx) + a2 * (bx * cy - by * cx);
    return det > 0;
}

} // namespace

// rect_w / rect_h > 0: the enclosing triangle is the one cv::Subdiv2D::initDelaunay builds for Rect(0, 0, rect_w, rect_h)
// -- (B, 0), (0, B), (-B, -B) with B = kSubdivBig * max(w, h) -- so that the triangulation of the points TOGETHER WITH
// these three virtual vertices, and with it the thin triangles along the convex hull that survive the reference's
// inside-the-image filter (main.cpp:140-146), is the one the reference computes.  The factor is OpenCV's and depends on
// its version: 6 in OpenCV 4.13 (the cv2 of this image, which the parity tests compare with: getVertex(1..3) of a fresh
// Subdiv2D((0, 0, 640, 480)) = (3840, 0), (0, 3840), (-3840, -3840)), 3 in older releases; the reference pins no version
// (README: OpenCV >= 2.4).  Only hull triangles a few pixels thin depend on it.
#ifndef ACMMP_SUBDIV_BIG
#define ACMMP_SUBDIV_BIG 6
#endif
std::vector<int> DelaunayIndices(const std::vector<cv::Point> &points, int rect_w, int rect_h)
{
    const int n = (int)points.size();
    std::vector<int> out;
    if (n < 3) return out;
    int64_t lo_x = points[0].x, hi_x = points[0].x, lo_y = points[0].y, hi_y = points[0].y;
    for (const auto &p : points) {
        lo_x = std::min<int64_t>(lo_x, p.x); hi_x = std::max<int64_t>(hi_x, p.x);
        lo_y = std::min<int64_t>(lo_y, p.y); hi_y = std::max<int64_t>(hi_y, p.y);
    }
    const int64_t S = std::max(hi_x - lo_x, hi_y - lo_y) + 16;
    std::vector<P2> pt(n + 3);
    for (int i = 0; i < n; ++i) pt[i] = P2{points[i].x, points[i].y};
    // enclosing triangle, counter-clockwise, far outside the points
    bool cv_triangle = rect_w > 0 && rect_h > 0;
    if (cv_triangle) {
        const int64_t big = (int64_t)ACMMP_SUBDIV_BIG * (int64_t)std::max(rect_w, rect_h);
        // every point must lie strictly inside it (true for pixels of the rectangle)
        for (const auto &p : points)
            if (p.x < 0 || p.y < 0 || p.x >= rect_w || p.y >= rect_h) cv_triangle = false;
        if (cv_triangle) {
            pt[n + 0] = P2{big, 0};
            pt[n + 1] = P2{0, big};
            pt[n + 2] = P2{-big, -big};
        }
    }
    if (!cv_triangle) {
        pt[n + 0] = P2{lo_x - 20 * S, lo_y - 10 * S};
        pt[n + 1] = P2{hi_x + 20 * S, lo_y - 10 * S};
        pt[n + 2] = P2{(lo_x + hi_x) / 2, hi_y + 30 * S};
    }

    std::vector<Tri> tris;
    tris.reserve((size_t)2 * n + 16);
    tris.push_back(Tri{{n, n + 1, n + 2}, {-1, -1, -1}, true});
    int last = 0;

    std::vector<int> cavity, stack;
    std::vector<int> in_cavity;                  // stamp per triangle
    std::vector<int> tri_of_a(n + 3, -1), tri_of_b(n + 3, -1);
    struct Edge { int a, b, outside; };
    std::vector<Edge> boundary;

    for (int ip = 0; ip < n; ++ip) {
        const P2 p = pt[ip];
        // ---- locate by walking
        int t = last;
        bool duplicate = false;
        for (int guard = 0; guard < (int)tris.size() + 8; ++guard) {
            const Tri &T = tris[t];
            int next = -1;
            for (int i = 0; i < 3; ++i) {
                const P2 &a = pt[T.v[(i + 1) % 3]], &b = pt[T.v[(i + 2) % 3]];
                if (orient(a, b, p) < 0) { next = T.n[i]; break; }
            }
            if (next < 0) break;
            t = next;
        }
        for (int i = 0; i < 3; ++i)
            if (pt[tris[t].v[i]].x == p.x && pt[tris[t].v[i]].y == p.y) duplicate = true;
        if (duplicate) continue;

        // ---- cavity: triangles whose circumcircle contains p, grown from t
        in_cavity.resize(tris.size(), -1);
        cavity.clear();
        stack.clear();
        stack.push_back(t);
        in_cavity[t] = ip;
        while (!stack.empty()) {
            const int c = stack.back();
            stack.pop_back();
            cavity.push_back(c);
            for (int i = 0; i < 3; ++i) {
                const int nb = tris[c].n[i];
                if (nb < 0 || in_cavity[nb] == ip) continue;
                const Tri &N = tris[nb];
                if (in_circle(pt[N.v[0]], pt[N.v[1]], pt[N.v[2]], p)) {
                    in_cavity[nb] = ip;
                    stack.push_back(nb);
                }
            }
        }
        // ---- boundary edges of the cavity, oriented counter-clockwise as seen from inside
        boundary.clear();
        for (const int c : cavity) {
            const Tri &T = tris[c];
            for (int i = 0; i < 3; ++i) {
                const int nb = T.n[i];
                if (nb >= 0 && in_cavity[nb] == ip) continue;
                boundary.push_back(Edge{T.v[(i + 1) % 3], T.v[(i + 2) % 3], nb});
            }
        }
        for (const int c : cavity) tris[c].alive = false;
        // ---- fan of new triangles (a, b, p); recycle the dead slots first
        size_t reuse = 0;
        std::vector<int> fresh;
        fresh.reserve(boundary.size());
        for (const Edge &e : boundary) {
            int id;
            if (reuse < cavity.size()) id = cavity[reuse++];
            else { id = (int)tris.size(); tris.push_back(Tri()); }
            tris[id] = Tri{{e.a, e.b, ip}, {-1, -1, e.outside}, true};
            if (e.outside >= 0) {
                Tri &O = tris[e.outside];
                for (int i = 0; i < 3; ++i) {
                    // the edge of O that is (b, a)
                    if (O.v[(i + 1) % 3] == e.b && O.v[(i + 2) % 3] == e.a) O.n[i] = id;
                }
            }
            tri_of_a[e.a] = id;
            tri_of_b[e.b] = id;
            fresh.push_back(id);
        }
        for (const int id : fresh) {
            Tri &T = tris[id];
            T.n[0] = tri_of_a[T.v[1]];       // across (b, p): the triangle whose a is my b
            T.n[1] = tri_of_b[T.v[0]];       // across (p, a): the triangle whose b is my a
        }
        in_cavity.resize(tris.size(), -1);
        for (const int id : fresh) in_cavity[id] = -1;
        last = fresh.empty() ? last : fresh.back();
    }

    for (const Tri &T : tris) {
        if (!T.alive) continue;
        if (T.v[0] >= n || T.v[1] >= n || T.v[2] >= n) continue;
        out.push_back(T.v[0]);
        out.push_back(T.v[1]);
        out.push_back(T.v[2]);
    }
    return out;
}
