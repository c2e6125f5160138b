/* oracle/acmmp_oracle.c
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C, CPU restatement of the reference's PatchMatch path
 * (reference = /root/reference/ACMMP.cu, a CUDA-only implementation; there is no CPU PatchMatch
 * upstream).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, load or
 * call this file; the product library never does.
 *
 * Parity pin: the reference ships no tests or golden vectors.  This restatement
 * is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref/libacmmp_ref.so (the
 * unmodified reference sources compiled for sm_100) is run on the GPU box on seeded synthetic scenes
 * and its outputs are committed under tests/golden/ (generator: tests/golden/make_golden.py);
 * tests/test_cpu_oracle.py checks this file against those vectors.
 *
 * Differences from the device arithmetic that cannot be removed on a CPU: the reference is built with
 * --use_fast_math (approximate rcp / rsqrt / sin / cos / ex2, FTZ) and reads images through the
 * texture unit.  Here libm is used and the bilinear filter is emulated with the documented 1.8
 * fixed-point fractions; agreement is to ~1e-4, not bit-exact.
 *
 * Each function cites the reference lines (file:line) it follows.
 */
#include "acmmp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------
 * row-parallel driver (pthreads; the image has no libgomp).  ORC_THREADS overrides the thread count.
 * ---------------------------------------------------------------------------------------------- */
typedef void (*orc_row_fn)(int y, void *arg);
typedef struct { int H; int next; orc_row_fn fn; void *arg; } orc_par;

static void *orc_par_worker(void *p)
{
    orc_par *j = (orc_par *)p;
    for (;;) {
        const int y = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (y >= j->H) break;
        j->fn(y, j->arg);
    }
    return NULL;
}

int orc_num_threads(void)
{
    const char *e = getenv("ORC_THREADS");
    int n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return n;
}

static void par_rows(int H, orc_row_fn fn, void *arg)
{
    orc_par job = {H, 0, fn, arg};
    const int n = orc_num_threads();
    if (n == 1) { orc_par_worker(&job); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < n - 1; ++i)
        if (pthread_create(&th[started], NULL, orc_par_worker, &job) == 0) started++;
    orc_par_worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

#define ORC_PINHOLE 0
#define ORC_SPHERE 11
#define ORC_PI_F 3.141592654f /* CUDART_PI_F */

/* ------------------------------------------------------------------------------------------------
 * texture fetch: float32, cudaFilterModeLinear, un-normalised coordinates -> clamp addressing
 * (ACMMP.cpp:698-704; CUDA Programming Guide "Linear Filtering": xB = x - 0.5, fractions kept in
 * 9-bit fixed point with 8 fractional bits)
 * ---------------------------------------------------------------------------------------------- */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

float orc_tex2d(const orc_image *img, float x, float y)
{
    const float xb = x - 0.5f, yb = y - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    const float a = floorf((xb - fx) * 256.0f + 0.5f) / 256.0f;
    const float b = floorf((yb - fy) * 256.0f + 0.5f) / 256.0f;
    int i = (int)fx, j = (int)fy;
    if (!(xb == xb)) i = 0;
    if (!(yb == yb)) j = 0;
    const int i0 = clampi(i, 0, img->width - 1), i1 = clampi(i + 1, 0, img->width - 1);
    const int j0 = clampi(j, 0, img->height - 1), j1 = clampi(j + 1, 0, img->height - 1);
    const float t00 = img->data[(size_t)j0 * img->width + i0], t10 = img->data[(size_t)j0 * img->width + i1];
    const float t01 = img->data[(size_t)j1 * img->width + i0], t11 = img->data[(size_t)j1 * img->width + i1];
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}

/* ------------------------------------------------------------------------------------------------
 * camera model
 * ---------------------------------------------------------------------------------------------- */
/* PixelToDir, ACMMP.cu:119-134 */
void orc_pixel_to_dir(const orc_camera *cam, int x, int y, float dir[3])
{
    if (cam->model == ORC_PINHOLE) {
        dir[0] = ((float)x - cam->K[2]) / cam->K[0];
        dir[1] = ((float)y - cam->K[5]) / cam->K[4];
        dir[2] = 1.f;
        const float inv = 1.0f / sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        dir[0] *= inv; dir[1] *= inv; dir[2] *= inv;
    } else {
        const float lon = ((float)x - cam->params[1]) / (float)cam->width * 2.0f * ORC_PI_F;
        const float lat = -((float)y - cam->params[2]) / (float)cam->height * ORC_PI_F;
        dir[0] = cosf(lat) * sinf(lon);
        dir[1] = -sinf(lat);
        dir[2] = cosf(lat) * cosf(lon);
    }
}

/* ComputeDepthfromPlaneHypothesis, ACMMP.cu:187-193 */
float orc_depth_from_plane(const orc_camera *cam, const float plane[4], int x, int y)
{
    float dir[3];
    orc_pixel_to_dir(cam, x, y, dir);
    const float denom = plane[0] * dir[0] + plane[1] * dir[1] + plane[2] * dir[2];
    return (fabsf(denom) < 1e-6f) ? 1e6f : (-plane[3] / denom);
}

/* GetDistance2Origin, ACMMP.cu:168-173 */
static float distance_to_origin(const orc_camera *cam, int x, int y, float depth, const float normal[3])
{
    float dir[3];
    orc_pixel_to_dir(cam, x, y, dir);
    return -(normal[0] * dir[0] * depth + normal[1] * dir[1] * depth + normal[2] * dir[2] * depth);
}

/* Get3DPointonWorld_cu, ACMMP.cu:565-600 */
void orc_point_on_world(float x, float y, float depth, const orc_camera *cam, float X[3])
{
    float pc[3];
    if (cam->model == ORC_SPHERE) {
        const float lon = (x - cam->params[1]) / (float)cam->width * 2.0f * ORC_PI_F;
        const float lat = -(y - cam->params[2]) / (float)cam->height * ORC_PI_F;
        pc[0] = cosf(lat) * sinf(lon) * depth;
        pc[1] = -sinf(lat) * depth;
        pc[2] = cosf(lat) * cosf(lon) * depth;
    } else {
        pc[0] = depth * (x - cam->K[2]) / cam->K[0];
        pc[1] = depth * (y - cam->K[5]) / cam->K[4];
        pc[2] = depth;
    }
    const float *R = cam->R, *t = cam->t;
    const float tx = R[0] * pc[0] + R[3] * pc[1] + R[6] * pc[2];
    const float ty = R[1] * pc[0] + R[4] * pc[1] + R[7] * pc[2];
    const float tz = R[2] * pc[0] + R[5] * pc[1] + R[8] * pc[2];
    const float Cx = -(R[0] * t[0] + R[3] * t[1] + R[6] * t[2]);
    const float Cy = -(R[1] * t[0] + R[4] * t[1] + R[7] * t[2]);
    const float Cz = -(R[2] * t[0] + R[5] * t[1] + R[8] * t[2]);
    X[0] = tx + Cx; X[1] = ty + Cy; X[2] = tz + Cz;
}

/* ProjectonCamera_cu, ACMMP.cu:602-644 */
void orc_project(const float X[3], const orc_camera *cam, float pt[2], float *depth)
{
    const float *R = cam->R, *t = cam->t;
    const float tx = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const float ty = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const float tz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    if (cam->model == ORC_SPHERE) {
        *depth = sqrtf(tx * tx + ty * ty + tz * tz);
        if (*depth < 1e-6f) {
            pt[0] = cam->params[1];
            pt[1] = cam->params[2];
            return;
        }
        const float latitude = -asinf(ty / *depth);
        const float longitude = atan2f(tx, tz);
        pt[0] = (longitude / (2.0f * ORC_PI_F)) * (float)cam->width + cam->params[1];
        pt[1] = (-latitude / ORC_PI_F) * (float)cam->height + cam->params[2];
    } else {
        *depth = tz;
        pt[0] = (cam->K[0] * tx + cam->K[1] * ty + cam->K[2] * tz) / *depth;
        pt[1] = (cam->K[3] * tx + cam->K[4] * ty + cam->K[5] * tz) / *depth;
    }
}

/* ------------------------------------------------------------------------------------------------
 * costs
 * ---------------------------------------------------------------------------------------------- */
/* ComputeBilateralWeight, ACMMP.cu:398-403 */
static float bilateral_weight(float xd, float yd, float pix, float center_pix, float sigma_spatial, float sigma_color)
{
    const float spatial_dist = sqrtf(xd * xd + yd * yd);
    const float color_dist = fabsf(pix - center_pix);
    return expf(-spatial_dist / (2.0f * sigma_spatial * sigma_spatial) - color_dist / (2.0f * sigma_color * sigma_color));
}

/* ComputeBilateralNCC, ACMMP.cu:405-516 (patch_size 11, radius_increment 2, sigma 5 / 3: ACMMP.h:34-39) */
float orc_bilateral_ncc(const orc_image *ref_img, const orc_camera *ref_cam, const orc_image *src_img,
                        const orc_camera *src_cam, int x, int y, const float plane[4])
{
    const float cost_max = 2.0f;
    const int radius = 11 / 2;
    const float sigma_spatial = 5.0f, sigma_color = 3.0f;

    const float depth_ref = orc_depth_from_plane(ref_cam, plane, x, y);
    float Pw[3], pt[2], dummy;
    orc_point_on_world((float)x, (float)y, depth_ref, ref_cam, Pw);
    orc_project(Pw, src_cam, pt, &dummy);
    if (src_cam->model == ORC_SPHERE) {
        pt[0] = pt[0] - floorf(pt[0] / (float)src_cam->width) * (float)src_cam->width;
        pt[1] = fminf(fmaxf(pt[1], 0.0f), (float)src_cam->height - 1.0f);
    } else {
        if (pt[0] < 0.0f || pt[0] >= src_cam->width || pt[1] < 0.0f || pt[1] >= src_cam->height) return cost_max;
    }
    float scale_x = 1.0f, scale_y = 1.0f, sigma_eff = sigma_spatial;
    if (ref_cam->model == ORC_SPHERE) {
        const float lat_c = -((float)y - ref_cam->params[2]) / (float)ref_cam->height * ORC_PI_F;
        scale_x = (2.0f * ORC_PI_F / (float)ref_cam->width) * cosf(lat_c);
        scale_y = (ORC_PI_F / (float)ref_cam->height);
        sigma_eff = sigma_spatial * (ORC_PI_F / (float)ref_cam->height);
    }
    const float ref_center_pix = orc_tex2d(ref_img, x + 0.5f, y + 0.5f);
    float sum_ref = 0, sum_ref_ref = 0, sum_src = 0, sum_src_src = 0, sum_ref_src = 0, sum_bw = 0;
    for (int i = -radius; i <= radius; i += 2) {
        for (int j = -radius; j <= radius; j += 2) {
            const int rx = x + i, ry = y + j;
            const float ref_pix = orc_tex2d(ref_img, rx + 0.5f, ry + 0.5f);
            const float depth_n = orc_depth_from_plane(ref_cam, plane, rx, ry);
            float Pn[3], sp[2], sd;
            orc_point_on_world((float)rx, (float)ry, depth_n, ref_cam, Pn);
            orc_project(Pn, src_cam, sp, &sd);
            if (src_cam->model == ORC_SPHERE) {
                sp[0] = sp[0] - floorf(sp[0] / (float)src_cam->width) * (float)src_cam->width;
                sp[1] = fminf(fmaxf(sp[1], 0.0f), (float)src_cam->height - 1.0f);
            } else {
                if (sp[0] < 0.0f || sp[0] >= src_cam->width || sp[1] < 0.0f || sp[1] >= src_cam->height) continue;
            }
            const float src_pix = orc_tex2d(src_img, sp[0] + 0.5f, sp[1] + 0.5f);
            const float dx = (ref_cam->model == ORC_SPHERE) ? (i * scale_x) : (float)i;
            const float dy = (ref_cam->model == ORC_SPHERE) ? (j * scale_y) : (float)j;
            const float w = bilateral_weight(dx, dy, ref_pix, ref_center_pix,
                                             (ref_cam->model == ORC_SPHERE) ? sigma_eff : sigma_spatial, sigma_color);
            sum_bw += w;
            sum_ref += w * ref_pix;
            sum_ref_ref += w * ref_pix * ref_pix;
            sum_src += w * src_pix;
            sum_src_src += w * src_pix * src_pix;
            sum_ref_src += w * ref_pix * src_pix;
        }
    }
    if (sum_bw < 1e-6f) return cost_max;
    const float inv_bw = 1.0f / sum_bw;
    const float m_ref = sum_ref * inv_bw, m_src = sum_src * inv_bw;
    const float var_ref = sum_ref_ref * inv_bw - m_ref * m_ref;
    const float var_src = sum_src_src * inv_bw - m_src * m_src;
    if (var_ref < 1e-5f || var_src < 1e-5f) return cost_max;
    const float covar = sum_ref_src * inv_bw - m_ref * m_src;
    float ncc_cost = 1.0f - covar / sqrtf(var_ref * var_src);
    ncc_cost = fmaxf(0.0f, fminf(cost_max, ncc_cost));
    return ncc_cost;
}

/* ComputeGeomConsistencyCost, ACMMP.cu:646-671 */
float orc_geom_cost(const orc_image *depth_img, const orc_camera *ref_cam, const orc_camera *src_cam,
                    const float plane[4], int x, int y)
{
    const float max_cost = 3.0f;
    const float depth = orc_depth_from_plane(ref_cam, plane, x, y);
    float fwd[3], sp[2], sd;
    orc_point_on_world((float)x, (float)y, depth, ref_cam, fwd);
    orc_project(fwd, src_cam, sp, &sd);
    /* (int) of a float: truncation, saturating, NaN -> 0 (cvt.rzi.s32.f32) */
    int ix = (sp[0] == sp[0]) ? (sp[0] >= 2147483648.0f ? 2147483647 : (sp[0] <= -2147483648.0f ? (-2147483647 - 1) : (int)sp[0])) : 0;
    int iy = (sp[1] == sp[1]) ? (sp[1] >= 2147483648.0f ? 2147483647 : (sp[1] <= -2147483648.0f ? (-2147483647 - 1) : (int)sp[1])) : 0;
    ix = clampi(ix, 0, depth_img->width - 1);
    iy = clampi(iy, 0, depth_img->height - 1);
    const float src_depth = depth_img->data[(size_t)iy * depth_img->width + ix];
    if (src_depth == 0.0f) return max_cost;
    float s3[3], bp[2], rd;
    orc_point_on_world(sp[0], sp[1], src_depth, src_cam, s3);
    orc_project(s3, ref_cam, bp, &rd);
    const float dc = x - bp[0], dr = y - bp[1];
    const float e = sqrtf(dc * dc + dr * dr);
    return (e == e) ? fminf(max_cost, e) : max_cost;
}

/* sort_small, ACMMP.cu:36-45 */
static void sort_small(float *d, int n)
{
    for (int i = 1; i < n; i++) {
        const float tmp = d[i];
        int j;
        for (j = i; j >= 1 && tmp < d[j - 1]; j--) d[j] = d[j - 1];
        d[j] = tmp;
    }
}

/* ComputeMultiViewInitialCostandSelectedViews, ACMMP.cu:519-556 */
float orc_init_cost(int n_images, const orc_image *imgs, const orc_camera *cams, int x, int y, const float plane[4],
                    uint32_t *selected_views)
{
    const float cost_max = 2.0f;
    float cv[32], cvc[32];
    int count = 0, valid = 0;
    for (int i = 1; i < n_images; ++i) {
        const float c = orc_bilateral_ncc(&imgs[0], &cams[0], &imgs[i], &cams[i], x, y, plane);
        cv[i - 1] = c; cvc[i - 1] = c; count++;
        if (c < cost_max) valid++;
    }
    sort_small(cv, count);
    *selected_views = 0;
    const int top_k = valid < 4 ? valid : 4;
    if (top_k > 0) {
        float cost = 0.0f;
        for (int i = 0; i < top_k; ++i) cost += cv[i];
        const float thr = cv[top_k - 1];
        for (int i = 0; i < n_images - 1; ++i)
            if (cvc[i] <= thr) *selected_views |= (1u << i);
        return cost / top_k;
    }
    return cost_max;
}

/* ------------------------------------------------------------------------------------------------
 * whole-map drivers
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const orc_image *a_img, *b_img; const orc_camera *a_cam, *b_cam; const float *planes4; float *out; float *out4;
    int n_images; const orc_image *imgs; const orc_camera *cams; uint32_t *views;
} map_args;

static void ncc_row(int y, void *p)
{
    map_args *a = (map_args *)p;
    const int W = a->a_cam->width;
    for (int x = 0; x < W; ++x)
        a->out[(size_t)y * W + x] = orc_bilateral_ncc(a->a_img, a->a_cam, a->b_img, a->b_cam, x, y, a->planes4 + 4 * ((size_t)y * W + x));
}

void orc_ncc_map(const orc_image *ref_img, const orc_camera *ref_cam, const orc_image *src_img, const orc_camera *src_cam,
                 const float *planes4, float *out)
{
    map_args a = {ref_img, src_img, ref_cam, src_cam, planes4, out, NULL, 0, NULL, NULL, NULL};
    par_rows(ref_cam->height, ncc_row, &a);
}

static void geom_row(int y, void *p)
{
    map_args *a = (map_args *)p;
    const int W = a->a_cam->width;
    for (int x = 0; x < W; ++x)
        a->out[(size_t)y * W + x] = orc_geom_cost(a->a_img, a->a_cam, a->b_cam, a->planes4 + 4 * ((size_t)y * W + x), x, y);
}

void orc_geom_map(const orc_image *depth_img, const orc_camera *ref_cam, const orc_camera *src_cam, const float *planes4,
                  float *out)
{
    map_args a = {depth_img, NULL, ref_cam, src_cam, planes4, out, NULL, 0, NULL, NULL, NULL};
    par_rows(ref_cam->height, geom_row, &a);
}

void orc_warp_map(const orc_camera *ref_cam, const orc_camera *src_cam, const float *planes4, float *out4)
{
    const int W = ref_cam->width, H = ref_cam->height;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float *pl = planes4 + 4 * ((size_t)y * W + x);
            const float depth = orc_depth_from_plane(ref_cam, pl, x, y);
            float X[3], pt[2], d;
            orc_point_on_world((float)x, (float)y, depth, ref_cam, X);
            orc_project(X, src_cam, pt, &d);
            float *o = out4 + 4 * ((size_t)y * W + x);
            o[0] = pt[0]; o[1] = pt[1]; o[2] = d; o[3] = depth;
        }
}

static void initcost_row(int y, void *p)
{
    map_args *a = (map_args *)p;
    const int W = a->cams[0].width;
    for (int x = 0; x < W; ++x)
        a->out[(size_t)y * W + x] = orc_init_cost(a->n_images, a->imgs, a->cams, x, y, a->planes4 + 4 * ((size_t)y * W + x), &a->views[(size_t)y * W + x]);
}

void orc_initcost_map(int n_images, const orc_image *imgs, const orc_camera *cams, const float *planes4, float *out,
                      uint32_t *views)
{
    map_args a = {NULL, NULL, NULL, NULL, planes4, out, NULL, n_images, imgs, cams, views};
    par_rows(cams[0].height, initcost_row, &a);
}

/* SpatialGauss / RangeGauss, ACMMP.cu:175-185 */
static float spatial_gauss(float x1, float y1, float x2, float y2, float sigma)
{
    const float dis = (float)(pow((double)(x1 - x2), 2) + pow((double)(y1 - y2), 2) - (double)0.0f);
    return (float)exp(-1.0 * dis / (2 * sigma * sigma));
}
static float range_gauss(float x, float sigma)
{
    const float x_p = x - 0.0f;
    return (float)exp(-1.0 * (x_p * x_p) / (2 * sigma * sigma));
}

/* JBU_cu, ACMMP.cu:1558-1616 with Imagescale = max(rows / s_rows, cols / s_cols) (ACMMP.cpp:1075) */
void orc_jbu(const float *image, int cols, int rows, const float *depth, int s_width, int s_height, float *out)
{
    const int a = rows / s_height, b = cols / s_width;
    const int Imagescale = a > b ? a : b;
    const float scale = (float)(1.0 * s_width / cols);
    const float sigmad = 0.50f, sigmar = 25.5f;
    const int num_neighbors = (Imagescale * Imagescale + 1) / 2;
    for (int py = 0; py < rows; ++py)
        for (int px = 0; px < cols; ++px) {
            const float o_y = py * scale, o_x = px * scale;
            const float refPix = image[(size_t)py * cols + px];
            float total = 0.0f, norm = 0.0f;
            for (int j = -num_neighbors; j <= num_neighbors; ++j) {
                int r_y = (int)(o_y + j);
                r_y = (r_y > 0 ? (r_y < s_height ? r_y : s_height - 1) : 0);
                int r_ys = py + j;
                r_ys = (r_ys > 0 ? (r_ys < rows ? r_ys : rows - 1) : 0);
                for (int i = -num_neighbors; i <= num_neighbors; ++i) {
                    int r_x = (int)(o_x + i);
                    r_x = (r_x > 0 ? (r_x < s_width ? r_x : s_width - 1) : 0);
                    const float srcPix = depth[(size_t)r_y * s_width + r_x];
                    int r_xs = px + i;
                    r_xs = (r_xs > 0 ? (r_xs < cols ? r_xs : cols - 1) : 0);
                    const float nb = image[(size_t)r_ys * cols + r_xs];
                    const float tg = spatial_gauss(o_x, o_y, (float)r_x, (float)r_y, sigmad) * range_gauss(fabsf(refPix - nb), sigmar);
                    norm += tg;
                    total += srcPix * tg;
                }
            }
            out[(size_t)py * cols + px] = total / norm;
        }
}

/* GetDepthandNormal, ACMMP.cu:1351-1364 */
void orc_depth_normal(const orc_camera *cam, float *planes4)
{
    const int W = cam->width, H = cam->height;
    const float *R = cam->R;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float *p = planes4 + 4 * ((size_t)y * W + x);
            const float d = orc_depth_from_plane(cam, p, x, y);
            const float nx = R[0] * p[0] + R[3] * p[1] + R[6] * p[2];
            const float ny = R[1] * p[0] + R[4] * p[1] + R[7] * p[2];
            const float nz = R[2] * p[0] + R[5] * p[1] + R[8] * p[2];
            p[0] = nx; p[1] = ny; p[2] = nz; p[3] = d;
        }
}

/* CheckerboardFilter, ACMMP.cu:1366-1480; colour 0 = black ((x+y) even), 1 = red */
void orc_median_filter(int W, int H, float *planes4, const float *costs, int colour)
{
#define PW_(idx) planes4[4 * (size_t)(idx) + 3]
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (((x + y) & 1) != colour) continue;
            const int c = y * W + x;
            float f[21];
            int n = 0;
            f[n++] = PW_(c);
            if (costs[c] < 0.001f) continue;
            const int left = c - 1, ll = c - 3, up = c - W, uu = c - 3 * W, down = c + W, dd = c + 3 * W, right = c + 1, rr = c + 3;
            if (y > 0) f[n++] = PW_(up);
            if (y > 2) f[n++] = PW_(uu);
            if (y > 4) f[n++] = PW_(uu - W * 2);
            if (y < H - 1) f[n++] = PW_(down);
            if (y < H - 3) f[n++] = PW_(dd);
            if (y < H - 5) f[n++] = PW_(dd + W * 2);
            if (x > 0) f[n++] = PW_(left);
            if (x > 2) f[n++] = PW_(ll);
            if (x > 4) f[n++] = PW_(ll - 2);
            if (x < W - 1) f[n++] = PW_(right);
            if (x < W - 3) f[n++] = PW_(rr);
            if (x < W - 5) f[n++] = PW_(rr + 2);
            if (y > 0 && x < W - 2) f[n++] = PW_(up + 2);
            if (y < H - 1 && x < W - 2) f[n++] = PW_(down + 2);
            if (y > 0 && x > 1) f[n++] = PW_(up - 2);
            if (y < H - 1 && x > 1) f[n++] = PW_(down - 2);
            if (x > 0 && y > 2) f[n++] = PW_(left - W * 2);
            if (x < W - 1 && y > 2) f[n++] = PW_(right - W * 2);
            if (x > 0 && y < H - 2) f[n++] = PW_(left + W * 2);
            if (x < W - 1 && y < H - 2) f[n++] = PW_(right + W * 2);
            sort_small(f, n);
            const int m = n / 2;
            PW_(c) = (n % 2 == 0) ? (f[m - 1] + f[m]) / 2 : f[m];
        }
#undef PW_
}

/* ------------------------------------------------------------------------------------------------
 * cuRAND XORWOW (CUDA 12.9 curand_kernel.h: curand_init -> _skipahead_sequence / _skipahead, curand(),
 * _curand_uniform).  The sequence skip is 2^67 steps per subsequence; its matrix is rebuilt here by
 * squaring the one-step matrix instead of reading curand_precalc.h.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t w[5]; } v160;
static v160 g_T[160];
static int g_T_ready = 0;

static v160 xw_step(v160 s)
{
    const uint32_t t = s.w[0] ^ (s.w[0] >> 2);
    v160 o;
    o.w[0] = s.w[1]; o.w[1] = s.w[2]; o.w[2] = s.w[3]; o.w[3] = s.w[4];
    o.w[4] = (s.w[4] ^ (s.w[4] << 4)) ^ (t ^ (t << 1));
    return o;
}

static v160 mat_apply(const v160 *m, v160 v)
{
    v160 r = {{0, 0, 0, 0, 0}};
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 32; ++j)
            if (v.w[i] & (1u << j))
                for (int k = 0; k < 5; ++k) r.w[k] ^= m[i * 32 + j].w[k];
    return r;
}

static void build_T(void)
{
    static pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_lock(&mtx);
    {
        if (!g_T_ready) {
            v160 m[160], sq[160];
            for (int r = 0; r < 160; ++r) {
                v160 e = {{0, 0, 0, 0, 0}};
                e.w[r / 32] = 1u << (r % 32);
                m[r] = xw_step(e);
            }
            for (int s = 0; s < 67; ++s) {
                for (int r = 0; r < 160; ++r) sq[r] = mat_apply(m, m[r]);
                memcpy(m, sq, sizeof(m));
            }
            memcpy(g_T, m, sizeof(m));
            __atomic_store_n(&g_T_ready, 1, __ATOMIC_RELEASE);
        }
    }
    pthread_mutex_unlock(&mtx);
}

void orc_curand_init(uint64_t seed, uint64_t subsequence, uint64_t offset, uint32_t st[6])
{
    if (!g_T_ready) build_T();
    const uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    uint32_t d = 6615241u + t1 + t0;
    v160 v;
    v.w[0] = 123456789u + t0; v.w[1] = 362436069u ^ t0; v.w[2] = 521288629u + t1;
    v.w[3] = 88675123u ^ t1; v.w[4] = 5783321u + t0;
    for (uint64_t s = 0; s < subsequence; ++s) v = mat_apply(g_T, v);
    for (uint64_t o = 0; o < offset; ++o) v = xw_step(v);
    d += 362437u * (uint32_t)offset;
    st[0] = d;
    for (int k = 0; k < 5; ++k) st[1 + k] = v.w[k];
}

uint32_t orc_curand(uint32_t st[6])
{
    const uint32_t t = st[1] ^ (st[1] >> 2);
    st[1] = st[2]; st[2] = st[3]; st[3] = st[4]; st[4] = st[5];
    st[5] = (st[5] ^ (st[5] << 4)) ^ (t ^ (t << 1));
    st[0] += 362437u;
    return st[5] + st[0];
}

float orc_curand_uniform(uint32_t st[6])
{
    /* nvcc contracts x * 2^-32 + 2^-33 into one FMA */
    return fmaf((float)orc_curand(st), 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

/* ------------------------------------------------------------------------------------------------
 * random hypotheses
 * ---------------------------------------------------------------------------------------------- */
static void normalize3(float v[3])
{
    const float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

/* GenerateRandomNormal, ACMMP.cu:194-220 */
static void random_normal(const orc_camera *cam, int x, int y, uint32_t st[6], float n[3])
{
    float q1 = 1.0f, q2 = 1.0f, s = 2.0f;
    while (s >= 1.0f) {
        q1 = 2.0f * orc_curand_uniform(st) - 1.0f;
        q2 = 2.0f * orc_curand_uniform(st) - 1.0f;
        s = q1 * q1 + q2 * q2;
    }
    const float sq = sqrtf(1.0f - s);
    n[0] = 2.0f * q1 * sq; n[1] = 2.0f * q2 * sq; n[2] = 1.0f - 2.0f * s;
    float vd[3];
    orc_pixel_to_dir(cam, x, y, vd);
    if (n[0] * vd[0] + n[1] * vd[1] + n[2] * vd[2] > 0.0f) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    normalize3(n);
}

/* GeneratePerturbedNormal, ACMMP.cu:222-257 */
static void perturbed_normal(const orc_camera *cam, int x, int y, const float normal[3], uint32_t st[6], float perturbation,
                             float out[3])
{
    float vd[3];
    orc_pixel_to_dir(cam, x, y, vd);
    const float a1 = (orc_curand_uniform(st) - 0.5f) * perturbation;
    const float a2 = (orc_curand_uniform(st) - 0.5f) * perturbation;
    const float a3 = (orc_curand_uniform(st) - 0.5f) * perturbation;
    const float s1 = sinf(a1), s2 = sinf(a2), s3 = sinf(a3), c1 = cosf(a1), c2 = cosf(a2), c3 = cosf(a3);
    float R[9];
    R[0] = c2 * c3; R[1] = c3 * s1 * s2 - c1 * s3; R[2] = s1 * s3 + c1 * c3 * s2;
    R[3] = c2 * s3; R[4] = c1 * c3 + s1 * s2 * s3; R[5] = c1 * s2 * s3 - c3 * s1;
    R[6] = -s2; R[7] = c2 * s1; R[8] = c1 * c2;
    out[0] = R[0] * normal[0] + R[1] * normal[1] + R[2] * normal[2];
    out[1] = R[3] * normal[0] + R[4] * normal[1] + R[5] * normal[2];
    out[2] = R[6] * normal[0] + R[7] * normal[1] + R[8] * normal[2];
    if (out[0] * vd[0] + out[1] * vd[1] + out[2] * vd[2] >= 0.0f) { out[0] = normal[0]; out[1] = normal[1]; out[2] = normal[2]; }
    normalize3(out);
}

/* SampleDepthInv, ACMMP.cu:14-22 */
static float sample_depth_inv(uint32_t st[6], float dmin, float dmax)
{
    dmin = fmaxf(dmin, 1e-6f);
    dmax = fmaxf(dmax, dmin + 1e-6f);
    const float inv_min = 1.0f / dmax, inv_max = 1.0f / dmin;
    const float u = orc_curand_uniform(st);
    const float inv = inv_min + u * (inv_max - inv_min);
    return 1.0f / inv;
}

/* RandomInitialization, branch !geom && !hierarchy (ACMMP.cu:683-689) */
typedef struct {
    int n_images; const orc_image *imgs; const orc_camera *cams; float depth_min, depth_max; uint64_t seed;
    float *planes4, *costs; uint32_t *views, *rand6; int with_costs; uint32_t *row_states;
} rinit_args;

static void rinit_row(int y, void *p)
{
    rinit_args *a = (rinit_args *)p;
    const int W = a->cams[0].width;
    uint32_t row[6];
    memcpy(row, a->row_states + 6 * y, sizeof(row));      /* curand_init(seed, y, 0) */
    for (int x = 0; x < W; ++x) {
        uint32_t st[6];
        memcpy(st, row, sizeof(st));                       /* curand_init(seed, y, x), ACMMP.cu:684 */
        orc_curand(row);
        const size_t c = (size_t)y * W + x;
        /* GenerateRandomPlaneHypothesis, ACMMP.cu:259-265 */
        const float depth = orc_curand_uniform(st) * (a->depth_max - a->depth_min) + a->depth_min;
        float *pl = a->planes4 + 4 * c;
        random_normal(&a->cams[0], x, y, st, pl);
        pl[3] = distance_to_origin(&a->cams[0], x, y, depth, pl);
        if (a->with_costs) a->costs[c] = orc_init_cost(a->n_images, a->imgs, a->cams, x, y, pl, &a->views[c]);
        memcpy(a->rand6 + 6 * c, st, sizeof(st));
    }
}

void orc_random_init(int n_images, const orc_image *imgs, const orc_camera *cams, float depth_min, float depth_max,
                     uint64_t seed, float *planes4, float *costs, uint32_t *views, uint32_t *rand6, int with_costs)
{
    rinit_args a = {n_images, imgs, cams, depth_min, depth_max, seed, planes4, costs, views, rand6, with_costs};
    if (!g_T_ready) build_T();
    /* row-start states: apply the 2^67-step matrix once per row, sequentially */
    const int H = cams[0].height;
    a.row_states = (uint32_t *)malloc(sizeof(uint32_t) * 6 * (size_t)H);
    orc_curand_init(seed, 0, 0, a.row_states);
    for (int y = 1; y < H; ++y) {
        v160 v;
        memcpy(v.w, a.row_states + 6 * (y - 1) + 1, 20);
        v = mat_apply(g_T, v);
        a.row_states[6 * y] = a.row_states[0];
        memcpy(a.row_states + 6 * y + 1, v.w, 20);
    }
    par_rows(H, rinit_row, &a);
    free(a.row_states);
}

/* ------------------------------------------------------------------------------------------------
 * one checkerboard pass
 * ---------------------------------------------------------------------------------------------- */
static int find_min(const float *c, int n) { float m = c[0]; int mi = 0; for (int i = 1; i < n; ++i) if (c[i] <= m) { m = c[i]; mi = i; } return mi; }
static int find_max(const float *c, int n) { float m = c[0]; int mi = 0; for (int i = 1; i < n; ++i) if (c[i] >= m) { m = c[i]; mi = i; } return mi; }
static float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

typedef struct {
    int n_images;
    const orc_image *imgs, *depth_imgs;
    const orc_camera *cams;
    const orc_pass_flags *fl;
} pass_ctx;

/* Race emulation for the planar-prior pass ("the reference's in-pass writes").  In prior mode the
 * reference writes plane_hypotheses[center] in the MIDDLE of a thread's work (ACMMP.cu:1283, :1295) and other threads
 * re-read plane_hypotheses[positions[..]] of same-colour pixels at the same point of THEIR work (:1262, :1279, :1291):
 * a real data race.  The pass proper always reads the pre-pass state (one legal outcome).  With `orc_race_late` set,
 * those re-reads of same-colour positions take the plane from that array instead (= the other extreme: every
 * in-pass write lands before every re-read); `orc_race_center_out` receives each updated pixel's in-pass write. */
static const float *orc_race_late = NULL;
static float *orc_race_center_out = NULL;
void orc_set_race_emulation(const float *late_planes, float *center_planes_out)
{
    orc_race_late = late_planes;
    orc_race_center_out = center_planes_out;
}
static const float *late_plane(const float *planes, int W, int center, int p)
{
    if (orc_race_late && p != center) {
        const int same = (((p % W) + (p / W)) & 1) == (((center % W) + (center / W)) & 1);
        if (same) return orc_race_late + 4 * (size_t)p;
    }
    return planes + 4 * (size_t)p;
}

/* debugging aid of the parity work: ORC_DEBUG_PIXEL="x,y" prints that pixel's decision variables to stderr */
static int orc_dbg_x = -2, orc_dbg_y = -2;
static int orc_dbg(int x, int y)
{
    if (orc_dbg_x == -2) {
        const char *e = getenv("ORC_DEBUG_PIXEL");
        orc_dbg_x = orc_dbg_y = -1;
        if (e) sscanf(e, "%d,%d", &orc_dbg_x, &orc_dbg_y);
    }
    return x == orc_dbg_x && y == orc_dbg_y;
}

static void cost_vector(const pass_ctx *pc, int x, int y, const float *plane, float *cv)
{
    for (int i = 1; i < pc->n_images; ++i)
        cv[i - 1] = orc_bilateral_ncc(&pc->imgs[0], &pc->cams[0], &pc->imgs[i], &pc->cams[i], x, y, plane);
}

/* PlaneHypothesisRefinement, ACMMP.cu:797-936 */
static void refine(const pass_ctx *pc, float *plane, float *depth, float *cost, uint32_t st[6], const float *view_weights,
                   float weight_norm, const float *prior_plane, uint32_t mask, float *restricted_cost, int x, int y)
{
    if (weight_norm <= 0.0f) return;
    const orc_pass_flags *fl = pc->fl;
    const orc_camera *cam0 = &pc->cams[0];
    const float perturbation = 0.02f, gamma = 0.5f, beta = 0.18f;
    const float depth_sigma = (fl->depth_max - fl->depth_min) / 64.0f;
    const float two_ds2 = 2 * depth_sigma * depth_sigma;
    const float angle_sigma = ORC_PI_F * (5.0f / 180.0f);
    const float two_as2 = 2 * angle_sigma * angle_sigma;
    const int use_prior = fl->prior && mask > 0;
    float depth_rand, n_rand[3];
    if (use_prior) {
        const float depth_prior = orc_depth_from_plane(cam0, prior_plane, x, y);
        depth_rand = sample_depth_inv(st, fmaxf(depth_prior - 3 * depth_sigma, fl->depth_min), fminf(depth_prior + 3 * depth_sigma, fl->depth_max));
        perturbed_normal(cam0, x, y, prior_plane, st, angle_sigma, n_rand);
    } else {
        depth_rand = sample_depth_inv(st, fl->depth_min, fl->depth_max);
        random_normal(cam0, x, y, st, n_rand);
    }
    float lo = fmaxf((1.0f - perturbation) * (*depth), fl->depth_min);
    float hi = fminf((1.0f + perturbation) * (*depth), fl->depth_max);
    if (!(hi > lo)) { lo = fl->depth_min; hi = fl->depth_max; }
    float depth_perturbed = *depth;
    int ok = 0;
    for (int k = 0; k < 32; ++k) {
        const float cand = sample_depth_inv(st, lo, hi);
        if (cand >= fl->depth_min && cand <= fl->depth_max) { depth_perturbed = cand; ok = 1; break; }
    }
    if (!ok) depth_perturbed = fminf(fmaxf(*depth, fl->depth_min), fl->depth_max);
    float n_pert[3];
    perturbed_normal(cam0, x, y, plane, st, perturbation * ORC_PI_F, n_pert);

    const float depths[5] = {depth_rand, *depth, depth_rand, *depth, depth_perturbed};
    float normals[5][3];
    memcpy(normals[0], plane, 12); memcpy(normals[1], n_rand, 12); memcpy(normals[2], n_rand, 12);
    memcpy(normals[3], n_pert, 12); memcpy(normals[4], plane, 12);
    for (int i = 0; i < 5; ++i) {
        float cv[32], tp[4];
        memcpy(tp, normals[i], 12);
        tp[3] = distance_to_origin(cam0, x, y, depths[i], tp);
        cost_vector(pc, x, y, tp, cv);
        float temp_cost = 0.0f;
        for (int j = 0; j < pc->n_images - 1; ++j) {
            if (view_weights[j] > 0.0f) {
                if (fl->geom) temp_cost += view_weights[j] * (cv[j] + 0.1f * orc_geom_cost(&pc->depth_imgs[j + 1], cam0, &pc->cams[j + 1], tp, x, y));
                else temp_cost += view_weights[j] * cv[j];
            }
        }
        temp_cost /= weight_norm;
        const float depth_before = orc_depth_from_plane(cam0, tp, x, y);
        if (orc_dbg(x, y)) fprintf(stderr, "  refine cand %d: plane %.7g %.7g %.7g %.7g depth %.7g cost %.7g (cost_now %.7g restricted %.7g)\n", i, tp[0], tp[1], tp[2], tp[3], depth_before, temp_cost, *cost, *restricted_cost);
        if (depth_before < fl->depth_min || depth_before > fl->depth_max || depth_before >= 1e6f) continue;
        if (use_prior) {
            const float depth_prior = orc_depth_from_plane(cam0, prior_plane, x, y);
            const float dd = depths[i] - depth_prior;
            float ac = dot3(prior_plane, tp);
            ac = fminf(fmaxf(ac, -1.0f), 1.0f);
            const float ad = acosf(ac);
            const float prior = gamma + expf(-dd * dd / two_ds2) * expf(-ad * ad / two_as2);
            const float rtc = expf(-temp_cost * temp_cost / beta) * prior;
            if (orc_dbg(x, y)) fprintf(stderr, "     restricted_temp_cost %.7g prior %.7g\n", rtc, prior);
            if (rtc > *restricted_cost) { *depth = depth_before; memcpy(plane, tp, 16); *cost = temp_cost; *restricted_cost = rtc; }
        } else if (temp_cost < *cost) {
            *depth = depth_before; memcpy(plane, tp, 16); *cost = temp_cost;
        }
    }
}

/* CheckerboardPropagation, ACMMP.cu:938-1325, for one pixel; neighbours read from *_in, result to *_out */
static void propagate_pixel(const pass_ctx *pc, int x, int y, int iter, const float *planes, const float *costs,
                            float *planes_out, float *costs_out, const float *pre_costs, uint32_t *selected_views,
                            uint32_t *rand6, const float *prior_planes, const uint32_t *plane_masks)
{
    const orc_pass_flags *fl = pc->fl;
    const orc_camera *cam0 = &pc->cams[0];
    const int W = cam0->width, H = cam0->height, nsrc = pc->n_images - 1;
    const int center = y * W + x;
    int pos[8];
    int flag[8] = {0};
    float cost_array[8][32];
    memset(cost_array, 0, sizeof(cost_array));
    cost_array[0][0] = 2.0f;                     /* `= {2.0f}`, ACMMP.cu:957 */
    /* order of evaluation in the reference: far first, then near; the order does not matter here */
    for (int l = 0; l < 8; ++l) {
        const int vertical = l < 4, du = (l & 2) ? 1 : -1, far_dir = l & 1;
        const int a = vertical ? y : x, A = vertical ? H : W, b = vertical ? x : y, B = vertical ? W : H;
        const int sa = vertical ? W : 1, sb = vertical ? 1 : W;
#define INB(s) (du < 0 ? (a - (s) >= 0) : (a + (s) <= A - 1))
        pos[l] = center;
        if (far_dir) {
            if (!INB(3)) continue;
            flag[l] = 1;
            int p = center + du * 3 * sa;
            float cmin = costs[p];
            for (int i = 1; i < 11; ++i)
                if (INB(3 + 2 * i)) { const int pt = center + du * (3 + 2 * i) * sa; if (costs[pt] < cmin) { cmin = costs[pt]; p = pt; } }
            pos[l] = p;
        } else {
            if (!INB(1)) continue;
            flag[l] = 1;
            int p = center + du * sa;
            float cmin = costs[p];
            for (int i = 0; i < 3; ++i)
                if (INB(2 + i)) {
                    if (b > i) { const int pt = center + du * (2 + i) * sa - i * sb; if (costs[pt] < cmin) { cmin = costs[pt]; p = pt; } }
                    if (b < B - 1 - i) { const int pt = center + du * (2 + i) * sa + i * sb; if (costs[pt] < cmin) { cmin = costs[pt]; p = pt; } }
                }
            pos[l] = p;
        }
#undef INB
        cost_vector(pc, x, y, planes + 4 * (size_t)pos[l], cost_array[l]);
    }
    /* view selection, ACMMP.cu:1146-1208 */
    float view_weights[32] = {0}, priors[32] = {0}, probs[32] = {0};
    const int nbpos[4] = {center - W, center + W, center - 1, center + 1};
    for (int i = 0; i < 4; ++i)
        if (flag[2 * i])
            for (int j = 0; j < nsrc; ++j) priors[j] += ((selected_views[nbpos[i]] >> j) & 1u) ? 0.9f : 0.1f;
    const float thr = (float)(0.8 * expf((iter) * (iter) / (-90.0f)));
    for (int i = 0; i < nsrc; ++i) {
        float count = 0, tmpw = 0;
        int count_false = 0;
        for (int j = 0; j < 8; ++j) {
            if (cost_array[j][i] < thr) { tmpw += expf(cost_array[j][i] * cost_array[j][i] / (-0.18f)); count++; }
            if (cost_array[j][i] > 1.2f) count_false++;
        }
        if (count > 2 && count_false < 3) probs[i] = tmpw / count;
        else if (count_false < 3) probs[i] = expf(thr * thr / (-0.32f));
        probs[i] = probs[i] * priors[i];
    }
    {
        float sum = 0.0f;
        for (int i = 0; i < nsrc; ++i) sum += probs[i];
        const float inv = 1.0f / sum;
        float cum = 0.0f;
        for (int i = 0; i < nsrc; ++i) { cum += probs[i] * inv; probs[i] = cum; }
    }
    uint32_t *st = rand6 + 6 * (size_t)center;
    for (int s = 0; s < 15; ++s) {
        const float r = orc_curand_uniform(st) - FLT_EPSILON;
        for (int id = 0; id < nsrc; ++id)
            if (probs[id] > r) { view_weights[id] += 1.0f; break; }
    }
    uint32_t temp_sel = 0;
    float weight_norm = 0;
    for (int i = 0; i < nsrc; ++i)
        if (view_weights[i] > 0) { temp_sel |= 1u << i; weight_norm += view_weights[i]; }
    float final_costs[8] = {0};
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < nsrc; ++j)
            if (view_weights[j] > 0) {
                if (fl->geom) {
                    if (flag[i]) final_costs[i] += view_weights[j] * (cost_array[i][j] + 0.2f * orc_geom_cost(&pc->depth_imgs[j + 1], cam0, &pc->cams[j + 1], planes + 4 * (size_t)pos[i], x, y));
                    else final_costs[i] += view_weights[j] * (cost_array[i][j] + 0.1f * 3.0f);
                } else final_costs[i] += view_weights[j] * cost_array[i][j];
            }
        final_costs[i] /= weight_norm;
    }
    const int min_idx = find_min(final_costs, 8);
    float cvn[32];
    const float *cur = planes + 4 * (size_t)center;
    cost_vector(pc, x, y, cur, cvn);
    float cost_now = 0.0f;
    for (int i = 0; i < nsrc; ++i) {
        if (fl->geom) cost_now += view_weights[i] * (cvn[i] + 0.2f * orc_geom_cost(&pc->depth_imgs[i + 1], cam0, &pc->cams[i + 1], cur, x, y));
        else cost_now += view_weights[i] * cvn[i];
    }
    cost_now /= weight_norm;
    float depth_now = orc_depth_from_plane(cam0, cur, x, y);
    float plane_center[4], cost_center = cost_now, restricted_cost = 0.0f;
    memcpy(plane_center, cur, 16);
    float plane_now[4], plane_intended[4];
    memcpy(plane_now, cur, 16);
    memcpy(plane_intended, cur, 16);
    int have_now = 0;
    const uint32_t mask = fl->prior ? plane_masks[center] : 0u;
    const float *pp = fl->prior ? prior_planes + 4 * (size_t)center : NULL;
    if (fl->prior) {
        const float gamma = 0.5f, beta = 0.18f;
        const float depth_sigma = (fl->depth_max - fl->depth_min) / 64.0f;
        const float two_ds2 = 2 * depth_sigma * depth_sigma;
        const float angle_sigma = (float)(3.14159265358979323846 * (5.0f / 180.0f));
        const float two_as2 = 2 * angle_sigma * angle_sigma;
        if (mask > 0) {
            const float depth_prior = orc_depth_from_plane(cam0, pp, x, y);
            float rfc[8] = {0};
            for (int i = 0; i < 8; ++i)
                if (flag[i]) {
                    const float *nb = late_plane(planes, W, center, pos[i]);
                    const float dd = orc_depth_from_plane(cam0, nb, x, y) - depth_prior;
                    const float ad = acosf(dot3(pp, nb));
                    const float prior = gamma + expf(-dd * dd / two_ds2) * expf(-ad * ad / two_as2);
                    rfc[i] = expf(-final_costs[i] * final_costs[i] / beta) * prior;
                }
            const int max_idx = find_max(rfc, 8);
            const float dd = depth_now - depth_prior;
            const float ad = acosf(dot3(pp, cur));
            const float prior = gamma + expf(-dd * dd / two_ds2) * expf(-ad * ad / two_as2);
            const float rcn = expf(-cost_now * cost_now / beta) * prior;
            if (flag[max_idx]) {
                const float *nb = late_plane(planes, W, center, pos[max_idx]);
                memcpy(plane_now, nb, 16); have_now = 1;
                const float db = orc_depth_from_plane(cam0, nb, x, y);
                if (db >= fl->depth_min && db <= fl->depth_max && rfc[max_idx] > rcn) {
                    memcpy(plane_center, nb, 16); memcpy(plane_intended, nb, 16);
                    cost_center = final_costs[max_idx];
                    restricted_cost = rfc[max_idx];
                    selected_views[center] = temp_sel;
                }
            }
        } else if (flag[min_idx]) {
            const float *nb = late_plane(planes, W, center, pos[min_idx]);
            memcpy(plane_now, nb, 16); have_now = 1;
            const float db = orc_depth_from_plane(cam0, nb, x, y);
            if (db >= fl->depth_min && db <= fl->depth_max && final_costs[min_idx] < cost_now) {
                depth_now = db;
                memcpy(plane_center, nb, 16); memcpy(plane_intended, nb, 16);
                cost_center = final_costs[min_idx];
            }
        }
    } else if (flag[min_idx]) {
        const float *nb = planes + 4 * (size_t)pos[min_idx];
        memcpy(plane_now, nb, 16); have_now = 1;
        const float db = orc_depth_from_plane(cam0, nb, x, y);
        if (db >= fl->depth_min && db <= fl->depth_max && final_costs[min_idx] < cost_now) {
            depth_now = db;
            cost_now = final_costs[min_idx];
            selected_views[center] = temp_sel;
            memcpy(plane_intended, nb, 16);
        }
    }
    if (orc_race_center_out) memcpy(orc_race_center_out + 4 * (size_t)center, plane_center, 16);
    /* ACMMP.cu:1301: plane_hypotheses_now is uninitialised; see DESIGN.md */
    if (!fl->as_compiled || !have_now) memcpy(plane_now, plane_intended, 16);
    if (orc_dbg(x, y)) {
        fprintf(stderr, "pixel %d,%d mask %u cost_now %.7g depth_now %.7g min_idx %d weight_norm %g sel %x\n", x, y, mask, cost_now, depth_now, min_idx, weight_norm, temp_sel);
        for (int i = 0; i < 8; ++i) {
            const float *nb = planes + 4 * (size_t)pos[i];
            fprintf(stderr, "  dir %d flag %d pos (%d,%d) final %.7g plane %.7g %.7g %.7g %.7g\n", i, flag[i], pos[i] % W, pos[i] / W, final_costs[i], nb[0], nb[1], nb[2], nb[3]);
        }
        fprintf(stderr, "  plane_now %.7g %.7g %.7g %.7g  center %.7g %.7g %.7g %.7g cost_center %.7g restricted %.7g\n", plane_now[0], plane_now[1], plane_now[2], plane_now[3], plane_center[0], plane_center[1], plane_center[2], plane_center[3], cost_center, restricted_cost);
    }
    refine(pc, plane_now, &depth_now, &cost_now, st, view_weights, weight_norm, pp, mask, &restricted_cost, x, y);
    float *po = planes_out + 4 * (size_t)center;
    if (fl->hierarchy && !(cost_now < pre_costs[center] - 0.1f)) {
        memcpy(po, plane_center, 16);
        costs_out[center] = cost_center;
    } else {
        memcpy(po, plane_now, 16);
        costs_out[center] = cost_now;
    }
}

typedef struct {
    const pass_ctx *pc; int colour, iter; const float *planes_in, *costs_in; float *planes_out, *costs_out;
    const float *pre_costs; uint32_t *selected_views, *rand6; const float *prior_planes4; const uint32_t *plane_masks;
} pass_args;

static void pass_row(int y, void *p)
{
    pass_args *a = (pass_args *)p;
    const int W = a->pc->cams[0].width;
    for (int x = 0; x < W; ++x)
        if (((x + y) & 1) == a->colour)
            propagate_pixel(a->pc, x, y, a->iter, a->planes_in, a->costs_in, a->planes_out, a->costs_out, a->pre_costs,
                            a->selected_views, a->rand6, a->prior_planes4, a->plane_masks);
}

void orc_checkerboard_pass(int n_images, const orc_image *imgs, const orc_image *depth_imgs, const orc_camera *cams,
                           const orc_pass_flags *flags, int colour, int iter, const float *planes_in,
                           const float *costs_in, float *planes_out, float *costs_out, const float *pre_costs,
                           uint32_t *selected_views, uint32_t *rand6, const float *prior_planes4,
                           const uint32_t *plane_masks)
{
    const int W = cams[0].width, H = cams[0].height;
    pass_ctx pc = {n_images, imgs, depth_imgs, cams, flags};
    memcpy(planes_out, planes_in, sizeof(float) * 4 * (size_t)W * H);
    memcpy(costs_out, costs_in, sizeof(float) * (size_t)W * H);
    /* selected_views: a pixel reads the masks of its 4-neighbours (other colour) and writes its own */
    pass_args a = {&pc, colour, iter, planes_in, costs_in, planes_out, costs_out, pre_costs, selected_views, rand6, prior_planes4, plane_masks};
    par_rows(H, pass_row, &a);
}

/* ------------------------------------------------------------------------------------------------
 * fusion: SimpleFusionKernel, ACMMP.cu:1664-1814, one reference view on the CPU
 * ---------------------------------------------------------------------------------------------- */
/* tex2D<float4>(image, c, r) of a LINEAR-filtered texture with un-normalised coordinates at an integer coordinate
 * (no half-texel offset): the sample point lies on the corner shared by the texels (c-1 .. c) x (r-1 .. r), both bilinear
 * fractions are exactly 0.5, addressing clamps (ACMMP.cu:1966-1972) => the mean of the four texels.  Texels are
 * value / 255 (convertTo(CV_32FC4, 1 / 255), :1955); the kernel multiplies back by 255 (:1704-1708). */
static float fuse_texel_mean(const float *gray, const unsigned char *bgr, int channel, int w, int h, int c, int r)
{
    const int c0 = c - 1 < 0 ? 0 : (c - 1 > w - 1 ? w - 1 : c - 1), c1 = c < 0 ? 0 : (c > w - 1 ? w - 1 : c);
    const int r0 = r - 1 < 0 ? 0 : (r - 1 > h - 1 ? h - 1 : r - 1), r1 = r < 0 ? 0 : (r > h - 1 ? h - 1 : r);
    const int at[4] = {r0 * w + c0, r0 * w + c1, r1 * w + c0, r1 * w + c1};
    float sum = 0.0f;
    for (int k = 0; k < 4; ++k) {
        const float level = bgr ? (float)bgr[3 * at[k] + channel] : gray[at[k]];
        sum += (float)(level * (1.0 / 255.0));
    }
    return 0.25f * sum * 255.0f;
}

/* depths[i]: w_i * h_i, normals3[i]: w_i * h_i * 3 (world frame), gray[i]: grey levels 0..255, bgr (may be NULL) / bgr[i]
 * (may be NULL): 3 bytes per pixel in OpenCV's B, G, R order; cams[i].width / height = the maps' size.
 * points: w * h * 9 floats (PointList: coord, normal, color -- color in the kernel's order, which is B, G, R: its texels
 * are RGBA after cvtColor(BGR2RGBA) and it sums .z first, :1704-1708), written where flags[idx] = 1. */
void orc_fuse_view(int n_views, const orc_camera *cams, const float *const *depths, const float *const *normals3,
                   const float *const *gray, const unsigned char *const *bgr, int ref, int n_src, const int *src_idx,
                   float *points, int *flags)
{
    const orc_camera *rc = &cams[ref];
    const int width = rc->width, height = rc->height;
    (void)n_views;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < height; ++r)
        for (int c = 0; c < width; ++c) {
            const int idx = r * width + c;
            flags[idx] = 0;
            const float ref_depth = depths[ref][idx];                       /* point-sampled at (c, r): texel (c, r) */
            if (ref_depth <= 0.0f) continue;
            float X[3];
            orc_point_on_world((float)c, (float)r, ref_depth, rc, X);
            const float *rn = normals3[ref] + 3 * idx;
            float psum[3] = {X[0], X[1], X[2]}, nsum[3] = {rn[0], rn[1], rn[2]}, csum[3];
            const unsigned char *rb = bgr ? bgr[ref] : 0;
            for (int k = 0; k < 3; ++k) csum[k] = fuse_texel_mean(gray[ref], rb, k, width, height, c, r);
            int num = 1;
            for (int j = 0; j < n_src && j < 32; ++j) {
                const int s = src_idx[j];
                if (s < 0) continue;
                const orc_camera *sc = &cams[s];
                float pp[2], pd;
                orc_project(X, sc, pp, &pd);
                const int src_c = (int)(pp[0] + 0.5f), src_r = (int)(pp[1] + 0.5f);
                if (src_c < 0 || src_c >= sc->width || src_r < 0 || src_r >= sc->height) continue;
                const int sidx = src_r * sc->width + src_c;
                const float src_depth = depths[s][sidx];
                if (src_depth <= 0.0f) continue;
                float Xs[3], rp[2], dummy;
                orc_point_on_world((float)src_c, (float)src_r, src_depth, sc, Xs);
                orc_project(Xs, rc, rp, &dummy);
                const float reproj_error = hypotf((float)c - rp[0], (float)r - rp[1]);
                const float relative_depth_diff = fabsf(pd - src_depth) / src_depth;
                const float *sn = normals3[s] + 3 * sidx;
                float dot = rn[0] * sn[0] + rn[1] * sn[1] + rn[2] * sn[2];
                dot = fmaxf(-1.0f, fminf(1.0f, dot));
                const float angle = acosf(dot);
                if (reproj_error < 1.0 && relative_depth_diff < 0.01f && angle < 0.149f) {
                    for (int k = 0; k < 3; ++k) { psum[k] += Xs[k]; nsum[k] += sn[k]; }
                    const unsigned char *sb = bgr ? bgr[s] : 0;
                    for (int k = 0; k < 3; ++k) csum[k] += fuse_texel_mean(gray[s], sb, k, sc->width, sc->height, src_c, src_r);
                    num++;
                }
            }
            if (num < 3) continue;
            float *o = points + 9 * (size_t)idx;
            float n[3] = {nsum[0] / num, nsum[1] / num, nsum[2] / num};
            const float len = hypotf(hypotf(n[0], n[1]), n[2]);
            if (len > 0.0f) { n[0] /= len; n[1] /= len; n[2] /= len; }
            for (int k = 0; k < 3; ++k) {
                o[k] = psum[k] / num;
                o[3 + k] = n[k];
                o[6 + k] = csum[k] / num;
            }
            flags[idx] = 1;
        }
}
