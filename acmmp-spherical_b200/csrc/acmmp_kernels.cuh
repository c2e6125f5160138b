// acmmp_kernels.cuh -- the __global__ entry points of the B200 PatchMatch path.
//
//   k_pass            : one red/black checkerboard pass      (reference Black/RedPixelUpdate,
//                       CheckerboardPropagation + PlaneHypothesisRefinement, ACMMP.cu:797-1349)
//   k_random_init     : RandomInitialization, all four branches            (ACMMP.cu:673-795)
//   k_probe, k_probe_quad, k_probe_coords : sub-kernel probes for parity tests (the device code of the two above:
//                       warp chain / geometric term, the quad-cooperative NCC, its fetch coordinates)
//   k_depth_normal    : GetDepthandNormal                                  (ACMMP.cu:1351-1364)
//   k_median_filter   : Black/RedPixelFilter                               (ACMMP.cu:1366-1504)
//   k_jbu             : JBU_cu                                             (ACMMP.cu:1558-1616)
//   k_support_cells, k_tri_planes, k_tri_raster[_long], k_prior_finish : the planar-prior stage on the device
//                       (reference: CPU, ACMMP.cpp:904-1011 + main.cpp:113-185)
//   k_pad_reference, k_rng_fill, k_export_depth, k_expand_prior, k_make_coarse, ... : data-layout helpers
//
// Work decomposition of k_pass (the kernel that is >93 % of the time): persistent, one 512-thread CTA per SM walking
// 8x16 pixel tiles (64 pixels of the active colour each) in lock step; each pixel is served by a GROUP OF 8 LANES
// (4 pixels per warp).  Lane l scans candidate direction l of the adaptive checkerboard; the 8 neighbour hypotheses
// are evaluated by the group's two quads (4 hypotheses each, the four lanes of a quad splitting the 36 taps into 2x2
// blocks: quad_ncc), argmin by warp shuffles; the current plane and the five refinement hypotheses are evaluated per
// (pixel, selected view) pair, the warp's pairs dealt across its eight quads (WarpPairs).  The reference tile (+5 px
// halo) arrives in shared memory by TMA bulk-tensor copies, double-buffered two tiles ahead; bilateral weights and
// tap rays are computed once per pixel visit; source views are bilinear R32F texture fetches (identical filtering to
// the reference's textures by construction).
#pragma once
#include <cuda.h>
#include "acmmp_device.cuh"

namespace acmmp {

// ------------------------------------------------------------------------------------------
// shared-memory carve-up helpers
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <int MODEL, int TW, int TH, int NPIX, int NT>
struct SmemLayout {
    typedef TileGeom<TW, TH> TG;
    typedef typename AuxType<MODEL>::type AuxT;
    size_t off_tile, off_aux, off_wr, off_rr, off_tq, off_vc, off_ncc, off_bar, off_cost, off_vw, off_probs, off_sums, total;
    size_t tile_stride, aux_stride;     // between the buffers of a multi-buffered tile (k_pass: 2)
    __host__ __device__ SmemLayout(int nsrc, int cost_rows, int group_rows, int tq_entries = kTaps, int nbuf = 1)
    {
        const int nvp = nsrc | 1;
        size_t o = 0;
        tile_stride = align_up(TG::kTileBytes, 128);
        aux_stride = align_up(sizeof(AuxT) * TG::RW * TG::RH, 16);
        off_tile = o; o += tile_stride * nbuf;
        off_aux = o; o += aux_stride * nbuf;
        off_wr = o; o += sizeof(float2) * kTaps * NPIX;
        off_rr = o; o += sizeof(float) * kTaps * NPIX;
        off_tq = o; o += sizeof(float) * (size_t)tq_entries * NT;
        off_vc = o; o += sizeof(ViewConst) * (size_t)nsrc;
        off_ncc = o; o += sizeof(NccConst) * (size_t)nsrc;
        off_bar = o; o += 16;
        off_cost = o; o += sizeof(float) * (size_t)cost_rows * nvp;
        off_vw = o; o += sizeof(float) * (size_t)group_rows * nvp;
        off_probs = o; o += sizeof(float) * (size_t)group_rows * nvp;
        o = align_up(o, 16);
        off_sums = o; o += sizeof(float4) * (size_t)group_rows;
        total = align_up(o, 16);
    }
};

// Stage the per-view constants and the reference tile; build the per-tile-pixel ray table.
// Must be called by every thread of the CTA.
template <int MODEL, int TW, int TH, int NT>
__device__ __forceinline__ void stage_tile(const FrameConst &fc, const NccTable &nt, const CUtensorMap *tmap, const int x0,
                                           const int y0, float *tile_r, typename AuxType<MODEL>::type *aux, ViewConst *s_vc,
                                           NccConst *s_ncc, unsigned long long *bar)
{
    typedef TileGeom<TW, TH> TG;
    const int tid = threadIdx.x;
    if (fc.use_tma) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            tma_load_tile_2d(tile_r, tmap, x0 - kHalo + kRefPad, y0 - kHalo + kRefPad, bar, TG::kTileBytes);
        }
    } else {
        for (int idx = tid; idx < TG::RW * TG::RH; idx += NT) {
            const int cx = idx % TG::RW, cy = idx / TG::RW;
            const int gy = min(y0 - kHalo + kRefPad + cy, fc.H + 2 * kRefPad - 1);
            const int gx = min(x0 - kHalo + kRefPad + cx, fc.ref_pitch - 1);
            tile_r[cy * TG::PW + cx] = __ldg(fc.ref_padded + (size_t)gy * fc.ref_pitch + gx);
        }
    }
    for (int i = tid; i < fc.nsrc * 16; i += NT) s_ncc[i >> 4].a[i & 15] = nt.c[i >> 4].a[i & 15];
    {   // view constants: nsrc * 72 words
        const uint32_t *src = reinterpret_cast<const uint32_t *>(fc.views);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_vc);
        const int nwords = fc.nsrc * (int)(sizeof(ViewConst) / 4);
        for (int i = tid; i < nwords; i += NT) dst[i] = __ldg(src + i);
    }
    for (int idx = tid; idx < TG::RW * TG::RH; idx += NT) {
        const int xx = x0 - kHalo + idx % TG::RW;
        const int yy = y0 - kHalo + idx / TG::RW;
        aux[idx] = make_aux<MODEL>(fc, xx, yy);
    }
    if (fc.use_tma) mbar_wait(bar, 0);
    __syncthreads();
}

template <int MODEL>
__device__ __forceinline__ PixCtx make_pix(const FrameConst &fc, const int x, const int y, const int x0, const int y0)
{
    PixCtx px;
    px.x = x; px.y = y;
    px.tx = x - x0 + kHalo; px.ty = y - y0 + kHalo;
    px.dx = static_cast<float>(x) - fc.cx;
    px.dy = static_cast<float>(y) - fc.cy;
    px.dir = pixel_dir<MODEL>(fc, x, y);
    px.Sw = px.Swr = px.Swrr = 0.f;
    return px;
}

// ComputeMultiViewInitialCostandSelectedViews, ACMMP.cu:519-556, in the quad-cooperative form: the four lanes of a quad
// evaluate their pixel's plane against every source view with quad_ncc (2x2 tap blocks: one texture wavefront per quad
// request; the lane-per-plane form this replaces needed 2.46), the view loop is warp-uniform, the cost row of the
// pixel lives in shared memory and all four lanes then run the (cheap) top-k selection redundantly.
// costrow: nsrc floats of scratch per pixel.  Must be called by all 32 lanes; `want` false = nothing evaluated.
template <int MODEL, int RW, int WRS, int TQS>
__device__ __forceinline__ float init_cost_and_views_quad(const FrameConst &fc, const NccTable &nt, const NccConst *s_ncc,
                                                          const typename AuxType<MODEL>::type *aux, const float2 *wr, const float *rr,
                                                          const PixCtx &px, const float4 &plane, const int q, const bool want,
                                                          float *costrow, uint32_t &selected, float *tq)
{
    constexpr unsigned FULL = 0xffffffffu;
    const float cost_max = 2.0f;
    __syncwarp(FULL);                                   // the previous evaluation's rows and tap depths are free
    const unsigned steps = (MODEL == kModelSphere) ? sphere_tap_steps<WRS>(wr, q, px.Sw, fc.tap_prune) : 0x1FFu;
    quad_fill_depths<MODEL, RW, TQS>(fc, aux, px, plane, q, tq, steps);
    __syncwarp(FULL);
    for (int v = 0; v < fc.nsrc; ++v) {
        const ViewK c = load_view(s_ncc + v);
        FetchView fetch;
        fetch.tex = (cudaTextureObject_t)nt.tex[v];
        quad_ncc<MODEL, 1, RW, WRS, TQS>(
            c, px, aux, wr, rr, tq, fetch, q, want ? 1u : 0u, [](const int) { return 0; },
            [&](const int, const float cst) { costrow[v] = cst; }, steps);
    }
    __syncwarp(FULL);
    selected = 0;
    if (!want) return cost_max;
    int num_valid = 0;
    for (int i = 0; i < fc.nsrc; ++i) num_valid += (costrow[i] < cost_max) ? 1 : 0;
    const int top_k = min(num_valid, 4);      // params.top_k, ACMMP.h:40
    if (top_k <= 0) return cost_max;
    // the top_k smallest costs in ascending order == the head of the reference's sorted vector
    uint32_t taken = 0;
    float cost = 0.0f, threshold = 0.0f;
    for (int k = 0; k < top_k; ++k) {
        float best = 0.f;
        int bi = -1;
        for (int i = 0; i < fc.nsrc; ++i) {
            if ((taken >> i) & 1u) continue;
            const float c = costrow[i];
            if (bi < 0 || c < best) { best = c; bi = i; }
        }
        taken |= 1u << bi;
        cost += best;
        threshold = best;
    }
    for (int i = 0; i < fc.nsrc; ++i) {
        if (costrow[i] <= threshold) selected |= 1u << i;
    }
    return cost / top_k;
}

// ------------------------------------------------------------------------------------------
// probes (parity tests) of the per-pixel geometry: mode 1 geom(view), 2 warp(view); thread per pixel
// ------------------------------------------------------------------------------------------
constexpr int kTpTW = 16, kTpTH = 8;

template <int MODEL>
__global__ void __launch_bounds__(128)
k_probe(const __grid_constant__ FrameConst fc, const int mode, const int view, const float4 *__restrict__ planes,
        float *__restrict__ out, float4 *__restrict__ out4)
{
    const int x = blockIdx.x * kTpTW + (threadIdx.x % kTpTW), y = blockIdx.y * kTpTH + (threadIdx.x / kTpTW);
    if (x >= fc.W || y >= fc.H) return;
    const int center = y * fc.W + x;
    const PixCtx px = make_pix<MODEL>(fc, x, y, 0, 0);
    const float4 plane = planes[center];
    const ViewConst &vc = fc.views[view - 1];
    if (mode == 1) {
        out[center] = geom_cost<MODEL>(fc, vc, px, plane);
    } else {
        const float depth = plane_depth(plane, px.dir);
        float sx, sy, sd;
        forward_project<MODEL>(fc, vc, px, depth, sx, sy, sd);
        out4[center] = make_float4(sx, sy, sd, depth);
    }
}

// ------------------------------------------------------------------------------------------
// probe of the NCC form every kernel here runs: fixed plane per pixel through quad_ncc (four lanes per pixel, one
// hypothesis) with the same tables fill_weights / full_sums / quad_fill_depths build inside k_pass / k_random_init.
//   view >= 1 : ComputeBilateralNCC against that source view                       -> out
//   view == 0 : ComputeMultiViewInitialCostandSelectedViews (all views + top-k)    -> out, out_views
// 16x8 pixel tile, 512 threads.
// ------------------------------------------------------------------------------------------
constexpr int kPqPix = kTpTW * kTpTH, kPqNT = 4 * kPqPix, kPqWRS = kPqPix + 8;

template <int MODEL>
__global__ void __launch_bounds__(kPqNT)
k_probe_quad(const __grid_constant__ FrameConst fc, const __grid_constant__ NccTable nt, const __grid_constant__ CUtensorMap tmap, const int view,
             const float4 *__restrict__ planes, float *__restrict__ out, uint32_t *__restrict__ out_views)
{
    typedef TileGeom<kTpTW, kTpTH> TG;
    typedef typename AuxType<MODEL>::type AuxT;
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout<MODEL, kTpTW, kTpTH, kPqWRS, kPqNT> L(fc.nsrc, kPqPix, 0, kTqPerHyp);
    float *tile_r = reinterpret_cast<float *>(smem + L.off_tile);
    AuxT *aux = reinterpret_cast<AuxT *>(smem + L.off_aux);
    ViewConst *s_vc = reinterpret_cast<ViewConst *>(smem + L.off_vc);
    NccConst *s_ncc = reinterpret_cast<NccConst *>(smem + L.off_ncc);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + L.off_bar);
    float *cost = reinterpret_cast<float *>(smem + L.off_cost);

    const int x0 = blockIdx.x * kTpTW, y0 = blockIdx.y * kTpTH;
    stage_tile<MODEL, kTpTW, kTpTH, kPqNT>(fc, nt, &tmap, x0, y0, tile_r, aux, s_vc, s_ncc, bar);

    const int tid = threadIdx.x;
    const int p = tid >> 2, q = tid & 3;               // pixel slot, lane of its quad
    const int xx = x0 + (p % kTpTW), yy = y0 + (p / kTpTW);
    const bool valid = xx < fc.W && yy < fc.H;
    const int x = min(xx, fc.W - 1), y = min(yy, fc.H - 1);
    const int center = y * fc.W + x;
    PixCtx px = make_pix<MODEL>(fc, x, y, x0, y0);
    float2 *wr = reinterpret_cast<float2 *>(smem + L.off_wr) + p;
    float *rr = reinterpret_cast<float *>(smem + L.off_rr) + p;
    float *tq = reinterpret_cast<float *>(smem + L.off_tq) + tid;
    fill_weights<MODEL, TG::PW, kPqWRS>(fc, tile_r, px, wr, rr, q, 4);
    __syncwarp(0xffffffffu);
    full_sums<kPqWRS>(wr, rr, px);
    if (view == 0) {
        uint32_t sel;
        const float c = init_cost_and_views_quad<MODEL, TG::RW, kPqWRS, kPqNT>(fc, nt, s_ncc, aux, wr, rr, px, planes[center], q, valid,
                                                                               cost + p * (fc.nsrc | 1), sel, tq);
        if (valid && q == 0) {
            out[center] = c;
            out_views[center] = sel;
        }
        return;
    }
    const unsigned steps = (MODEL == kModelSphere) ? sphere_tap_steps<kPqWRS>(wr, q, px.Sw, fc.tap_prune) : 0x1FFu;
    quad_fill_depths<MODEL, TG::RW, kPqNT>(fc, aux, px, planes[center], q, tq, steps);
    __syncwarp(0xffffffffu);
    const ViewK c = load_view(s_ncc + (view - 1));
    FetchView fetch;
    fetch.tex = (cudaTextureObject_t)nt.tex[view - 1];
    quad_ncc<MODEL, 1, TG::RW, kPqWRS, kPqNT>(
        c, px, aux, wr, rr, tq, fetch, q, valid ? 1u : 0u, [](const int) { return 0; }, [&](const int, const float cst) { out[center] = cst; },
        steps);
}

// ------------------------------------------------------------------------------------------
// probe of the sample COORDINATES quad_ncc fetches at: the same device functions in the same composition (folded ray of
// the pixel -> shifted to the quad lane's first tap -> to the block row -> to the block column; PlaneRay depth of the
// tap; tap_coords).  72 floats per pixel: (u, v) of tap k = ii * 6 + jj, texel-centre shift included -- to be compared
// with the reference's unfolded chain (ACMMP.cu:458-476): a parity test uses it to show that the NCC residue above 1e-4
// sits exactly on the pixels where a coordinate crosses a 1/256 boundary of the texture unit's 1.8 fixed-point fraction.
// ------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(128)
k_probe_coords(const __grid_constant__ FrameConst fc, const __grid_constant__ NccTable nt, const int view, const int variant,
               const float4 *__restrict__ planes, float *__restrict__ out)
{
    typedef typename AuxType<MODEL>::type AuxT;
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (x >= fc.W || y >= fc.H) return;
    const int center = y * fc.W + x;
    PixCtx px = make_pix<MODEL>(fc, x, y, 0, 0);
    ViewK c;
#pragma unroll
    for (int i = 0; i < 16; ++i) c.a[i] = nt.c[view - 1].a[i];
    ViewPix<MODEL> vp;
    vp.init(c, px);
    PlaneRay<MODEL> ray;
    ray.init(fc, px, planes[center]);
    for (int q = 0; q < 4; ++q) {
        const int qi = q & 1, qj = q >> 1;
        const ViewPix<MODEL> vq = shift_y(c, shift_x(c, vp, (float)(2 * qi - 5)), (float)(2 * qj - 5));
        for (int by = 0; by < 3; ++by) {
            const ViewPix<MODEL> vrow = shift_y(c, vq, (float)(4 * by));
            for (int bx = 0; bx < 3; ++bx) {
                const ViewPix<MODEL> vt = (bx == 0) ? vrow : shift_x(c, vrow, (float)(4 * bx));
                const int i = 2 * qi - 5 + 4 * bx, j = 2 * qj - 5 + 4 * by;
                const AuxT a = make_aux<MODEL>(fc, x + i, y + j);
                const float t = ray.depth(a, i, j);
                float u, v;
                tap_coords(c, vt, a, t, c.a[11], u, v);
                if (MODEL == kModelSphere && variant != 0) {
                    // the two-hypothesis packed form of the pass (sphere_coords2), this plane as its first / second
                    // hypothesis next to an unrelated one: must give the same bits as the scalar form
                    float2 uv0, uv1;
                    const float4 *dir = reinterpret_cast<const float4 *>(&a);
                    if (variant == 1) sphere_coords2(c, *dir, make_float2(t, 1.01f * t), uv0, uv1);
                    else sphere_coords2(c, *dir, make_float2(0.97f * t, t), uv1, uv0);
                    u = uv0.x; v = uv0.y;
                }
                const int k = (2 * bx + qi) * 6 + (2 * by + qj);
                out[(size_t)center * 72 + 2 * k] = u;
                out[(size_t)center * 72 + 2 * k + 1] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// RandomInitialization, ACMMP.cu:673-795
// ------------------------------------------------------------------------------------------
// SpatialGauss / RangeGauss, ACMMP.cu:175-185 (double precision inside, float in/out)
__device__ __forceinline__ float spatial_gauss(float x1, float y1, float x2, float y2, float sigma)
{
    const double ddx = (double)(x1 - x2), ddy = (double)(y1 - y2);      // pow(., 2) of a float is exact in double
    const float dis = (float)(ddx * ddx + ddy * ddy - (double)0.0f);
    return (float)exp(-1.0 * (double)dis / (double)(2 * sigma * sigma));
}
__device__ __forceinline__ float range_gauss(float x, float sigma)
{
    const float x_p = x - 0.0f;
    return (float)exp(-1.0 * (double)(x_p * x_p) / (double)(2 * sigma * sigma));
}

// Four lanes per pixel (16x8 pixel tile, 512 threads): the per-pixel part (RNG draws, plane construction) is computed by
// all four lanes redundantly -- same inputs, same operations, same bits -- and the cost evaluations run in the
// quad-cooperative NCC form (init_cost_and_views_quad).
#ifndef ACMMP_INIT_MIN_CTAS
#define ACMMP_INIT_MIN_CTAS 2      // 64 registers: two 512-thread CTAs per SM (~92 KB of shared memory each)
#endif
template <int MODEL>
__global__ void __launch_bounds__(kPqNT, ACMMP_INIT_MIN_CTAS)
k_random_init(const __grid_constant__ FrameConst fc, const __grid_constant__ NccTable nt, const __grid_constant__ CUtensorMap tmap)
{
    typedef TileGeom<kTpTW, kTpTH> TG;
    typedef typename AuxType<MODEL>::type AuxT;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout<MODEL, kTpTW, kTpTH, kPqWRS, kPqNT> L(fc.nsrc, kPqPix, 0, kTqPerHyp);
    float *tile_r = reinterpret_cast<float *>(smem + L.off_tile);
    AuxT *aux = reinterpret_cast<AuxT *>(smem + L.off_aux);
    ViewConst *s_vc = reinterpret_cast<ViewConst *>(smem + L.off_vc);
    NccConst *s_ncc = reinterpret_cast<NccConst *>(smem + L.off_ncc);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + L.off_bar);
    float *cost = reinterpret_cast<float *>(smem + L.off_cost);

    const int x0 = blockIdx.x * kTpTW, y0 = blockIdx.y * kTpTH;
    stage_tile<MODEL, kTpTW, kTpTH, kPqNT>(fc, nt, &tmap, x0, y0, tile_r, aux, s_vc, s_ncc, bar);

    // pixels outside the image work on a clamped pixel and store nothing (lock-step view loops)
    const int tid = threadIdx.x;
    const int p = tid >> 2, q = tid & 3;               // pixel slot, lane of its quad
    const int lane = tid & 31, qbase = lane & ~3;
    const int xx = x0 + (p % kTpTW), yy = y0 + (p / kTpTW);
    const bool valid = xx < fc.W && yy < fc.H;
    const int x = min(xx, fc.W - 1), y = min(yy, fc.H - 1);
    const int center = y * fc.W + x;
    PixCtx px = make_pix<MODEL>(fc, x, y, x0, y0);
    float2 *mywr = reinterpret_cast<float2 *>(smem + L.off_wr) + p;
    float *myrr = reinterpret_cast<float *>(smem + L.off_rr) + p;
    float *tq = reinterpret_cast<float *>(smem + L.off_tq) + tid;
    float *costrow = cost + p * (fc.nsrc | 1);
    fill_weights<MODEL, TG::PW, kPqWRS>(fc, tile_r, px, mywr, myrr, q, 4);
    __syncwarp(FULL);
    full_sums<kPqWRS>(mywr, myrr, px);
    auto evaluate = [&](const float4 &pl, uint32_t &sel_out) {
        return init_cost_and_views_quad<MODEL, TG::RW, kPqWRS, kPqNT>(fc, nt, s_ncc, aux, mywr, myrr, px, pl, q, valid, costrow, sel_out, tq);
    };

    Rng rs = rng_load(fc.rng_seeded + 3 * (size_t)center);     // curand_init(seed, y, x), ACMMP.cu:684
    uint32_t sel = 0;
    float4 plane;
    float c;

    if (!fc.geom && !fc.hierarchy) {
        // GenerateRandomPlaneHypothesis, ACMMP.cu:259-265
        const float depth = rng_uniform(rs) * (fc.depth_max - fc.depth_min) + fc.depth_min;
        plane = random_normal(rs, px.dir);
        plane.w = plane_offset(plane, px.dir, depth);
        c = evaluate(plane, sel);
    } else if (fc.prior) {
        if (fc.plane_masks[center] > 0 && fc.costs[center] >= 0.1f) {      // ACMMP.cu:691-703
            const float perturbation = 0.02f;
            const float4 prior = fc.prior_planes[center];
            float depth_perturbed = prior.w;
            const float depth_min_perturbed = (1 - 3 * perturbation) * depth_perturbed;
            const float depth_max_perturbed = (1 + 3 * perturbation) * depth_perturbed;
            depth_perturbed = rng_uniform(rs) * (depth_max_perturbed - depth_min_perturbed) + depth_min_perturbed;
            plane = perturbed_normal(rs, px.dir, prior, (float)(3 * perturbation * 3.14159265358979323846));
            plane.w = depth_perturbed;
        } else {                                                          // ACMMP.cu:704-710
            plane = fc.planes[center];
            const float depth = plane.w;
            plane.w = plane_offset(plane, px.dir, depth);
        }
        c = evaluate(plane, sel);
    } else if (fc.upsample) {
        // joint-bilateral NORMAL upsampling from the coarse level, ACMMP.cu:713-779.  The window's taps (j outer, i inner)
        // are dealt to the quad's four lanes four at a time -- the double-precision Gauss weights are the expensive
        // part -- and every lane then accumulates the four in the reference's order, so the sums keep their bits.
        const float scaled_cols = (float)fc.scaled_cols, scaled_rows = (float)fc.scaled_rows;
        const float scale = (float)(1.0 * scaled_cols / fc.W);
        const float sigmad = 0.50f, sigmar = 25.5f;
        const int Imagescale = (int)fmaxf(fc.W / scaled_cols, fc.H / scaled_rows);
        const int WinWidth = Imagescale * Imagescale + 1;
        const int num_neighbors = WinWidth / 2;           // host guarantees <= kHalo
        const int side = 2 * num_neighbors + 1, ntaps = side * side;
        const float o_y = y * scale, o_x = x * scale;
        const float refPix = tile_r[px.ty * TG::PW + px.tx];
        float normalizing_factor = 0.0f;
        float4 n_total = make_float4(0.f, 0.f, 0.f, 0.f);
        auto tap_of = [&](const int t, int &r_x, int &r_y, int &i, int &j) {
            j = t / side - num_neighbors;
            i = t % side - num_neighbors;
            r_y = (int)(o_y + j);
            r_y = (r_y > 0 ? (r_y < scaled_rows ? r_y : (int)(scaled_rows - 1)) : 0);
            r_x = (int)(o_x + i);
            r_x = (r_x > 0 ? (r_x < scaled_cols ? r_x : (int)(scaled_cols - 1)) : 0);
        };
        for (int t0 = 0; t0 < ntaps; t0 += 4) {
            float mine = 0.f;
            if (t0 + q < ntaps) {
                int r_x, r_y, i, j;
                tap_of(t0 + q, r_x, r_y, i, j);
                const float neighborPix = tile_r[(px.ty + j) * TG::PW + (px.tx + i)];
                const float sgauss = spatial_gauss(o_x, o_y, (float)r_x, (float)r_y, sigmad);
                const float rgauss = range_gauss(fabsf(refPix - neighborPix), sigmar);
                mine = sgauss * rgauss;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float totalgauss = __shfl_sync(FULL, mine, qbase + k);
                if (t0 + k < ntaps) {
                    int r_x, r_y, i, j;
                    tap_of(t0 + k, r_x, r_y, i, j);
                    const int s_center = (int)(r_y * scaled_cols + r_x);
                    float4 srcNorm = fc.coarse_planes[s_center];
                    normalizing_factor += totalgauss;
                    srcNorm.x = srcNorm.x * totalgauss;
                    srcNorm.y = srcNorm.y * totalgauss;
                    srcNorm.z = srcNorm.z * totalgauss;
                    n_total.x = n_total.x + srcNorm.x;
                    n_total.y = n_total.y + srcNorm.y;
                    n_total.z = n_total.z + srcNorm.z;
                }
            }
        }
        n_total.x = n_total.x / normalizing_factor;
        n_total.y = n_total.y / normalizing_factor;
        n_total.z = n_total.z / normalizing_factor;
        normalize3(n_total);
        // cost of the plane exactly as uploaded (normal part defined as 0 here) -> pre_costs, :770-771
        const float4 uploaded = fc.planes[center];
        uint32_t sel0;
        const float c0 = evaluate(uploaded, sel0);
        if (valid && q == 0) fc.pre_costs[center] = c0;
        plane = normal_to_cam(fc, n_total);
        plane.w = plane_offset(plane, px.dir, uploaded.w);
        c = evaluate(plane, sel);
    } else {
        // reload, ACMMP.cu:780-793
        plane = fc.hierarchy ? fc.coarse_planes[center] : fc.planes[center];
        plane = normal_to_cam(fc, plane);
        const float depth = plane.w;
        plane.w = plane_offset(plane, px.dir, depth);
        c = evaluate(plane, sel);
    }
    if (valid && q == 0) {
        fc.planes[center] = plane;
        fc.costs[center] = c;
        fc.selected_views[center] = sel;
        rng_store(fc.rng + 3 * (size_t)center, rs);
    }
}

// ------------------------------------------------------------------------------------------
// checkerboard pass
// ------------------------------------------------------------------------------------------
// Measured on B200 (C2, 6 photometric passes): 8x4 tiles / 128 threads x 4 CTAs 212 ms, 8x6 / 192 x 3 205 ms,
// 8x8 / 256 x 2 176 ms, 8x16 / 512 x 1 168 ms -- larger tiles keep the warps of an SM on one image region
// (texture L1 reuse) and in the same code region (instruction cache); register-capped variants with more
// resident warps (96 / 80 registers) lose to their spills.
#ifndef ACMMP_PASS_MIN_CTAS
#define ACMMP_PASS_MIN_CTAS 1
#endif
#ifndef ACMMP_PASS_TH
#define ACMMP_PASS_TH 16
#endif
constexpr int kPassTW = 8, kPassTH = ACMMP_PASS_TH, kPassPix = kPassTW * kPassTH / 2, kPassNT = 8 * kPassPix;
// Tap stride of the weight tables in elements.  The four lanes of a quad read taps {0, 1, 6, 7} + const of four
// neighbouring pixels in one LDS.64: with a stride = 4 (mod 16) float2 the four taps land in four different bank
// groups (one wavefront per load; the unpadded power-of-two stride cost 8 -- ncu r1f: 42 % of all shared wavefronts).
#ifndef ACMMP_PASS_WPAD
#define ACMMP_PASS_WPAD 4
#endif
constexpr int kPassWRS = kPassPix + ACMMP_PASS_WPAD;
constexpr int kPassTq = 4 * kTqPerHyp;      // tap-depth table entries per lane: phase A holds 4 hypotheses x (9 taps + centre), the
                                             // refinement 3 (five hypotheses spread over the two quads), view sampling 15 draws

// FindMinCostIndex / FindMaxCostIndex, ACMMP.cu:62-86 (ties -> last index)
__device__ __forceinline__ int find_min_idx(const float (&c)[8])
{
    float m = c[0];
    int mi = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        if (c[i] <= m) { m = c[i]; mi = i; }
    }
    return mi;
}
__device__ __forceinline__ int find_max_idx(const float (&c)[8])
{
    float m = c[0];
    int mi = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        if (c[i] >= m) { m = c[i]; mi = i; }
    }
    return mi;
}

__device__ __forceinline__ float4 shfl_plane(const unsigned gmask, const float4 &p, const int src)
{
    float4 r;
    r.x = __shfl_sync(gmask, p.x, src);
    r.y = __shfl_sync(gmask, p.y, src);
    r.z = __shfl_sync(gmask, p.z, src);
    r.w = __shfl_sync(gmask, p.w, src);
    return r;
}

// The (pixel, selected view) pairs of the four pixels of a warp, dealt to its eight quads: pair j = 8 r + (lane >> 2)
// in round r.  With each pixel group working through its own selected views two at a time (one per quad) the warp ran
// max_i ceil(n_i / 2) rounds with idle quads wherever a pixel has fewer views than its busiest neighbour; dealt across
// the warp it runs ceil(sum_i n_i / 8).  A quad that serves another group's pixel takes that pixel's weight tables,
// tap-depth columns, reference-side sums and cost rows from shared memory (all four pixels sit in one tile row).
struct WarpPairs {
    uint32_t m[4];         // selected-view masks of the warp's four pixel groups (0 = group takes no part)
    int pre[5];            // prefix sums of their pair counts
    __device__ __forceinline__ void build(const uint32_t my_mask)
    {
        pre[0] = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            m[g] = __shfl_sync(0xffffffffu, my_mask, 8 * g);
            pre[g + 1] = pre[g] + __popc(m[g]);
        }
    }
    // pair j -> owner group and source view; false when j is past the end
    __device__ __forceinline__ bool get(const int j, int &g, int &vsel) const
    {
        if (j >= pre[4]) return false;
        g = (j >= pre[1]) + (j >= pre[2]) + (j >= pre[3]);
        const uint32_t mg = g == 0 ? m[0] : (g == 1 ? m[1] : (g == 2 ? m[2] : m[3]));
        const int base = g == 0 ? pre[0] : (g == 1 ? pre[1] : (g == 2 ? pre[2] : pre[3]));
        vsel = (int)__fns(mg, 0, j - base + 1);
        return true;
    }
};

// MODE: which stage variant this instance serves -- the reference's flag combinations are exclusive
// (main.cpp:427-473: prior stages have geom_consistency off, geometric stages have planar_prior off), and a
// specialised instance carries only its own code (the instruction cache is a measured bottleneck of this kernel).
constexpr int kModePhoto = 0, kModePrior = 1, kModeGeom = 2;

// Persistent form: the grid is one CTA per SM; CTA b walks the tiles b, b + gridDim.x, ... (tile index = row-major
// over the tile grid) and its 16 warps meet at a __syncthreads() after every tile -- exactly the lock step separate
// CTAs per tile have (the instruction cache needs it: warps left to drift across tiles ran 1.75x slower, 46 % of the
// warp time waiting for instructions).  What the loop saves is everything between two tiles: the CTA launch gap, the
// per-view constants (staged once per CTA), and the latency of the reference tile -- tile and ray table are
// double-buffered, the tile after next is requested (TMA) and its ray table computed right after the barrier.
template <int MODEL, int MODE>
__global__ void __launch_bounds__(kPassNT, ACMMP_PASS_MIN_CTAS)
k_pass(const __grid_constant__ FrameConst fc, const __grid_constant__ NccTable nt, const __grid_constant__ CUtensorMap tmap,
       const int colour, const int iter, const int tiles_x, const int ntiles)
{
    typedef TileGeom<kPassTW, kPassTH> TG;
    typedef typename AuxType<MODEL>::type AuxT;
    constexpr bool kPrior = (MODE == kModePrior), kGeom = (MODE == kModeGeom);
    constexpr int WRS = kPassWRS;
    constexpr int TQS = kPassNT;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout<MODEL, kPassTW, kPassTH, kPassWRS, kPassNT> L(fc.nsrc, kPassNT, kPassPix, kPassTq, 2);
    float2 *wr_all = reinterpret_cast<float2 *>(smem + L.off_wr);
    float *rr_all = reinterpret_cast<float *>(smem + L.off_rr);
    float *tq_all = reinterpret_cast<float *>(smem + L.off_tq);
    ViewConst *s_vc = reinterpret_cast<ViewConst *>(smem + L.off_vc);
    NccConst *s_ncc = reinterpret_cast<NccConst *>(smem + L.off_ncc);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + L.off_bar);        // full[2]
    float *cost_all = reinterpret_cast<float *>(smem + L.off_cost);
    float *vw_all = reinterpret_cast<float *>(smem + L.off_vw);
    float *probs_all = reinterpret_cast<float *>(smem + L.off_probs);
    float4 *sums_all = reinterpret_cast<float4 *>(smem + L.off_sums);      // per pixel slot: (Sw, Swr, Swrr, -)

    const int W = fc.W, H = fc.H;
    const int tid = threadIdx.x;
    const int G = gridDim.x;
    const int g = tid >> 3;            // pixel slot in the CTA
    const int gl = tid & 7;            // lane in the pixel group == candidate direction
    const int lane = tid & 31;
    const int gbase = lane & ~7;       // first lane of the group inside the warp
    const int nsrc = fc.nsrc;
    const int nvp = nsrc | 1;
    float2 *wr = wr_all + g;
    float *rr = rr_all + g;
    float *tq = tq_all + tid;                  // this lane's column of the tap-depth table
    float *costrow = cost_all + (g * 8 + gl) * nvp;
    float *cost_grp = cost_all + (g * 8) * nvp;
    float *vw = vw_all + g * nvp;
    float *probs = probs_all + g * nvp;
    const float4 *planes_in = fc.planes;
    const float *costs_in = fc.costs;

    // buffer `buf` for tile t: ray table by all threads, reference tile by one TMA copy (or plain loads, debug aid)
    auto fill_buffer = [&](const int buf, const int t) {
        AuxT *a = reinterpret_cast<AuxT *>(smem + L.off_aux + buf * L.aux_stride);
        float *dst = reinterpret_cast<float *>(smem + L.off_tile + buf * L.tile_stride);
        const int tx0 = (t % tiles_x) * kPassTW - kHalo, ty0 = (t / tiles_x) * kPassTH - kHalo;
        for (int idx = tid; idx < TG::RW * TG::RH; idx += kPassNT) a[idx] = make_aux<MODEL>(fc, tx0 + idx % TG::RW, ty0 + idx / TG::RW);
        if (fc.use_tma) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tma_load_tile_2d(dst, &tmap, tx0 + kRefPad, ty0 + kRefPad, bars + buf, TG::kTileBytes);
            }
        } else {
            for (int idx = tid; idx < TG::RW * TG::RH; idx += kPassNT) {
                const int cx = idx % TG::RW, cy = idx / TG::RW;
                const int gy = min(ty0 + kRefPad + cy, fc.H + 2 * kRefPad - 1);
                const int gx = min(tx0 + kRefPad + cx, fc.ref_pitch - 1);
                dst[cy * TG::PW + cx] = __ldg(fc.ref_padded + (size_t)gy * fc.ref_pitch + gx);
            }
        }
    };

    // ---- prologue: per-view constants, barriers, the first two tiles -------------------------------------------
    for (int i = tid; i < nsrc * 16; i += kPassNT) s_ncc[i >> 4].a[i & 15] = nt.c[i >> 4].a[i & 15];
    {   // view constants: nsrc * 72 words
        const uint32_t *src = reinterpret_cast<const uint32_t *>(fc.views);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_vc);
        const int nwords = nsrc * (int)(sizeof(ViewConst) / 4);
        for (int i = tid; i < nwords; i += kPassNT) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
    }
    __syncthreads();
    for (int k = 0; k < 2; ++k)
        if ((int)blockIdx.x + k * G < ntiles) fill_buffer(k, blockIdx.x + k * G);
    __syncthreads();

#pragma unroll 1
    for (int it = 0, tile = blockIdx.x; tile < ntiles; ++it, tile += G) {
    const int buf = it & 1;
    const float *tile_r = reinterpret_cast<const float *>(smem + L.off_tile + buf * L.tile_stride);
    const AuxT *aux = reinterpret_cast<const AuxT *>(smem + L.off_aux + buf * L.aux_stride);
    if (fc.use_tma) {
        if (lane == 0) mbar_wait(bars + buf, (unsigned)((it >> 1) & 1));
        __syncwarp(FULL);
    }
    const int x0 = (tile % tiles_x) * kPassTW, y0 = (tile / tiles_x) * kPassTH;

    // pixels of the other colour are carried over to the output buffers unchanged (each warp: its own row)
    if (lane < 4) {
        const int yy = y0 + (tid >> 5);
        const int xx = x0 + 2 * lane + ((yy + colour + 1) & 1);
        if (xx < W && yy < H) {
            const int cc = yy * W + xx;
            fc.planes_alt[cc] = fc.planes[cc];
            fc.costs_alt[cc] = fc.costs[cc];
        }
    }

    const int yy_ = y0 + (g >> 2);
    const int xpar = (yy_ + colour) & 1;               // x offset of the active colour in this tile row
    const int xx_ = x0 + 2 * (g & 3) + xpar;
    // Groups that fall outside the image stay alive (clamped coordinates, nothing stored) so that the
    // warp walks the view loops in lock step; see ncc_views.
    const bool valid = xx_ < W && yy_ < H;
    const int x = min(xx_, W - 1), y = min(yy_, H - 1);
    const int center = y * W + x;
    PixCtx px = make_pix<MODEL>(fc, x, y, x0, y0);

    fill_weights<MODEL, TG::PW, WRS>(fc, tile_r, px, wr, rr, gl, 8);

    // ---- adaptive checkerboard sampling: lane l scans direction l (ACMMP.cu:965-1143) -------
    // 0 up_near 1 up_far 2 down_near 3 down_far 4 left_near 5 left_far 6 right_near 7 right_far
    bool flag = false;
    int pos = center;
    {
        const bool vertical = gl < 4;
        const int du = (gl & 2) ? 1 : -1;
        const bool far_dir = (gl & 1) != 0;
        const int a = vertical ? y : x, A = vertical ? H : W;
        const int b = vertical ? x : y, B = vertical ? W : H;
        const int sa = vertical ? W : 1, sb = vertical ? 1 : W;
#define ACMMP_INB(s) (du < 0 ? (a - (s) >= 0) : (a + (s) <= A - 1))
        // All cost reads of a direction are issued before the first comparison (positions that fall outside read
        // nothing and count as +inf): at the start of a tile the 16 warps of the CTA are all here at once, nobody hides
        // a chain of dependent L2 round trips (the scan took 2.6 % of the warp time as load -> compare -> load).
        constexpr float kInf = 3.0e38f;
        if (far_dir) {
            flag = ACMMP_INB(3);
            if (flag) {
                float cv[11];
                int pt[11];
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const bool ok = ACMMP_INB(3 + 2 * i);
                    pt[i] = center + du * (3 + 2 * i) * sa;
                    cv[i] = ok ? __ldg(costs_in + pt[i]) : kInf;
                }
                pos = pt[0];
                float cmin = cv[0];
#pragma unroll
                for (int i = 1; i < 11; ++i) {
                    if (cv[i] < cmin) { cmin = cv[i]; pos = pt[i]; }
                }
            }
        } else {
            flag = ACMMP_INB(1);
            if (flag) {
                float cv[7];
                int pt[7];
                pt[0] = center + du * sa;
                cv[0] = __ldg(costs_in + pt[0]);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const bool row_ok = ACMMP_INB(2 + i);
                    pt[1 + 2 * i] = center + du * (2 + i) * sa - i * sb;
                    pt[2 + 2 * i] = center + du * (2 + i) * sa + i * sb;
                    cv[1 + 2 * i] = (row_ok && b > i) ? __ldg(costs_in + pt[1 + 2 * i]) : kInf;
                    cv[2 + 2 * i] = (row_ok && b < B - 1 - i) ? __ldg(costs_in + pt[2 + 2 * i]) : kInf;
                }
                pos = pt[0];
                float cmin = cv[0];
#pragma unroll
                for (int i = 1; i < 7; ++i) {
                    if (cv[i] < cmin) { cmin = cv[i]; pos = pt[i]; }
                }
            }
        }
#undef ACMMP_INB
    }
    float4 cand = make_float4(0.f, 0.f, 0.f, 0.f);
    if (flag) cand = planes_in[pos];
    // view bitmask of the immediate neighbour in this direction (near lanes; ACMMP.cu:1149-1160)
    uint32_t nbsel = 0;
    if (flag && !(gl & 1)) {
        const int nb = (gl < 4) ? center + ((gl & 2) ? W : -W) : center + ((gl & 2) ? 1 : -1);
        nbsel = fc.selected_views[nb];
    }
    const unsigned flagbits = (__ballot_sync(FULL, flag) >> gbase) & 0xFFu;

    __syncwarp(FULL);
    full_sums<WRS>(wr, rr, px);
    if (gl == 0) sums_all[g] = make_float4(px.Sw, px.Swr, px.Swrr, 0.f);

    // ---- phase A: 8 neighbour hypotheses x all views (ACMMP.cu:981-1142) ---------------------
    // quad Q of the group evaluates candidates 4Q..4Q+3 (the planes its own four lanes hold), tap-split
    const int Q = gl >> 2, q = gl & 3;
    const int qbase = lane & ~3;
    // SPHERE: the tap steps worth sampling for the four pixels of this warp (all nine unless pruning is on)
    const unsigned steps = (MODEL == kModelSphere) ? sphere_tap_steps<WRS>(wr, q, px.Sw, fc.tap_prune) : 0x1FFu;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const float4 hp = shfl_plane(FULL, cand, qbase + h);
        quad_fill_depths<MODEL, TG::RW, TQS>(fc, aux, px, hp, q, tq + h * kTqPerHyp * TQS, steps);
    }
    {
        const unsigned wantA = valid ? ((flagbits >> (4 * Q)) & 0xFu) : 0u;
        for (int v = 0; v < nsrc; ++v) {
            const ViewK c = load_view(s_ncc + v);
            FetchView fetch;
            fetch.tex = (cudaTextureObject_t)nt.tex[v];
            quad_ncc<MODEL, 4, TG::RW, WRS, TQS>(
                c, px, aux, wr, rr, tq, fetch, q, wantA, [](const int h) { return h * kTqPerHyp * TQS; },
                [&](const int h, const float cst) { cost_grp[(4 * Q + h) * nvp + v] = cst; }, steps);
        }
    }
    if (!flag) {
        // `float cost_array[8][32] = {2.0f}` (ACMMP.cu:957): only element [0][0] is 2, the rest 0
        for (int v = 0; v < nsrc; ++v) costrow[v] = (gl == 0 && v == 0) ? 2.0f : 0.0f;
    }
    __syncwarp(FULL);

    // ---- multi-hypothesis joint view selection (ACMMP.cu:1146-1208) --------------------------
    {
        const float cost_threshold = (float)(0.8 * (double)expf((iter) * (iter) / (-90.0f)));
        uint32_t nsel[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) nsel[n] = __shfl_sync(FULL, nbsel, gbase + 2 * n);
        for (int i = gl; i < nsrc; i += 8) {
            float prior = 0.0f;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                if ((flagbits >> (2 * n)) & 1u) prior += ((nsel[n] >> i) & 1u) ? 0.9f : 0.1f;
            }
            float count = 0;
            int count_false = 0;
            float tmpw = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float cj = cost_grp[j * nvp + i];
                if (cj < cost_threshold) {
                    tmpw += expf(cj * cj / (-0.18f));
                    count++;
                }
                if (cj > 1.2f) count_false++;
            }
            float prob = 0.0f;
            if (count > 2 && count_false < 3) {
                prob = tmpw / count;
            } else if (count_false < 3) {
                prob = expf(cost_threshold * cost_threshold / (-0.32f));
            }
            probs[i] = prob * prior;
        }
    }
    __syncwarp(FULL);

    Rng rs;
    uint32_t temp_selected_views = 0;
    float weight_norm = 0;
    // Lane 0 of the group builds the CDF and draws the 15 uniforms (the XORWOW stream is sequential); the 15
    // inverse-CDF searches are then spread over the 8 lanes and the per-view counts come from ballots.
    if (gl == 0) {
        rs = rng_load(fc.rng + 3 * (size_t)center);
        // TransformPDFToCDF, ACMMP.cu:137-151
        float prob_sum = 0.0f;
        for (int i = 0; i < nsrc; ++i) prob_sum += probs[i];
        const float inv_prob_sum = 1.0f / prob_sum;
        float cum_prob = 0.0f;
        for (int i = 0; i < nsrc; ++i) {
            const float prob = probs[i] * inv_prob_sum;
            cum_prob += prob;
            probs[i] = cum_prob;
        }
#pragma unroll 1
        for (int sample = 0; sample < 15; ++sample) tq[sample * TQS] = rng_uniform(rs) - FLT_EPSILON;     // ACMMP.cu:1188
    }
    __syncwarp(FULL);
    {
        const float *draws = tq - gl;              // column of lane 0 of the group
        int id_a = -1, id_b = -1;
        {
            const float ra = draws[gl * TQS];
            for (int image_id = 0; image_id < nsrc; ++image_id) {
                if (probs[image_id] > ra) { id_a = image_id; break; }
            }
            if (gl < 7) {
                const float rb = draws[(gl + 8) * TQS];
                for (int image_id = 0; image_id < nsrc; ++image_id) {
                    if (probs[image_id] > rb) { id_b = image_id; break; }
                }
            }
        }
        // Per-view draw counts of the group without votes (a vote loop here costs the persistent tile loop its
        // warp-uniform texture handles, see the kernel header): every lane adds its one or two draws into 4-bit
        // counters (15 draws at most; 8 views per word), three xor-shuffles sum the words over the group's 8 lanes.
        unsigned cw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (id_a >= 0 && (id_a >> 3) == k) cw[k] += 1u << (4 * (id_a & 7));
            if (id_b >= 0 && (id_b >> 3) == k) cw[k] += 1u << (4 * (id_b & 7));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (8 * k < nsrc) {          // warp-uniform
                cw[k] += __shfl_xor_sync(FULL, cw[k], 1);
                cw[k] += __shfl_xor_sync(FULL, cw[k], 2);
                cw[k] += __shfl_xor_sync(FULL, cw[k], 4);
            }
        }
        unsigned sel_part = 0u, cnt_part = 0u;           // lane gl: views gl, gl + 8, ...
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = gl + 8 * k;
            const unsigned cnt = (cw[k] >> (4 * gl)) & 15u;
            if (i < nsrc) {
                vw[i] = (float)cnt;
                if (cnt > 0u) sel_part |= 1u << i;
                cnt_part += cnt;
            }
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            sel_part |= __shfl_xor_sync(FULL, sel_part, off);
            cnt_part += __shfl_xor_sync(FULL, cnt_part, off);
        }
        temp_selected_views = sel_part;
        weight_norm = (float)cnt_part;                   // small integers: exact in any order
    }
    __syncwarp(FULL);

    // ---- weighted neighbour costs and arg-min over the 8 lanes (ACMMP.cu:1210-1230) ----------
    float final_cost = 0.0f;
    if (kGeom && flag) {
        final_cost = weighted_geom_sum<MODEL>(fc, s_vc, px, cand, vw, costrow, 0.2f, nsrc);
    } else {
        for (int j = 0; j < nsrc; ++j) {
            const float wj = vw[j];
            if (wj > 0) final_cost += kGeom ? wj * (costrow[j] + 0.1f * 3.0f) : wj * costrow[j];
        }
    }
    final_cost /= weight_norm;
    float final_costs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) final_costs[k] = __shfl_sync(FULL, final_cost, gbase + k);
    const int min_cost_idx = find_min_idx(final_costs);
    __syncwarp(FULL);          // the group's cost rows and tap-depth columns are free from here on

    // ---- cost of the current plane (ACMMP.cu:1232-1245) --------------------------------------
    // zero-weight views contribute exactly 0 to the reference's sum, so only the selected views are
    // evaluated: the k-th selected view goes to lane k of the group (one (plane, view) pair per lane).
    const float4 cur_plane = planes_in[center];
    float cost_now = 0.0f;
    {
        // the warp's (pixel, selected view) pairs, one per quad and round (WarpPairs); tap-split inside the quad
        quad_fill_depths<MODEL, TG::RW, TQS>(fc, aux, px, cur_plane, q, tq, steps);
        const int n_sel = valid ? __popc(temp_selected_views) : 0;
        WarpPairs wp;
        wp.build(valid ? temp_selected_views : 0u);
        __syncwarp(FULL);                                  // every group's tap depths and sums are in shared memory
        const int Qw = lane >> 2, g_own = lane >> 3;
        const int rounds = (wp.pre[4] + 7) >> 3;
        float *row_now = cost_grp + 5 * nvp;
        for (int r = 0; r < rounds; ++r) {
            int go = g_own, vsel = 0;
            const bool want = wp.get(8 * r + Qw, go, vsel);
            const int o = (tid >> 5) * 4 + go;             // pixel slot of the pair's owner
            PixCtx po = px;                                // same tile row; x from the owner's (unclamped, valid) position
            po.tx = 2 * go + xpar + kHalo;
            po.dx = static_cast<float>(x0 + 2 * go + xpar) - fc.cx;
            const float4 so = sums_all[o];
            po.Sw = so.x; po.Swr = so.y; po.Swrr = so.z;
            float *row_o = cost_all + (o * 8 + 5) * nvp;
            const ViewK c = load_view(s_ncc + vsel);
            FetchLayer fetch;
            fetch.tex = (cudaTextureObject_t)fc.tex_src;
            fetch.layer = vsel;
            quad_ncc<MODEL, 1, TG::RW, WRS, TQS>(
                c, po, aux, wr_all + o, rr_all + o, tq_all + o * 8 + q, fetch, q, want ? 1u : 0u, [](const int) { return 0; },
                [&](const int, const float cst) { row_o[vsel] = cst; }, steps);
        }
        __syncwarp(FULL);
        // weight (and, in geometric mode, add the geometric term): lane k of the group takes the k-th selected
        // view, so that the group's neighbour-depth loads are in flight together instead of one after the other
        for (int idx = gl; idx < n_sel; idx += 8) {
            const int vsel = (int)__fns(temp_selected_views, 0, idx + 1);
            float cst = row_now[vsel];
            if (kGeom) cst += 0.2f * geom_cost<MODEL>(fc, s_vc[vsel], px, cur_plane);
            row_now[vsel] = vw[vsel] * cst;
        }
        __syncwarp(FULL);
        for (int j = 0; j < nsrc; ++j) {
            if (vw[j] > 0) cost_now += row_now[j];
        }
    }
    cost_now /= weight_norm;
    float depth_now = plane_depth(cur_plane, px.dir);

    // what memory holds for [center] unless the final write-back replaces it
    float4 plane_center = cur_plane;
    float cost_center = cost_now;                        // ACMMP.cu:1244
    uint32_t sel_center = 0;
    bool sel_dirty = false;
    float restricted_cost = 0.0f;
    float4 plane_now = cur_plane;
    bool have_plane_now = false;                         // as-compiled semantics, see below
    float4 plane_intended = cur_plane;                   // "the plane the pixel currently holds"

    const uint32_t mask_c = kPrior ? fc.plane_masks[center] : 0u;
    float4 prior_plane = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kPrior && mask_c > 0) prior_plane = fc.prior_planes[center];
    const float depth_sigma = (fc.depth_max - fc.depth_min) / 64.0f;
    const float two_depth_sigma_squared = 2 * depth_sigma * depth_sigma;
    const float beta = 0.18f;
    const float gamma = 0.5f;
    const float angle_sigma_r = CUDART_PI_F * (5.0f / 180.0f);
    const float two_angle_sigma_squared_r = 2 * angle_sigma_r * angle_sigma_r;

    // All shuffles below run with the full-warp mask outside divergent code: a shuffle under a group mask costs
    // a convergence barrier sequence each.  Values a group does not need are computed and dropped.
    const float4 nb_min = shfl_plane(FULL, cand, gbase + min_cost_idx);
    if (kPrior) {                                        // ACMMP.cu:1247-1299
        const float angle_sigma = (float)(3.14159265358979323846 * (double)(5.0f / 180.0f));
        const float two_angle_sigma_squared = 2 * angle_sigma * angle_sigma;
        const float depth_prior = plane_depth(prior_plane, px.dir);
        float restricted = 0.0f;
        if (flag && mask_c > 0) {
            const float depth_c = plane_depth(cand, px.dir);
            const float depth_diff = depth_c - depth_prior;
            const float angle_cos = dot3(prior_plane, cand);
            const float angle_diff = acosf(angle_cos);
            const float prior = gamma + expf(-depth_diff * depth_diff / two_depth_sigma_squared) *
                                            expf(-angle_diff * angle_diff / two_angle_sigma_squared);
            restricted = expf(-final_cost * final_cost / beta) * prior;
        }
        float restricted_final_costs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) restricted_final_costs[k] = __shfl_sync(FULL, restricted, gbase + k);
        const int max_cost_idx = find_max_idx(restricted_final_costs);
        const float4 nb_max = shfl_plane(FULL, cand, gbase + max_cost_idx);
        if (mask_c > 0) {
            float restricted_cost_now;
            {
                const float depth_diff = depth_now - depth_prior;
                const float angle_cos = dot3(prior_plane, cur_plane);
                const float angle_diff = acosf(angle_cos);
                const float prior = gamma + expf(-depth_diff * depth_diff / two_depth_sigma_squared) *
                                                expf(-angle_diff * angle_diff / two_angle_sigma_squared);
                restricted_cost_now = expf(-cost_now * cost_now / beta) * prior;
            }
            if ((flagbits >> max_cost_idx) & 1u) {
                plane_now = nb_max;
                have_plane_now = true;
                const float depth_before = plane_depth(nb_max, px.dir);
                if (depth_before >= fc.depth_min && depth_before <= fc.depth_max &&
                    restricted_final_costs[max_cost_idx] > restricted_cost_now) {
                    // note: the reference updates a SHADOWING depth_now here (ACMMP.cu:1271, :1282)
                    plane_center = nb_max;
                    plane_intended = nb_max;
                    cost_center = final_costs[max_cost_idx];
                    restricted_cost = restricted_final_costs[max_cost_idx];
                    sel_center = temp_selected_views;
                    sel_dirty = true;
                }
            }
        } else if ((flagbits >> min_cost_idx) & 1u) {
            plane_now = nb_min;
            have_plane_now = true;
            const float depth_before = plane_depth(nb_min, px.dir);
            if (depth_before >= fc.depth_min && depth_before <= fc.depth_max && final_costs[min_cost_idx] < cost_now) {
                depth_now = depth_before;
                plane_center = nb_min;
                plane_intended = nb_min;
                cost_center = final_costs[min_cost_idx];
            }
        }
    } else if ((flagbits >> min_cost_idx) & 1u) {        // ACMMP.cu:1301-1311
        const float depth_before = plane_depth(nb_min, px.dir);
        const bool accept = depth_before >= fc.depth_min && depth_before <= fc.depth_max && final_costs[min_cost_idx] < cost_now;
        plane_now = nb_min;
        have_plane_now = true;
        if (accept) {
            depth_now = depth_before;
            cost_now = final_costs[min_cost_idx];
            sel_center = temp_selected_views;
            sel_dirty = true;
            plane_intended = nb_min;      // registers only: memory keeps the old plane (ACMMP.cu:1307)
        }
    }
    // `float4 plane_hypotheses_now;` is uninitialised in the reference (ACMMP.cu:1301).  The nvcc
    // 12.9 / sm_100 binary keeps the best neighbour's plane there whenever that neighbour exists
    // (as_compiled); the intended meaning is "the plane currently stored for the pixel".
    if (!fc.as_compiled || !have_plane_now) plane_now = plane_intended;
    __syncwarp(FULL);

    // ---- PlaneHypothesisRefinement, ACMMP.cu:797-936 -----------------------------------------
    const bool do_refine = weight_norm > 0.0f;      // group-uniform
    const bool use_prior = kPrior && do_refine && mask_c > 0;
    float cdepth = depth_now;
    float4 temp_plane = plane_now;
    float depth_prior = 0.f;
    {
        const float perturbation = 0.02f;
        const float angle_sigma = angle_sigma_r;
        if (use_prior) depth_prior = plane_depth(prior_plane, px.dir);
        float depth_rand = 0.f, depth_perturbed = 0.f;
        float4 n_rand = make_float4(0.f, 0.f, 0.f, 0.f), n_pert = n_rand;
        if (do_refine && gl == 0) {
            if (use_prior) {
                depth_rand = sample_depth_inv(rs, fmaxf(depth_prior - 3 * depth_sigma, fc.depth_min),
                                              fminf(depth_prior + 3 * depth_sigma, fc.depth_max));
                n_rand = perturbed_normal(rs, px.dir, prior_plane, angle_sigma);
            } else {
                depth_rand = sample_depth_inv(rs, fc.depth_min, fc.depth_max);
                n_rand = random_normal(rs, px.dir);
            }
            float lo = fmaxf((1.0f - perturbation) * depth_now, fc.depth_min);
            float hi = fminf((1.0f + perturbation) * depth_now, fc.depth_max);
            if (!(hi > lo)) { lo = fc.depth_min; hi = fc.depth_max; }
            depth_perturbed = depth_now;
            bool ok = false;
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
                const float cnd = sample_depth_inv(rs, lo, hi);
                if (cnd >= fc.depth_min && cnd <= fc.depth_max) {
                    depth_perturbed = cnd;
                    ok = true;
                    break;
                }
            }
            if (!ok) depth_perturbed = fminf(fmaxf(depth_now, fc.depth_min), fc.depth_max);
            n_pert = perturbed_normal(rs, px.dir, plane_now, perturbation * CUDART_PI_F);
        }
        __syncwarp(FULL);
        depth_rand = __shfl_sync(FULL, depth_rand, gbase);
        depth_perturbed = __shfl_sync(FULL, depth_perturbed, gbase);
        n_rand = shfl_plane(FULL, n_rand, gbase);
        n_pert = shfl_plane(FULL, n_pert, gbase);
        if (do_refine) {
            // candidate gl (0..4): depths {rand, now, rand, now, pert}, normals {now, rand, rand, pert, now}
            if (gl == 0 || gl == 2) cdepth = depth_rand;
            if (gl == 4) cdepth = depth_perturbed;
            if (gl == 1 || gl == 2) temp_plane = n_rand;
            if (gl == 3) temp_plane = n_pert;
            temp_plane.w = plane_offset(temp_plane, px.dir, cdepth);
        }
    }
    __syncwarp(FULL);
    // The 5 refinement hypotheses are evaluated view by view, only for the selected views (the reference
    // evaluates every view and then ignores the zero-weight ones, ACMMP.cu:882-897): quad Q of the group
    // takes the selected views Q, Q+2, ... with all five hypotheses at once (tap-split inside the quad).
    // Tap depths: lanes of quad 0 fill hypotheses 0-2, lanes of quad 1 hypotheses 3-4; both quads read both.
    // What is stored per (hypothesis, view) is ncc + 0.1 * geometric cost, the bracket of ACMMP.cu:890.
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int h = 3 * Q + k;
            const float4 hp = shfl_plane(FULL, temp_plane, gbase + min(h, 4));
            if (h < 5) quad_fill_depths<MODEL, TG::RW, TQS>(fc, aux, px, hp, q, tq + k * kTqPerHyp * TQS, steps);
        }
        __syncwarp(FULL);
        const int n_sel = (do_refine && valid) ? __popc(temp_selected_views) : 0;
        WarpPairs wp;
        wp.build((do_refine && valid) ? temp_selected_views : 0u);
        const int Qw = lane >> 2, g_own = lane >> 3;
        const int rounds = (wp.pre[4] + 7) >> 3;
        for (int r = 0; r < rounds; ++r) {
            int go = g_own, vsel = 0;
            const bool want = wp.get(8 * r + Qw, go, vsel);
            const int o = (tid >> 5) * 4 + go;             // pixel slot of the pair's owner
            PixCtx po = px;                                // same tile row; x from the owner's (unclamped, valid) position
            po.tx = 2 * go + xpar + kHalo;
            po.dx = static_cast<float>(x0 + 2 * go + xpar) - fc.cx;
            const float4 so = sums_all[o];
            po.Sw = so.x; po.Swr = so.y; po.Swrr = so.z;
            float *grp_o = cost_all + (o * 8) * nvp;
            const ViewK c = load_view(s_ncc + vsel);
            FetchLayer fetch;
            fetch.tex = (cudaTextureObject_t)fc.tex_src;
            fetch.layer = vsel;
            // tap depths of the owner: hypotheses 0-2 in the columns of its first quad, 3-4 in those of its second
            quad_ncc<MODEL, 5, TG::RW, WRS, TQS>(
                c, po, aux, wr_all + o, rr_all + o, tq_all + o * 8 + q, fetch, q, want ? 0x1Fu : 0u,
                [&](const int h) { return (h < 3 ? h : h - 3) * kTqPerHyp * TQS + (h < 3 ? 0 : 4); },
                [&](const int h, const float cst) { grp_o[h * nvp + vsel] = cst; }, steps);
        }
        __syncwarp(FULL);
        if (kGeom) {
            // geometric term of every (hypothesis, selected view) pair, dealt out to the 8 lanes of the group with
            // four neighbour-depth loads in flight per lane; the five planes travel through the tap-depth table
            if (gl < 5) {
                tq[0 * TQS] = temp_plane.x; tq[1 * TQS] = temp_plane.y; tq[2 * TQS] = temp_plane.z; tq[3 * TQS] = temp_plane.w;
            }
            __syncwarp(FULL);
            const float *planes5 = tq - gl;
            const int n_pairs = 5 * n_sel;
            for (int p0 = gl; p0 < n_pairs; p0 += 32) {
                float sx[4], sy[4], sd[4];
                int hh[4], vv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pidx = p0 + 8 * k;
                    hh[k] = -1;
                    sd[k] = 0.f;
                    if (pidx < n_pairs) {
                        hh[k] = pidx % 5;
                        vv[k] = (int)__fns(temp_selected_views, 0, pidx / 5 + 1);
                        const float4 hp = make_float4(planes5[hh[k]], planes5[hh[k] + TQS], planes5[hh[k] + 2 * TQS], planes5[hh[k] + 3 * TQS]);
                        sd[k] = __ldg(geom_address<MODEL>(fc, s_vc[vv[k]], px, hp, sx[k], sy[k]));
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (hh[k] >= 0)
                        cost_grp[hh[k] * nvp + vv[k]] += 0.1f * geom_finish<MODEL>(fc, s_vc[vv[k]], px, sx[k], sy[k], sd[k]);
                }
            }
            __syncwarp(FULL);
        }
    }
    {
        const float two_angle_sigma_squared = two_angle_sigma_squared_r;
        float temp_cost = 0.0f;
        float depth_before = 0.f;
        bool cand_ok = false;
        float restricted_temp_cost = 0.f;
        if (do_refine && gl < 5) {
            for (int j = 0; j < nsrc; ++j) {
                const float wj = vw[j];
                if (wj > 0.0f) temp_cost += wj * costrow[j];
            }
            temp_cost /= weight_norm;
            depth_before = plane_depth(temp_plane, px.dir);
            cand_ok = !(depth_before < fc.depth_min || depth_before > fc.depth_max || depth_before >= 1e6f);
            if (use_prior) {
                const float depth_diff = cdepth - depth_prior;
                float angle_cos = dot3(prior_plane, temp_plane);
                angle_cos = fminf(fmaxf(angle_cos, -1.0f), 1.0f);
                const float angle_diff = acosf(angle_cos);
                const float prior = gamma + expf(-depth_diff * depth_diff / two_depth_sigma_squared) *
                                                expf(-angle_diff * angle_diff / two_angle_sigma_squared);
                restricted_temp_cost = expf(-temp_cost * temp_cost / beta) * prior;
            }
        }
        // sequential acceptance over the five candidates (ACMMP.cu:876-935)
        int best = -1;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float tc_i = __shfl_sync(FULL, temp_cost, gbase + i);
            const float rc_i = kPrior ? __shfl_sync(FULL, restricted_temp_cost, gbase + i) : 0.f;
            const int ok_i = __shfl_sync(FULL, (int)cand_ok, gbase + i);
            if (ok_i) {
                if (use_prior) {
                    if (rc_i > restricted_cost) { restricted_cost = rc_i; cost_now = tc_i; best = i; }
                } else {
                    if (tc_i < cost_now) { cost_now = tc_i; best = i; }
                }
            }
        }
        const int src = gbase + (best < 0 ? 0 : best);
        const float4 bp = shfl_plane(FULL, temp_plane, src);
        const float bd = __shfl_sync(FULL, depth_before, src);
        if (best >= 0) {
            plane_now = bp;
            depth_now = bd;
        }
    }

    // ---- write-back (ACMMP.cu:1315-1324) ------------------------------------------------------
    if (gl == 0 && valid) {
        float4 out_plane = plane_now;
        float out_cost = cost_now;
        if (fc.hierarchy) {
            if (!(cost_now < fc.pre_costs[center] - 0.1f)) {
                out_plane = plane_center;
                out_cost = cost_center;
            }
        }
        fc.planes_alt[center] = out_plane;
        fc.costs_alt[center] = out_cost;
        if (sel_dirty) fc.selected_views[center] = sel_center;
        rng_store(fc.rng + 3 * (size_t)center, rs);
    }

    // ---- every warp is through the tile: its buffer is free for the tile after next ---------------------------
    __syncthreads();
    if (tile + 2 * G < ntiles) fill_buffer(buf, tile + 2 * G);
    }   // tile loop
}

// ------------------------------------------------------------------------------------------
// GetDepthandNormal, ACMMP.cu:1351-1364: (n_cam, d) -> (n_world, depth); one float4 per thread
// ------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(256)
k_depth_normal(const __grid_constant__ FrameConst fc)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= fc.W * fc.H) return;
    const int x = idx % fc.W, y = idx / fc.W;
    float4 p = fc.planes[idx];
    p.w = plane_depth(p, pixel_dir<MODEL>(fc, x, y));
    fc.planes[idx] = normal_to_world(fc, p);
}

// CheckerboardFilter, ACMMP.cu:1366-1480: median of up to 21 depths, in place, one colour per
// launch.  All 20 neighbour offsets have odd Manhattan distance, i.e. the other colour, so a
// launch never reads what it writes.  A CTA owns a 32x16 pixel tile (256 pixels of the active colour, one per
// thread) and first stages the depths of the tile + 5 px halo in shared memory with coalesced float4 loads: the
// 21 scattered .w reads per pixel from the float4 map made the direct form L1-bound (0.42 ms per launch at
// 3200x2130, 0.55 TB/s effective).
constexpr int kMedTW = 32, kMedTH = 16, kMedHalo = 5, kMedPW = kMedTW + 2 * kMedHalo, kMedPH = kMedTH + 2 * kMedHalo;

__global__ void __launch_bounds__(256)
k_median_filter(const __grid_constant__ FrameConst fc, const int colour)
{
    __shared__ float sd[kMedPH][kMedPW + 1];
    const int W = fc.W, H = fc.H;
    const int x0 = blockIdx.x * kMedTW, y0 = blockIdx.y * kMedTH;
    float4 *ph = fc.planes;
    for (int i = threadIdx.x; i < kMedPW * kMedPH; i += 256) {
        const int lx = i % kMedPW, ly = i / kMedPW;
        const int gx = x0 + lx - kMedHalo, gy = y0 + ly - kMedHalo;
        float d = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) d = ph[(size_t)gy * W + gx].w;
        sd[ly][lx] = d;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4;                                  // 16 rows x 16 colour pixels
    const int y = y0 + ty;
    const int x = x0 + 2 * (threadIdx.x & 15) + ((y + colour) & 1);
    if (x >= W || y >= H) return;
    const int center = y * W + x;
    if (fc.costs[center] < 0.001f) return;
    const int cx = x - x0 + kMedHalo, cy = y - y0 + kMedHalo;
#define ACMMP_D(dx, dy) sd[cy + (dy)][cx + (dx)]
    float filter[21];
    int index = 0;
    filter[index++] = ACMMP_D(0, 0);
    // same order as the reference (ACMMP.cu:1393-1455); the order does not change the median
    if (y > 0) filter[index++] = ACMMP_D(0, -1);
    if (y > 2) filter[index++] = ACMMP_D(0, -3);
    if (y > 4) filter[index++] = ACMMP_D(0, -5);
    if (y < H - 1) filter[index++] = ACMMP_D(0, 1);
    if (y < H - 3) filter[index++] = ACMMP_D(0, 3);
    if (y < H - 5) filter[index++] = ACMMP_D(0, 5);
    if (x > 0) filter[index++] = ACMMP_D(-1, 0);
    if (x > 2) filter[index++] = ACMMP_D(-3, 0);
    if (x > 4) filter[index++] = ACMMP_D(-5, 0);
    if (x < W - 1) filter[index++] = ACMMP_D(1, 0);
    if (x < W - 3) filter[index++] = ACMMP_D(3, 0);
    if (x < W - 5) filter[index++] = ACMMP_D(5, 0);
    if (y > 0 && x < W - 2) filter[index++] = ACMMP_D(2, -1);
    if (y < H - 1 && x < W - 2) filter[index++] = ACMMP_D(2, 1);
    if (y > 0 && x > 1) filter[index++] = ACMMP_D(-2, -1);
    if (y < H - 1 && x > 1) filter[index++] = ACMMP_D(-2, 1);
    if (x > 0 && y > 2) filter[index++] = ACMMP_D(-1, -2);
    if (x < W - 1 && y > 2) filter[index++] = ACMMP_D(1, -2);
    if (x > 0 && y < H - 2) filter[index++] = ACMMP_D(-1, 2);
    if (x < W - 1 && y < H - 2) filter[index++] = ACMMP_D(1, 2);
#undef ACMMP_D
    // sort_small, ACMMP.cu:36-45
    for (int i = 1; i < index; i++) {
        const float tmp = filter[i];
        int j;
        for (j = i; j >= 1 && tmp < filter[j - 1]; j--) filter[j] = filter[j - 1];
        filter[j] = tmp;
    }
    const int median_index = index / 2;
    if (index % 2 == 0) {
        ph[center].w = (filter[median_index - 1] + filter[median_index]) / 2;
    } else {
        ph[center].w = filter[median_index];
    }
}

// ------------------------------------------------------------------------------------------
// JBU_cu, ACMMP.cu:1558-1616.  image: fine grey image (w x h), depth_in: coarse (sw x sh).
// Both are read at integer texel centres in the reference, i.e. exact values: plain loads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_jbu(const float *__restrict__ image, const int cols, const int rows, const float *__restrict__ depth_in, const int s_width,
      const int s_height, const int Imagescale, float *__restrict__ depth_out)
{
    const int px = blockIdx.x * 16 + (threadIdx.x & 15);
    const int py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= cols || py >= rows) return;
    const int center = py * cols + px;
    const float scale = (float)(1.0 * s_width / cols);
    const float sigmad = 0.50f, sigmar = 25.5f;
    const int WinWidth = Imagescale * Imagescale + 1;
    const int num_neighbors = WinWidth / 2;
    const float o_y = py * scale, o_x = px * scale;
    const float refPix = image[center];
    float total_val = 0.0f, normalizing_factor = 0.0f;
    for (int j = -num_neighbors; j <= num_neighbors; ++j) {
        int r_y = (int)(o_y + j);
        r_y = (r_y > 0 ? (r_y < s_height ? r_y : s_height - 1) : 0);
        int r_ys = py + j;
        r_ys = (r_ys > 0 ? (r_ys < rows ? r_ys : rows - 1) : 0);
        for (int i = -num_neighbors; i <= num_neighbors; ++i) {
            int r_x = (int)(o_x + i);
            r_x = (r_x > 0 ? (r_x < s_width ? r_x : s_width - 1) : 0);
            const float srcPix = __ldg(depth_in + r_y * s_width + r_x);
            int r_xs = px + i;
            r_xs = (r_xs > 0 ? (r_xs < cols ? r_xs : cols - 1) : 0);
            const float neighborPix = __ldg(image + r_ys * cols + r_xs);
            const float sgauss = spatial_gauss(o_x, o_y, (float)r_x, (float)r_y, sigmad);
            const float rgauss = range_gauss(fabsf(refPix - neighborPix), sigmar);
            const float totalgauss = sgauss * rgauss;
            normalizing_factor += totalgauss;
            total_val += srcPix * totalgauss;
        }
    }
    depth_out[center] = total_val / normalizing_factor;
}

// ------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------
// dense W x H image -> border-replicated pitch-linear image (clamp addressing of the reference's
// reference-view texture, ACMMP.cpp:698-704, made explicit so that TMA never leaves the tensor)
__global__ void __launch_bounds__(256)
k_pad_reference(const float *__restrict__ src, const int W, const int H, float *__restrict__ dst, const int pitch)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y;
    if (px >= pitch) return;
    const int sx = min(max(px - kRefPad, 0), W - 1);
    const int sy = min(max(py - kRefPad, 0), H - 1);
    dst[(size_t)py * pitch + px] = src[(size_t)sy * W + sx];
}

// XORWOW state of every pixel right after curand_init(seed, subsequence = y, offset = x):
// row_states[y] = state after the sequence skip (computed on the host from the 2^67-step matrix),
// then x single steps along the row.  One thread per row.
__global__ void __launch_bounds__(64)
k_rng_fill(const uint32_t *__restrict__ row_states, const int W, const int H, uint2 *__restrict__ out)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= H) return;
    Rng s;
    s.d = row_states[6 * y + 0];
    s.v0 = row_states[6 * y + 1]; s.v1 = row_states[6 * y + 2]; s.v2 = row_states[6 * y + 3];
    s.v3 = row_states[6 * y + 4]; s.v4 = row_states[6 * y + 5];
    uint2 *row = out + 3 * (size_t)y * W;
    for (int x = 0; x < W; ++x) {
        rng_store(row + 3 * x, s);
        rng_next(s);
    }
}

// w x h image -> W x H image with the last column / row replicated (layer padding, see fetch_src)
__global__ void __launch_bounds__(256)
k_replicate_pad(const float *__restrict__ src, const int w, const int h, float *__restrict__ dst, const int W, const int H)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y;
    if (px >= W || py >= H) return;
    dst[(size_t)py * W + px] = src[(size_t)min(py, h - 1) * w + min(px, w - 1)];
}

__global__ void __launch_bounds__(256)
k_export_depth(const float4 *__restrict__ planes, const int n, float *__restrict__ depth)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) depth[idx] = planes[idx].w;
}

// CudaPlanarPriorInitialization, ACMMP.cpp:855-863, on the device: masks (1-based triangle id as float)
// -> plane_masks (u32) and the per-pixel prior plane
__global__ void __launch_bounds__(256)
k_expand_prior(const float *__restrict__ masks, const float4 *__restrict__ params, const int n_planes, const int n,
               float4 *__restrict__ prior_planes, uint32_t *__restrict__ plane_masks)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float m = masks[idx];
    plane_masks[idx] = (uint32_t)m;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m > 0) {
        const int id = min(max((int)(m - 1), 0), n_planes - 1);
        p = params[id];
    }
    prior_planes[idx] = p;
}

// ------------------------------------------------------------------------------------------
// Planar-prior stage on the device.  The reference runs it on the CPU between the photometric
// and the prior stage (ACMMP.cpp:904-1011, main.cpp:113-185); here only the Delaunay triangulation stays on the
// host (host/delaunay.cpp).  Every floating-point step is written with round-to-nearest intrinsics in the order the
// host code (host/acmmp_host.cpp, acmmp_main.cpp: PlanarPriorStage) evaluates it: no contraction, no fast-math
// substitution, so that PINHOLE results are bit-identical to the CPU stage (SPHERE differs in the last bit of sin/cos).
// ------------------------------------------------------------------------------------------
struct PriorCam {
    int model, W, H;
    float K0, K2, K4, K5;      // PINHOLE fx, cx, fy, cy
    float cx, cy;              // SPHERE params[1], params[2]
    float depth_min, depth_max;
};

// GetSupportPoints, ACMMP.cpp:904-930: per 5x5 cell (columns outer, rows inner) the first pixel of least cost, kept
// when that cost is below 0.1.  cells[cx * cells_y + cy] = (x, y) or (-1, -1); the host compacts in that order.
__global__ void __launch_bounds__(256)
k_support_cells(const float *__restrict__ costs, const int W, const int H, const int cells_x, const int cells_y, int2 *__restrict__ cells)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cells_x * cells_y) return;
    const int col = (idx / cells_y) * 5, row = (idx % cells_y) * 5;
    float best = 2.0f;
    int2 where = make_int2(-1, -1);
    const int c_end = min(W, col + 5), r_end = min(H, row + 5);
    for (int c = col; c < c_end; ++c)
        for (int r = row; r < r_end; ++r) {
            const float cst = costs[(size_t)r * W + c];
            if (cst < 2.0f && best > cst) {
                where = make_int2(c, r);
                best = cst;
            }
        }
    cells[idx] = (best < 0.1f) ? where : make_int2(-1, -1);
}

// Get3DPointonRefCam, ACMMP.cpp:287-312 (host twin: acmmp_host.cpp)
__device__ __forceinline__ void prior_lift(const PriorCam &cam, const int x, const int y, const float depth, float &X, float &Y, float &Z)
{
    if (cam.model == kModelSphere) {
        const float lon = __fmul_rn(__fmul_rn(__fdiv_rn(__fsub_rn((float)x, cam.cx), (float)cam.W), 2.0f), (float)M_PI);
        const float lat = __fmul_rn(__fdiv_rn(-__fsub_rn((float)y, cam.cy), (float)cam.H), (float)M_PI);
        const float cl = (float)cos((double)lat), sl = (float)sin((double)lat);
        const float co = (float)cos((double)lon), so = (float)sin((double)lon);
        X = __fmul_rn(__fmul_rn(cl, so), depth);
        Y = __fmul_rn(-sl, depth);
        Z = __fmul_rn(__fmul_rn(cl, co), depth);
    } else {
        X = __fdiv_rn(__fmul_rn(depth, __fsub_rn((float)x, cam.K2)), cam.K0);
        Y = __fdiv_rn(__fmul_rn(depth, __fsub_rn((float)y, cam.K5)), cam.K4);
        Z = depth;
    }
}

// GetPriorPlaneParams, ACMMP.cpp:956-989, in the closed form of acmmp_host.cpp: n = (b - a) x (c - a), w = -n.a,
// normalised to |n| = 1 and w >= 0; double precision, one triangle per thread.  tri = 6 ints (x1 y1 x2 y2 x3 y3).
__global__ void __launch_bounds__(256)
k_tri_planes(const int *__restrict__ tri, const int n_tri, const float4 *__restrict__ planes, const PriorCam cam, float4 *__restrict__ params)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tri) return;
    float P[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int x = tri[6 * t + 2 * k], y = tri[6 * t + 2 * k + 1];
        prior_lift(cam, x, y, planes[(size_t)y * cam.W + x].w, P[k][0], P[k][1], P[k][2]);
    }
    const double ux = __dsub_rn((double)P[1][0], (double)P[0][0]), uy = __dsub_rn((double)P[1][1], (double)P[0][1]), uz = __dsub_rn((double)P[1][2], (double)P[0][2]);
    const double vx = __dsub_rn((double)P[2][0], (double)P[0][0]), vy = __dsub_rn((double)P[2][1], (double)P[0][1]), vz = __dsub_rn((double)P[2][2], (double)P[0][2]);
    const double nx = __dsub_rn(__dmul_rn(uy, vz), __dmul_rn(uz, vy));
    const double ny = __dsub_rn(__dmul_rn(uz, vx), __dmul_rn(ux, vz));
    const double nz = __dsub_rn(__dmul_rn(ux, vy), __dmul_rn(uy, vx));
    const double w = -__dadd_rn(__dadd_rn(__dmul_rn(nx, (double)P[0][0]), __dmul_rn(ny, (double)P[0][1])), __dmul_rn(nz, (double)P[0][2]));
    double norm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    if (w < 0) norm = -norm;
    if (norm == 0.0) norm = 1.0;
    params[t] = make_float4((float)__ddiv_rn(nx, norm), (float)__ddiv_rn(ny, norm), (float)__ddiv_rn(nz, norm), (float)__ddiv_rn(w, norm));
}

// The reference's barycentric stepping rasteriser (main.cpp:153-159); the sequential loop lets a later triangle
// overwrite an earlier one, i.e. the largest id wins: atomicMax.  Triangles up to kRasterSmall pixels of longest edge
// (the bulk: support points sit on a 5-pixel grid) take one thread each and the pass appends the others to a list: the
// long ones along image borders and across texture-less regions are 10^5 .. 10^7 steps each (L^2 / 2 for a longest
// edge of L pixels) -- on one thread, and then on one warp, they WERE the stage (66 ms per 3200x2130 view).  They are
// now spread over the whole grid: work item = (long triangle, chunk of kRasterChunk consecutive outer-loop iterations),
// one thread per outer iteration, items dealt to the blocks round-robin.  The outer variable is the reference's running
// float sum (p += step), so a thread reaches its iteration by the same sequence of additions.
constexpr float kRasterSmall = 48.0f;

struct RasterTri {
    float fx1, fx2, fy1, fy2, step, max_edge;
    double dx3, dy3;
};
__device__ __forceinline__ RasterTri raster_setup(const int *__restrict__ tri, const int t)
{
    const int x1 = tri[6 * t], y1 = tri[6 * t + 1], x2 = tri[6 * t + 2], y2 = tri[6 * t + 3], x3 = tri[6 * t + 4], y3 = tri[6 * t + 5];
    // std::sqrt(std::pow(int, 2) + std::pow(int, 2)) in double, stored to float
    const float L01 = (float)__dsqrt_rn((double)((long long)(x1 - x2) * (x1 - x2) + (long long)(y1 - y2) * (y1 - y2)));
    const float L02 = (float)__dsqrt_rn((double)((long long)(x1 - x3) * (x1 - x3) + (long long)(y1 - y3) * (y1 - y3)));
    const float L12 = (float)__dsqrt_rn((double)((long long)(x2 - x3) * (x2 - x3) + (long long)(y2 - y3) * (y2 - y3)));
    RasterTri r;
    r.max_edge = fmaxf(L01, fmaxf(L02, L12));
    r.step = (float)__ddiv_rn(1.0, (double)r.max_edge);
    r.fx1 = (float)x1; r.fx2 = (float)x2; r.fy1 = (float)y1; r.fy2 = (float)y2;
    r.dx3 = (double)x3; r.dy3 = (double)y3;
    return r;
}
// the inner loop of one outer iteration
__device__ __forceinline__ void raster_row(const RasterTri &r, const float p, const int W, const uint32_t id, uint32_t *__restrict__ mask)
{
    const double one_minus_p = __dsub_rn(1.0, (double)p);
    for (float q = 0; (double)q < one_minus_p; q = __fadd_rn(q, r.step)) {
        // p * x1 + q * x2: float;  (1.0 - p - q) * x3: double;  sum in double, truncated
        const double w3 = __dsub_rn(one_minus_p, (double)q);
        const int x = __double2int_rz(__dadd_rn((double)__fadd_rn(__fmul_rn(p, r.fx1), __fmul_rn(q, r.fx2)), __dmul_rn(w3, r.dx3)));
        const int y = __double2int_rz(__dadd_rn((double)__fadd_rn(__fmul_rn(p, r.fy1), __fmul_rn(q, r.fy2)), __dmul_rn(w3, r.dy3)));
        atomicMax(mask + (size_t)y * W + x, id);
    }
}

constexpr int kRasterChunk = 256;          // outer iterations per work item = threads per block of k_tri_raster_long

// long_list: cleared to -1 by the caller; long_items: 0.  A triangle whose items do not fit the list any more (thousands
// of image-sized slivers) is rasterised by its own thread, its reserved slots stay -1.
__global__ void __launch_bounds__(128)
k_tri_raster(const int *__restrict__ tri, const int n_tri, const int W, uint32_t *__restrict__ mask, int *__restrict__ long_list,
             const int list_capacity, int *__restrict__ long_items)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tri) return;
    const RasterTri r = raster_setup(tri, t);
    if (r.max_edge > kRasterSmall) {
        // outer iterations: p = 0, step, 2 step, ... < 1  =>  at most ceil(max_edge) + 1 of them (step = 1 / max_edge in
        // float; the running sum may undershoot 1 once more); an item past the end finds nothing to do
        const int chunks = ((int)r.max_edge + 2 + kRasterChunk - 1) / kRasterChunk;
        const int first = atomicAdd(long_items, chunks);
        if (chunks <= 64 && first + chunks <= list_capacity) {
            for (int c = 0; c < chunks; ++c) long_list[first + c] = t * 64 + c;
            return;
        }
    }
    for (float p = 0; (double)p < 1.0; p = __fadd_rn(p, r.step)) raster_row(r, p, W, (uint32_t)(t + 1), mask);
}

__global__ void __launch_bounds__(kRasterChunk)
k_tri_raster_long(const int *__restrict__ tri, const int W, uint32_t *__restrict__ mask, const int *__restrict__ long_list,
                  const int list_capacity, const int *__restrict__ long_items)
{
    const int n_items = min(*long_items, list_capacity);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int item = long_list[w];
        if (item < 0) continue;
        const int t = item >> 6, chunk = item & 63;
        const RasterTri r = raster_setup(tri, t);
        const int it = chunk * kRasterChunk + threadIdx.x;
        float p = 0;
        for (int i = 0; i < it && (double)p < 1.0; ++i) p = __fadd_rn(p, r.step);
        if ((double)p < 1.0) raster_row(r, p, W, (uint32_t)(t + 1), mask);
    }
}

// main.cpp:168-181 + CudaPlanarPriorInitialization (ACMMP.cpp:847-867): drop the pixels whose prior depth
// (GetDepthFromPlaneParam, ACMMP.cpp:991-1011) leaves the depth range, then mask + per-pixel prior plane
__global__ void __launch_bounds__(256)
k_prior_finish(const uint32_t *__restrict__ mask, const float4 *__restrict__ params, const PriorCam cam, float4 *__restrict__ prior_planes,
               uint32_t *__restrict__ plane_masks)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cam.W * cam.H) return;
    uint32_t m = mask[idx];
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m > 0) {
        p = params[m - 1];
        const int x = idx % cam.W, y = idx / cam.W;
        float d;
        if (cam.model == kModelSphere) {
            const float lon = __fmul_rn(__fmul_rn(__fdiv_rn(__fsub_rn((float)x, cam.cx), (float)cam.W), 2.0f), (float)M_PI);
            const float lat = __fmul_rn(__fdiv_rn(-__fsub_rn((float)y, cam.cy), (float)cam.H), (float)M_PI);
            const float cl = (float)cos((double)lat), sl = (float)sin((double)lat);
            const float dx = __fmul_rn(cl, (float)sin((double)lon)), dy = -sl, dz = __fmul_rn(cl, (float)cos((double)lon));
            const float denom = __fadd_rn(__fadd_rn(__fmul_rn(p.x, dx), __fmul_rn(p.y, dy)), __fmul_rn(p.z, dz));
            d = (fabsf(denom) < 1e-6f) ? 1e6f : __fdiv_rn(-p.w, denom);
        } else {
            const float t0 = __fmul_rn(__fsub_rn((float)x, cam.K2), p.x);
            const float t1 = __fmul_rn(__fmul_rn(__fdiv_rn(cam.K0, cam.K4), __fsub_rn((float)y, cam.K5)), p.y);
            const float t2 = __fmul_rn(cam.K0, p.z);
            d = __fdiv_rn(__fmul_rn(-p.w, cam.K0), __fadd_rn(__fadd_rn(t0, t1), t2));
        }
        if (!(d <= cam.depth_max && d >= cam.depth_min)) {
            m = 0;
            p = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    plane_masks[idx] = m;
    prior_planes[idx] = p;
}

// hierarchy hand-over between pyramid levels, all on the device (ACMMP.cpp:816-840):
// coarse (normal, cost) from the previous level's (n_world, depth) planes + costs
__global__ void __launch_bounds__(256)
k_make_coarse(const float4 *__restrict__ planes, const float *__restrict__ costs, const int n, float4 *__restrict__ coarse,
              float *__restrict__ coarse_depth)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float4 p = planes[idx];
    coarse[idx] = make_float4(p.x, p.y, p.z, costs[idx]);
    coarse_depth[idx] = p.w;
}

// same-size hierarchy level (!upsample): the reference stores the DEPTH, not the cost, in scaled_plane_hypotheses.w
// (ACMMP.cpp:826-828) and the reload branch of RandomInitialization reads it as depth (ACMMP.cu:781-792)
__global__ void __launch_bounds__(256)
k_coarse_w_from_depth(const float *__restrict__ depth, const int n, float4 *__restrict__ coarse)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) coarse[idx].w = depth[idx];
}

// plane = (0, 0, 0, fine depth): what the reference uploads in hierarchy mode (ACMMP.cpp:833-840)
__global__ void __launch_bounds__(256)
k_seed_planes_from_depth(const float *__restrict__ depth, const int n, float4 *__restrict__ planes)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) planes[idx] = make_float4(0.f, 0.f, 0.f, depth[idx]);
}

__global__ void __launch_bounds__(256)
k_import_planes(const float4 *__restrict__ normals_depth, const int n, float4 *__restrict__ planes)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) planes[idx] = normals_depth[idx];
}

} // namespace acmmp
