// acmmp_types.cuh -- POD argument blocks shared by host code and kernels.
//
// The reference passes `Camera` (120 B) and `PatchMatchParams` (68 B) by value into every
// device helper (reference ACMMP.cu:405-412, main.h:40-54) and re-derives the camera centre
// and both rigid transforms per NCC sample (ACMMP.cu:585-599, :607-614).  Here everything
// that depends only on (reference view, source view) is folded on the host, in double
// precision, into one `ViewConst` per source view.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace acmmp {

constexpr int kMaxSrc = 32;          // reference ACMMP.cu:522, 32-bit view mask
constexpr int kModelPinhole = 0;     // reference main.h:35-38
constexpr int kModelSphere = 11;
constexpr int kHalo = 5;             // patch_size 11 -> radius 5   (ACMMP.h:34, ACMMP.cu:415)
constexpr int kTaps = 36;            // i,j in {-5,-3,-1,1,3,5}       (ACMMP.cu:450-451)
constexpr int kRefPad = 9;           // border replication of the pitch-linear reference image; 9 - kHalo = 4 makes every
                                     // TMA box start 16-byte aligned (tile origins are multiples of 8): without swizzle the
                                     // B200 TMA unit faults ("illegal instruction") on box starts that are not 16-byte aligned

// Per source view, 72 words.
struct ViewConst {
    // PINHOLE forward (ref pixel -> src pixel):  x~ = t * (Mx*(x-cx) + My*(y-cy) + Mz) + b
    //   with M = K_s R_s R_r^T (third row: depth), Mx = M[:,0]/fx_r, My = M[:,1]/fy_r, Mz = M[:,2],
    //   b = K_s (t_s - R_s R_r^T t_r).                      (ACMMP.cu:579-599 then :607-643)
    float Mx[3], My[3], Mz[3], b[3];
    // same with the +0.5 texel-centre offset of the bilinear fetch (ACMMP.cu:476) folded in:
    // row0 += 0.5*row2, row1 += 0.5*row2
    float Fx[3], Fy[3], Fz[3], fb[3];
    // PINHOLE inverse (src pixel + src depth -> ref pixel): y~ = d*(Ix*(sx-cxs)+Iy*(sy-cys)+Iz)+ib
    float Ix[3], Iy[3], Iz[3], ib[3];
    // SPHERE: X_s = R X_r + t ;  X_r = Ri X_s + ti
    float R[9], t[3];
    float Ri[9], ti[3];
    float cx, cy;        // source principal point (PINHOLE K[2],K[5]; SPHERE params[1],params[2])
    float Wf, Hf;        // source image size
    unsigned long long tex;      // R32F bilinear texture of the source image
    const float *depth;          // neighbour depth map (geom consistency), dense float32, or null
    int dW, dH;                  // its size
    int pad_[2];
};
static_assert(sizeof(ViewConst) == 72 * 4, "ViewConst layout");

// Per source view, what the NCC sample loop needs, 16 floats.  Lives in KERNEL PARAMETER space
// (constant bank) so that a warp-uniform view index turns into uniform-register operands and a
// uniform texture handle -- no per-lane register copies, no non-uniform-handle replay loop.
//   PINHOLE: a[0..2]=Fx a[3..5]=Fy a[6..8]=Fz a[9..11]=fb a[12]=W+0.5 a[13]=H+0.5
//   SPHERE : a[0..8]=R  a[9..11]=t a[12]=cx a[13]=cy a[14]=W a[15]=H
struct alignas(16) NccConst {
    float a[16];
};
// The source views exist twice on the device: as one 2-D R32F texture per view (tex[v]: the handle is
// warp-uniform in the view loops, the fetch is a plain TEX.LZ with no extra operands) and as the layers of
// ONE layered texture (FrameConst::tex_src) for the places where the lanes of a warp work on different
// views at the same time (one handle, per-lane layer).
struct NccTable {
    NccConst c[kMaxSrc];
    unsigned long long tex[kMaxSrc];
};
static_assert(sizeof(NccTable) == kMaxSrc * 72, "NccTable layout");

// Per reference view + stage; passed by value (__grid_constant__).
struct FrameConst {
    int W, H;              // reference image size
    int model;             // kModelPinhole / kModelSphere (all views share one model)
    int nsrc;              // num_images - 1
    float cx, cy;          // PINHOLE K[2],K[5]; SPHERE params[1],params[2]
    float ifx, ify;        // PINHOLE 1/K[0], 1/K[4]
    float Wf, Hf;
    float R[9];            // world -> reference camera
    float depth_min, depth_max;   // params.depth_min/max (ACMMP.cpp:645-646)
    int geom, prior, hierarchy, upsample;
    int scaled_cols, scaled_rows;
    int as_compiled;       // plane_hypotheses_now semantics, see acmmp_b200.h
    int ref_pitch;         // floats per row of the padded reference image
    int use_tma;           // 1: tile staged by a TMA bulk-tensor copy; 0: same tile by plain loads (debug aid)
    float tap_prune;       // SPHERE: window taps whose bilateral weight is below tap_prune * (sum of the 36) are not sampled; 0 = off
    unsigned long long tex_src;   // layered R32F bilinear texture of the source views
    const float *ref_padded;      // (H + 2*kRefPad) rows, border replicated
    const ViewConst *views;       // nsrc entries (device)
    // state
    float4 *planes;        // working plane hypotheses (n_cam, d) / (n_world, depth) after finalize
    float4 *planes_alt;    // ping-pong twin of `planes`
    float *costs;
    float *costs_alt;
    float *pre_costs;
    uint32_t *selected_views;
    uint2 *rng;            // 3 x uint2 per pixel: {d,v0},{v1,v2},{v3,v4}
    const uint2 *rng_seeded;      // state right after curand_init(seed, y, x)
    const float4 *prior_planes;
    const uint32_t *plane_masks;
    const float4 *coarse_planes;  // hierarchy: (normal, cost|depth) at scaled_cols x scaled_rows
};

} // namespace acmmp
