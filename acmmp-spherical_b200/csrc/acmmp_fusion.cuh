// acmmp_fusion.cuh -- depth-map fusion on the device.
//
//   k_fuse_view    : SimpleFusionKernel, ACMMP.cu:1664-1814 -- per pixel of a reference view: lift to the world point,
//                    project into every source view, read that view's depth / normal at the rounded pixel, lift it,
//                    project it back; a source view is consistent when reprojection error < 1 px, relative depth
//                    difference < 1 % and the normals are < 0.149 rad apart; pixels with >= 3 consistent views
//                    (itself included) give one point = the average of the consistent points / normals / colours.
//   k_scan_blocks, k_compact_points : the points of a view compacted ON THE DEVICE in pixel order -- the reference copies
//                    36 bytes + a flag for EVERY pixel to the host and filters there (ACMMP.cu:2056-2076).
//
// Memory layout: the reference reads depth / normal / colour through textures with point (depth, normal) or linear
// (colour) filtering at INTEGER coordinates without the half-texel offset (ACMMP.cu:1685-1699, :1735-1758, :1786): the point
// reads are exact texels -> plain linear memory here (float depth, float4 normal: one 16-byte load); the linear read at
// an integer coordinate is the mean of the 2x2 block ending at that texel (both bilinear fractions are exactly 0.5, clamp
// addressing) -> four loads.  The kernel is bound by these scattered reads (L2 / HBM), one thread per pixel, 256-pixel
// row-major blocks so that a block's points are contiguous in the output.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "../../include/acmmp_b200.h"

namespace acmmp {

constexpr int kFuseMaxSrc = 32;          // FusionProblem, ACMMP.cu:1656-1661
constexpr int kFuseBlock = 256;

struct FusionViewDev {
    acmmp_camera cam;                    // scaled to the depth map's size (RescaleImageAndCamera, ACMMP.cpp:213-245)
    const float *depth;
    const float4 *normal;                // world-frame normal (normals.dmb), w unused
    const float *gray;                   // grey levels 0..255 at the depth map's size (colour = (g, g, g))
    const uchar4 *bgr;                   // optional: the colour image at that size, (B, G, R, -) per pixel; null = use gray
};

struct FusionProblemDev {
    int num_src;
    int src[kFuseMaxSrc];                // index into the view table, -1 = not available
};

// Get3DPointonWorld_cu, ACMMP.cu:565-600
__device__ __forceinline__ float3 fuse_lift(const acmmp_camera &cam, const float x, const float y, const float depth)
{
    float3 pc;
    if (cam.model == ACMMP_MODEL_SPHERE) {
        const float lon = (x - cam.params[1]) / static_cast<float>(cam.width) * 2.0f * CUDART_PI_F;
        const float lat = -(y - cam.params[2]) / static_cast<float>(cam.height) * CUDART_PI_F;
        pc.x = cosf(lat) * sinf(lon) * depth;
        pc.y = -sinf(lat) * depth;
        pc.z = cosf(lat) * cosf(lon) * depth;
    } else {
        pc.x = depth * (x - cam.K[2]) / cam.K[0];
        pc.y = depth * (y - cam.K[5]) / cam.K[4];
        pc.z = depth;
    }
    float3 tmp;
    tmp.x = cam.R[0] * pc.x + cam.R[3] * pc.y + cam.R[6] * pc.z;
    tmp.y = cam.R[1] * pc.x + cam.R[4] * pc.y + cam.R[7] * pc.z;
    tmp.z = cam.R[2] * pc.x + cam.R[5] * pc.y + cam.R[8] * pc.z;
    float3 C;
    C.x = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
    C.y = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
    C.z = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
    return make_float3(tmp.x + C.x, tmp.y + C.y, tmp.z + C.z);
}

// ProjectonCamera_cu, ACMMP.cu:602-644
__device__ __forceinline__ void fuse_project(const acmmp_camera &cam, const float3 X, float2 &pt, float &depth)
{
    float3 tmp;
    tmp.x = cam.R[0] * X.x + cam.R[1] * X.y + cam.R[2] * X.z + cam.t[0];
    tmp.y = cam.R[3] * X.x + cam.R[4] * X.y + cam.R[5] * X.z + cam.t[1];
    tmp.z = cam.R[6] * X.x + cam.R[7] * X.y + cam.R[8] * X.z + cam.t[2];
    if (cam.model == ACMMP_MODEL_SPHERE) {
        depth = sqrtf(tmp.x * tmp.x + tmp.y * tmp.y + tmp.z * tmp.z);
        if (depth < 1e-6f) {
            pt.x = cam.params[1];
            pt.y = cam.params[2];
            return;
        }
        const float latitude = -asinf(tmp.y / depth);
        const float longitude = atan2f(tmp.x, tmp.z);
        pt.x = (longitude / (2.0f * CUDART_PI_F)) * static_cast<float>(cam.width) + cam.params[1];
        pt.y = (-latitude / CUDART_PI_F) * static_cast<float>(cam.height) + cam.params[2];
    } else {
        depth = tmp.z;
        pt.x = (cam.K[0] * tmp.x + cam.K[1] * tmp.y + cam.K[2] * tmp.z) / depth;
        pt.y = (cam.K[3] * tmp.x + cam.K[4] * tmp.y + cam.K[5] * tmp.z) / depth;
    }
}

// tex2D<float4>(image, c, r) of a linear-filtered texture at an integer coordinate (no half-texel offset): the mean of
// the texels (c-1 .. c) x (r-1 .. r) with clamp addressing, as levels 0..255.  Returned in the order the reference sums
// them into PointList::color (ACMMP.cu:1704-1708, :1769-1771): (B, G, R) -- its texels are RGBA after cvtColor(BGR2RGBA)
// and it takes .z first; the PLY writer reads them back in that order (ACMMP.cpp:509-511).
__device__ __forceinline__ float3 fuse_colour(const FusionViewDev &v, const int c, const int r)
{
    const int w = v.cam.width, h = v.cam.height;
    const int c0 = min(max(c - 1, 0), w - 1), c1 = min(max(c, 0), w - 1);
    const int r0 = min(max(r - 1, 0), h - 1), r1 = min(max(r, 0), h - 1);
    if (v.bgr) {
        const uchar4 a = __ldg(v.bgr + (size_t)r0 * w + c0), b = __ldg(v.bgr + (size_t)r0 * w + c1);
        const uchar4 d = __ldg(v.bgr + (size_t)r1 * w + c0), e = __ldg(v.bgr + (size_t)r1 * w + c1);
        return make_float3(0.25f * (((float)a.x + (float)b.x) + ((float)d.x + (float)e.x)),
                           0.25f * (((float)a.y + (float)b.y) + ((float)d.y + (float)e.y)),
                           0.25f * (((float)a.z + (float)b.z) + ((float)d.z + (float)e.z)));
    }
    const float a = __ldg(v.gray + (size_t)r0 * w + c0), b = __ldg(v.gray + (size_t)r0 * w + c1);
    const float d = __ldg(v.gray + (size_t)r1 * w + c0), e = __ldg(v.gray + (size_t)r1 * w + c1);
    const float g = 0.25f * ((a + b) + (d + e));
    return make_float3(g, g, g);
}

// One thread per pixel of the reference view, row-major blocks of kFuseBlock pixels.
__global__ void __launch_bounds__(kFuseBlock)
k_fuse_view(const FusionViewDev *__restrict__ views, const int ref, const __grid_constant__ FusionProblemDev problem,
            acmmp_point *__restrict__ dense, unsigned char *__restrict__ flags, int *__restrict__ block_counts)
{
    const FusionViewDev &rv = views[ref];
    const int width = rv.cam.width, height = rv.cam.height;
    const int idx = blockIdx.x * kFuseBlock + threadIdx.x;
    bool valid = false;
    if (idx < width * height) {
        const int c = idx % width, r = idx / width;
        const float ref_depth = __ldg(rv.depth + idx);
        if (ref_depth > 0.0f) {
            const acmmp_camera ref_cam = rv.cam;
            const float3 PointX = fuse_lift(ref_cam, static_cast<float>(c), static_cast<float>(r), ref_depth);
            const float4 rn = __ldg(rv.normal + idx);
            float3 point_sum = PointX;
            float3 normal_sum = make_float3(rn.x, rn.y, rn.z);
            float3 colour_sum = fuse_colour(rv, c, r);           // the reference's texels are x / 255, multiplied back by 255
            int num_consistent = 1;
            for (int j = 0; j < problem.num_src; ++j) {
                const int s = problem.src[j];
                if (s < 0) continue;
                const FusionViewDev &sv = views[s];
                float2 proj;
                float proj_depth;
                fuse_project(sv.cam, PointX, proj, proj_depth);
                const int src_c = static_cast<int>(proj.x + 0.5f);
                const int src_r = static_cast<int>(proj.y + 0.5f);
                if (src_c < 0 || src_c >= sv.cam.width || src_r < 0 || src_r >= sv.cam.height) continue;
                const size_t sidx = (size_t)src_r * sv.cam.width + src_c;
                const float src_depth = __ldg(sv.depth + sidx);
                if (src_depth <= 0.0f) continue;
                const float3 Xs = fuse_lift(sv.cam, static_cast<float>(src_c), static_cast<float>(src_r), src_depth);
                float2 reproj;
                float dummy;
                fuse_project(ref_cam, Xs, reproj, dummy);
                const float reproj_error = hypotf(c - reproj.x, r - reproj.y);
                const float relative_depth_diff = fabsf(proj_depth - src_depth) / src_depth;
                const float4 sn = __ldg(sv.normal + sidx);
                float dot_product = rn.x * sn.x + rn.y * sn.y + rn.z * sn.z;
                dot_product = fmaxf(-1.0f, fminf(1.0f, dot_product));
                const float angle = acosf(dot_product);
                if (reproj_error < 1.0 && relative_depth_diff < 0.01f && angle < 0.149f) {
                    point_sum.x += Xs.x; point_sum.y += Xs.y; point_sum.z += Xs.z;
                    normal_sum.x += sn.x; normal_sum.y += sn.y; normal_sum.z += sn.z;
                    const float3 cs = fuse_colour(sv, src_c, src_r);
                    colour_sum.x += cs.x; colour_sum.y += cs.y; colour_sum.z += cs.z;
                    num_consistent++;
                }
            }
            if (num_consistent >= 3) {
                acmmp_point p;
                p.coord[0] = point_sum.x / num_consistent;
                p.coord[1] = point_sum.y / num_consistent;
                p.coord[2] = point_sum.z / num_consistent;
                float3 n = make_float3(normal_sum.x / num_consistent, normal_sum.y / num_consistent, normal_sum.z / num_consistent);
                const float len = hypotf(hypotf(n.x, n.y), n.z);
                if (len > 0.0f) { n.x /= len; n.y /= len; n.z /= len; }
                p.normal[0] = n.x; p.normal[1] = n.y; p.normal[2] = n.z;
                p.color[0] = colour_sum.x / num_consistent;
                p.color[1] = colour_sum.y / num_consistent;
                p.color[2] = colour_sum.z / num_consistent;
                dense[idx] = p;
                valid = true;
            }
        }
        flags[idx] = valid ? 1 : 0;
    }
    const int n = __syncthreads_count(valid);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = n;
}

// exclusive scan of the per-block counts in place (one CTA; a few 10^4 blocks at most); total -> *total
__global__ void __launch_bounds__(1024)
k_scan_blocks(int *__restrict__ counts, const int n, int *__restrict__ total)
{
    __shared__ int part[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {            // Hillis-Steele inclusive scan of the partial sums
        const int v = (threadIdx.x >= off) ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;                      // exclusive prefix of this thread's chunk
    for (int i = lo; i < hi; ++i) {
        const int c = counts[i];
        counts[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) *total = part[1023];
}

// the valid points of a block, in pixel order, to out[block offset + rank]
__global__ void __launch_bounds__(kFuseBlock)
k_compact_points(const acmmp_point *__restrict__ dense, const unsigned char *__restrict__ flags, const int *__restrict__ block_offsets,
                 const int npx, acmmp_point *__restrict__ out, const int capacity)
{
    __shared__ int warp_base[kFuseBlock / 32];
    const int idx = blockIdx.x * kFuseBlock + threadIdx.x;
    const bool valid = idx < npx && flags[idx];
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_base[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < kFuseBlock / 32; ++w) {
            const int c = warp_base[w];
            warp_base[w] = run;
            run += c;
        }
    }
    __syncthreads();
    if (valid) {
        const int pos = block_offsets[blockIdx.x] + warp_base[warp] + __popc(ballot & ((1u << lane) - 1u));
        if (pos < capacity) out[pos] = dense[idx];
    }
}

// The same compaction straight into the PLY's 27-byte vertex records (StoreColorPlyFileBinaryPointCloud, ACMMP.cpp:481-534:
// x y z nx ny nz as little-endian floats, then red green blue = (char)(int) of color.z, .y, .x; a point with a non-finite
// coordinate is written at the origin, :505-508): the host appends the bytes to the file as they are.
__global__ void __launch_bounds__(kFuseBlock)
k_compact_ply(const acmmp_point *__restrict__ dense, const unsigned char *__restrict__ flags, const int *__restrict__ block_offsets,
              const int npx, unsigned char *__restrict__ out, const int capacity)
{
    __shared__ int warp_base[kFuseBlock / 32];
    const int idx = blockIdx.x * kFuseBlock + threadIdx.x;
    const bool valid = idx < npx && flags[idx];
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_base[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < kFuseBlock / 32; ++w) {
            const int c = warp_base[w];
            warp_base[w] = run;
            run += c;
        }
    }
    __syncthreads();
    if (valid) {
        const int pos = block_offsets[blockIdx.x] + warp_base[warp] + __popc(ballot & ((1u << lane) - 1u));
        if (pos < capacity) {
            const acmmp_point p = dense[idx];
            float v[6] = {p.coord[0], p.coord[1], p.coord[2], p.normal[0], p.normal[1], p.normal[2]};
            const float big = 3.402823466e+38f;
            if (!(v[0] < big && v[0] > -big) || !(v[1] < big && v[1] > -big) || !(v[2] < big && v[2] >= -big)) v[0] = v[1] = v[2] = 0.0f;
            unsigned char *o = out + (size_t)27 * pos;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const unsigned u = __float_as_uint(v[k]);
                o[4 * k] = (unsigned char)u; o[4 * k + 1] = (unsigned char)(u >> 8); o[4 * k + 2] = (unsigned char)(u >> 16); o[4 * k + 3] = (unsigned char)(u >> 24);
            }
            o[24] = (unsigned char)(int)p.color[2];
            o[25] = (unsigned char)(int)p.color[1];
            o[26] = (unsigned char)(int)p.color[0];
        }
    }
}

// (H, W, 3) normals -> float4 per pixel
__global__ void __launch_bounds__(256)
k_pack_normals(const float *__restrict__ n3, const int npx, float4 *__restrict__ n4)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < npx) n4[idx] = make_float4(n3[3 * idx], n3[3 * idx + 1], n3[3 * idx + 2], 0.f);
}

} // namespace acmmp
