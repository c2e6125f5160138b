// Development aid: bilinear texture fetch throughput on B200 for the formats the NCC kernel could use.
// Prints fetches/s for a cache-friendly access pattern (each warp samples a compact 2-D neighbourhood).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_fetch(cudaTextureObject_t tex, float *out, int iters, float spread, int W, int H)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // warp origin scattered over the image; lanes form an 8x4 block of nearby sample points
    float x0 = (float)((warp * 37) % (W - 64)) + 16.f + (lane & 7) * spread;
    float y0 = (float)((warp * 101) % (H - 64)) + 16.f + (lane >> 3) * spread;
    float acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const float u = x0 + 0.37f * k + 0.11f * (it & 15);
            const float v = y0 + 0.23f * k + 0.07f * (it & 15);
            acc[k] += tex2D<float>(tex, u, v);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    if (s == 12345.678f) out[0] = s;
}

enum Fmt { R32F, R16F, R16U, R8U };

static cudaTextureObject_t make_tex(Fmt f, int W, int H, cudaArray_t *arr, bool linear)
{
    cudaChannelFormatDesc d;
    std::vector<unsigned char> bytes;
    size_t es = 4;
    if (f == R32F) { d = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat); es = 4; }
    if (f == R16F) { d = cudaCreateChannelDescHalf(); es = 2; }
    if (f == R16U) { d = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned); es = 2; }
    if (f == R8U) { d = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned); es = 1; }
    bytes.resize((size_t)W * H * es);
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        const unsigned v = (unsigned)((i * 2654435761u) >> 24);
        if (f == R32F) ((float *)bytes.data())[i] = (float)v;
        if (f == R16F) ((__half *)bytes.data())[i] = __float2half((float)v);
        if (f == R16U) ((unsigned short *)bytes.data())[i] = (unsigned short)(v * 257);
        if (f == R8U) bytes[i] = (unsigned char)v;
    }
    cudaMallocArray(arr, &d, W, H);
    cudaMemcpy2DToArray(*arr, 0, 0, bytes.data(), W * es, W * es, H, cudaMemcpyHostToDevice);
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = *arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = linear ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = (f == R16U || f == R8U) ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t t = 0;
    cudaCreateTextureObject(&t, &res, &td, nullptr);
    return t;
}

int main()
{
    const int W = 3200, H = 2130;
    float *out;
    cudaMalloc(&out, 4);
    const char *names[4] = {"R32F", "R16F", "R16unorm", "R8unorm"};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int linear = 1; linear >= 0; --linear)
        for (int f = 0; f < 4; ++f) {
            cudaArray_t arr;
            cudaTextureObject_t tex = make_tex((Fmt)f, W, H, &arr, linear != 0);
            for (float spread : {0.6f, 2.5f}) {
                const int blocks = 148 * 16, threads = 256, iters = 400;
                k_fetch<8><<<blocks, threads>>>(tex, out, 20, spread, W, H);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                k_fetch<8><<<blocks, threads>>>(tex, out, iters, spread, W, H);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                const double fetches = (double)blocks * threads * iters * 8;
                printf("{\"format\": \"%s\", \"filter\": \"%s\", \"spread_px\": %.1f, \"ms\": %.3f, \"gfetch_per_s\": %.1f, \"err\": \"%s\"}\n",
                       names[f], linear ? "linear" : "point", spread, ms, fetches / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
            }
            cudaDestroyTextureObject(tex);
            cudaFreeArray(arr);
        }
    return 0;
}
