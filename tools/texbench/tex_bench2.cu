// Development aid (round 1): what bounds bilinear fetches in the access pattern of k_pass, and whether
// narrower texel formats reproduce the R32F filter result.
//   part 1: throughput when the 4 lanes of a QUAD sample within `qspread` px of each other and the 8
//           quads of a warp are unrelated (== 4 hypotheses of one pixel per quad, 4 pixels per warp)
//   part 2: max |difference| (grey levels) of R8unorm / R16unorm / R16F bilinear against R32F bilinear
//           on (a) an integer-valued 0..255 image, (b) a fractional-valued image
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int ILP>
__global__ void k_fetch_quads(cudaTextureObject_t tex, float *out, int iters, float qspread, int W, int H)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int quad = gtid >> 2;
    const unsigned hq = hash32(quad * 2654435761u + 17u), hl = hash32(gtid * 40503u + 3u);
    // quad origin anywhere in the image; lanes of the quad inside a qspread x qspread box
    float x0 = 16.f + (float)(hq % (unsigned)(W - 96)) + qspread * (float)(hl & 1023) * (1.f / 1024.f);
    float y0 = 16.f + (float)((hq >> 12) % (unsigned)(H - 96)) + qspread * (float)((hl >> 10) & 1023) * (1.f / 1024.f);
    float acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = 0.f;
    for (int it = 0; it < iters; ++it) {
        // walk a 6x6 window with stride ~2 px like the NCC taps
        const float bx = x0 + 2.1f * (float)(it % 6), by = y0 + 1.9f * (float)((it / 6) % 6);
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] += tex2DLod<float>(tex, bx + 0.13f * k, by + 2.05f * k, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    if (s == 12345.678f) out[0] = s;
}

__global__ void k_sample(cudaTextureObject_t tex, const float2 *pts, int n, float scale, float *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2DLod<float>(tex, pts[i].x, pts[i].y, 0.f) * scale;
}

enum Fmt { R32F, R16F, R16U, R8U };

static cudaTextureObject_t make_tex(Fmt f, int W, int H, const std::vector<float> &img, cudaArray_t *arr)
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
        const float v = img[i];
        if (f == R32F) ((float *)bytes.data())[i] = v;
        if (f == R16F) ((__half *)bytes.data())[i] = __float2half(v);
        if (f == R16U) ((unsigned short *)bytes.data())[i] = (unsigned short)lrintf(v * (65535.f / 255.f));
        if (f == R8U) bytes[i] = (unsigned char)lrintf(v);
    }
    cudaMallocArray(arr, &d, W, H);
    cudaMemcpy2DToArray(*arr, 0, 0, bytes.data(), W * es, W * es, H, cudaMemcpyHostToDevice);
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = *arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
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
    // smooth-ish integer image (band-limited noise look-alike) and a fractional twin
    std::vector<float> img_int((size_t)W * H), img_frac((size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float v = 127.5f + 60.f * sinf(0.31f * x + 0.17f * y) + 50.f * cosf(0.23f * y - 0.11f * x) +
                            15.f * sinf(1.3f * x) * cosf(1.1f * y);
            const float c = fminf(fmaxf(v, 0.f), 255.f);
            img_int[(size_t)y * W + x] = rintf(c);
            img_frac[(size_t)y * W + x] = c;
        }
    // ---- part 1
    for (int f = 0; f < 4; ++f) {
        if (f == 1) continue;
        cudaArray_t arr;
        cudaTextureObject_t tex = make_tex((Fmt)f, W, H, img_int, &arr);
        for (float qs : {0.0f, 0.5f, 1.0f, 2.0f, 4.0f, 8.0f, 32.0f}) {
            const int blocks = 148 * 16, threads = 256, iters = 360;
            k_fetch_quads<6><<<blocks, threads>>>(tex, out, 36, qs, W, H);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_fetch_quads<6><<<blocks, threads>>>(tex, out, iters, qs, W, H);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fetches = (double)blocks * threads * iters * 6;
            printf("{\"part\": 1, \"format\": \"%s\", \"quad_spread_px\": %.1f, \"ms\": %.3f, \"gfetch_per_s\": %.1f, \"err\": \"%s\"}\n",
                   names[f], qs, ms, fetches / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
        cudaDestroyTextureObject(tex);
        cudaFreeArray(arr);
    }
    // ---- part 2
    const int n = 1 << 20;
    std::vector<float2> pts(n);
    srand(7);
    for (int i = 0; i < n; ++i) pts[i] = make_float2(0.5f + (W - 1) * (rand() / (float)RAND_MAX), 0.5f + (H - 1) * (rand() / (float)RAND_MAX));
    float2 *dp;
    float *dv;
    cudaMalloc(&dp, sizeof(float2) * n);
    cudaMalloc(&dv, sizeof(float) * n);
    cudaMemcpy(dp, pts.data(), sizeof(float2) * n, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant) {
        const std::vector<float> &img = variant ? img_frac : img_int;
        std::vector<float> ref(n), got(n);
        for (int f = 0; f < 4; ++f) {
            cudaArray_t arr;
            cudaTextureObject_t tex = make_tex((Fmt)f, W, H, img, &arr);
            const float scale = (f == R16U || f == R8U) ? 255.f : 1.f;
            k_sample<<<(n + 255) / 256, 256>>>(tex, dp, n, scale, dv);
            cudaMemcpy(f == 0 ? ref.data() : got.data(), dv, sizeof(float) * n, cudaMemcpyDeviceToHost);
            if (f > 0) {
                double mx = 0, sum = 0;
                int exact = 0;
                for (int i = 0; i < n; ++i) {
                    const double d = fabs((double)got[i] - (double)ref[i]);
                    mx = d > mx ? d : mx;
                    sum += d;
                    exact += (got[i] == ref[i]);
                }
                printf("{\"part\": 2, \"image\": \"%s\", \"format\": \"%s\", \"max_abs_diff\": %.6g, \"mean_abs_diff\": %.6g, \"bit_equal_frac\": %.4f}\n",
                       variant ? "fractional" : "integer", names[f], mx, sum / n, exact / (double)n);
            }
            cudaDestroyTextureObject(tex);
            cudaFreeArray(arr);
        }
    }
    return 0;
}
