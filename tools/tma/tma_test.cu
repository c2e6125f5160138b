// Development aid: minimal TMA 2D tile load variants to find which form the B200 accepts.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BW, int BH, bool TILE_Q, bool INIT_FENCE>
__device__ void body(const CUtensorMap *tm, float *out, int c0, int c1)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 8192);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        if (INIT_FENCE) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BW * BH * 4) : "memory");
        if (TILE_Q)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tile)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tile)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

template <int BW, int BH, bool TILE_Q, bool INIT_FENCE>
__global__ void k_param(const __grid_constant__ CUtensorMap tm, float *out, int c0, int c1) { body<BW, BH, TILE_Q, INIT_FENCE>(&tm, out, c0, c1); }

template <int BW, int BH, bool TILE_Q, bool INIT_FENCE>
__global__ void k_global(const CUtensorMap *tm, float *out, int c0, int c1) { body<BW, BH, TILE_Q, INIT_FENCE>(tm, out, c0, c1); }

struct Pad { int a[54]; };   // pushes the tensor map to a later, still 64-byte aligned, parameter offset
template <int BW, int BH>
__global__ void k_param_late(const __grid_constant__ Pad pad, const __grid_constant__ CUtensorMap tm, float *out, int c0, int c1)
{
    if (pad.a[0] == 12345) out[0] = 1.f;
    body<BW, BH, true, true>(&tm, out, c0, c1);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 336, H = 256;
    std::vector<float> h((size_t)W * H);
    for (int i = 0; i < W * H; ++i) h[i] = (float)i;
    float *d = nullptr, *out = nullptr;
    cudaMalloc(&d, sizeof(float) * W * H);
    cudaMalloc(&out, sizeof(float) * 64 * 64);
    cudaMemcpy(d, h.data(), sizeof(float) * W * H, cudaMemcpyHostToDevice);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    int bw = 28, bh = 18;
    if (variant == 2) { bw = 32; bh = 16; }
    if (variant == 6) { bw = 20; bh = 18; }
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, variant == 5 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode=%d\n", variant, (int)r);
    const int c0 = 3, c1 = 3;
    const size_t smem = 16384;
    switch (variant) {
    case 0: k_param<28, 18, true, true><<<1, 128, smem>>>(tm, out, c0, c1); break;      // what the library does
    case 1: k_param<28, 18, false, true><<<1, 128, smem>>>(tm, out, c0, c1); break;     // no .tile qualifier
    case 2: k_param<32, 16, true, true><<<1, 128, smem>>>(tm, out, c0, c1); break;      // 128-byte rows
    case 3: {                                                                            // descriptor in global memory
        CUtensorMap *dtm = nullptr;
        cudaMalloc(&dtm, sizeof(CUtensorMap));
        cudaMemcpy(dtm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        k_global<28, 18, true, true><<<1, 128, smem>>>(dtm, out, c0, c1);
        break;
    }
    case 4: k_param<28, 18, true, false><<<1, 128, smem>>>(tm, out, c0, c1); break;     // no mbarrier_init fence
    case 5: k_param<28, 18, true, true><<<1, 128, smem>>>(tm, out, c0, c1); break;      // L2 promotion none
    case 6: k_param<20, 18, true, true><<<1, 128, smem>>>(tm, out, c0, c1); break;      // the pass kernel's box
    case 7: { Pad pad; memset(&pad, 0, sizeof(pad)); k_param_late<28, 18><<<1, 128, smem>>>(pad, tm, out, c0, c1); break; }
    case 8: k_param<28, 18, true, true><<<1, 128, smem>>>(tm, out, 4, c1); break;       // 16-byte aligned start column
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(64 * 64, -1.f);
    cudaMemcpy(o.data(), out, sizeof(float) * bw * bh, cudaMemcpyDeviceToHost);
    const int cc0 = variant == 8 ? 4 : c0;
    int bad = 0;
    for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x)
            if (o[y * bw + x] != (float)((y + c1) * W + x + cc0)) bad++;
    printf("variant %d: sync=%s mismatches=%d first=%g\n", variant, cudaGetErrorString(e), bad, o[0]);
    return e == cudaSuccess && bad == 0 ? 0 : 1;
}
