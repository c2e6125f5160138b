// acmmp_api.cu -- host side of libacmmp_b200.so: context, uploads, launch logic, C ABI.
//
// Replaces the device-facing half of the reference host class: ACMMP::CudaSpaceInitialization
// (ACMMP.cpp:681-845), ACMMP::CudaPlanarPriorInitialization (:847-867), ACMMP::RunPatchMatch
// (ACMMP.cu:1506-1556), RunJBU / JBU::CudaRun (ACMMP.cpp:1071-1122, ACMMP.cu:1617-1649).
// No CPU fallback exists: without a usable CUDA device every compute entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/acmmp_b200.h"
#include "acmmp_kernels.cuh"
#include "acmmp_fusion.cuh"

using namespace acmmp;

static_assert(sizeof(acmmp_camera) == 120, "acmmp_camera must mirror the reference Camera (main.h:40-54)");
static_assert(sizeof(acmmp_params) == 68, "acmmp_params must mirror PatchMatchParams (ACMMP.h:32-55)");

// ---------------------------------------------------------------------------------------------
// XORWOW sequence skipping on the host: state(y) = T^y state(0), T = (one step)^(2^67) over GF(2)
// (what curand_init's _skipahead_sequence does with precalc_xorwow_matrix, curand_kernel.h)
// ---------------------------------------------------------------------------------------------
namespace {

struct Bits160 {
    uint32_t w[5];
};

struct XorwowSkip {
    Bits160 rows[160];     // rows[r] = image of basis vector r
};

inline Bits160 xorwow_step(const Bits160 &s)
{
    const uint32_t t = s.w[0] ^ (s.w[0] >> 2);
    Bits160 o;
    o.w[0] = s.w[1]; o.w[1] = s.w[2]; o.w[2] = s.w[3]; o.w[3] = s.w[4];
    o.w[4] = (s.w[4] ^ (s.w[4] << 4)) ^ (t ^ (t << 1));
    return o;
}

inline Bits160 mat_apply(const XorwowSkip &m, const Bits160 &v)
{
    Bits160 r = {{0, 0, 0, 0, 0}};
    for (int i = 0; i < 5; ++i) {
        uint32_t bits = v.w[i];
        while (bits) {
            const int j = __builtin_ctz(bits);
            bits &= bits - 1;
            const Bits160 &row = m.rows[i * 32 + j];
            for (int k = 0; k < 5; ++k) r.w[k] ^= row.w[k];
        }
    }
    return r;
}

const XorwowSkip &xorwow_sequence_matrix()
{
    static XorwowSkip T;
    static std::once_flag once;
    std::call_once(once, [] {
        XorwowSkip m;
        for (int r = 0; r < 160; ++r) {
            Bits160 e = {{0, 0, 0, 0, 0}};
            e.w[r / 32] = 1u << (r % 32);
            m.rows[r] = xorwow_step(e);
        }
        for (int s = 0; s < 67; ++s) {       // square 67 times: one step -> 2^67 steps
            XorwowSkip sq;
            for (int r = 0; r < 160; ++r) sq.rows[r] = mat_apply(m, m.rows[r]);
            m = sq;
        }
        T = m;
    });
    return T;
}

// {d, v0..v4} of curand_init(seed, subsequence = y, offset = 0) for y = 0..H-1
void xorwow_row_states(uint64_t seed, int H, std::vector<uint32_t> &out)
{
    const uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    const uint32_t d = 6615241u + t1 + t0;
    Bits160 v;
    v.w[0] = 123456789u + t0;
    v.w[1] = 362436069u ^ t0;
    v.w[2] = 521288629u + t1;
    v.w[3] = 88675123u ^ t1;
    v.w[4] = 5783321u + t0;
    const XorwowSkip &T = xorwow_sequence_matrix();
    out.resize((size_t)6 * H);
    for (int y = 0; y < H; ++y) {
        out[6 * y + 0] = d;       // the sequence skip leaves d untouched (multiple of 2^32 steps)
        for (int k = 0; k < 5; ++k) out[6 * y + 1 + k] = v.w[k];
        v = mat_apply(T, v);
    }
}

// ---------------------------------------------------------------------------------------------
// 3x3 helpers, double precision
// ---------------------------------------------------------------------------------------------
struct M3 {
    double m[9];
};
inline M3 mul(const M3 &a, const M3 &b)
{
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return r;
}
inline M3 transpose(const M3 &a)
{
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * j + i];
    return r;
}
inline void mulv(const M3 &a, const double v[3], double o[3])
{
    for (int i = 0; i < 3; ++i) o[i] = a.m[3 * i] * v[0] + a.m[3 * i + 1] * v[1] + a.m[3 * i + 2] * v[2];
}
inline M3 from_f(const float f[9])
{
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = f[i];
    return r;
}
// projection matrix the reference applies in ProjectonCamera_cu (ACMMP.cu:634-642): rows
// (K0 K1 K2), (K3 K4 K5), (0 0 1) -- the third row of K is never read.
inline M3 kproj(const float K[9])
{
    M3 r;
    r.m[0] = K[0]; r.m[1] = K[1]; r.m[2] = K[2];
    r.m[3] = K[3]; r.m[4] = K[4]; r.m[5] = K[5];
    r.m[6] = 0; r.m[7] = 0; r.m[8] = 1;
    return r;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// ACMMP_TRACE=1: host wall clock per entry point / phase / allocator call, summed over the process and printed to
// stderr at exit (a development aid for the driver's end-to-end time; off by default, two clock reads per scope)
// ---------------------------------------------------------------------------------------------
struct TraceTable {
    std::mutex m;
    std::map<std::string, std::pair<double, long>> acc;
    const bool on = std::getenv("ACMMP_TRACE") != nullptr;
    void add(const std::string &k, double dt)
    {
        std::lock_guard<std::mutex> lock(m);
        auto &e = acc[k];
        e.first += dt;
        e.second++;
    }
    ~TraceTable()
    {
        if (!on) return;
        for (auto &kv : acc) std::fprintf(stderr, "[acmmp trace] %-44s %9.3f ms  %6ld calls\n", kv.first.c_str(), kv.second.first * 1e3, kv.second.second);
    }
};
static TraceTable g_trace;
static inline double trace_now()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
struct Trace {
    const char *name;
    double t0, tp;
    explicit Trace(const char *n) : name(n), t0(g_trace.on ? trace_now() : 0.0), tp(t0) {}
    void mark(const char *phase)
    {
        if (!g_trace.on) return;
        const double t = trace_now();
        g_trace.add(std::string(name) + ":" + phase, t - tp);
        tp = t;
    }
    ~Trace()
    {
        if (g_trace.on) g_trace.add(name, trace_now() - t0);
    }
};

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
// Size-keyed free lists for device memory, pinned host memory and CUDA arrays, in two tiers.  Every context has its
// own (unsynchronised) pool: a view walks through the same three pyramid-level shapes over and over, and a block freed
// by a context is handed back to that context in stream order.  Behind it sits one pool per DEVICE, shared by the
// contexts of that device (mutex): a context that parks (acmmp_park) or is destroyed synchronises its stream and moves
// its free blocks there, and a context that misses in its own lists looks there before it calls cudaMalloc /
// cudaMallocHost / cudaMallocArray.  A resident scene keeps one context per view alive (the stage state lives in it);
// with every context allocating its own ~1.1 GB of scratch per level, the allocator calls were 6 of 16 s of an
// 11-view 3200x2130 scene (ACMMP_TRACE).  (The reference allocates and frees everything per ProcessProblem call,
// ACMMP.cpp:685-724 / :101-143.)
struct MemPool {
    std::multimap<size_t, void *> dev_free, host_free;
    std::unordered_map<void *, size_t> dev_size, host_size;      // every block this pool owns, free or handed out
    typedef std::tuple<int, int, int> ArrKey;
    std::multimap<ArrKey, cudaArray_t> arr_free;
    std::map<cudaArray_t, ArrKey> arr_key;
    struct DeviceShared *shared = nullptr;                        // second tier (null for the shared pool itself)
    const struct DeviceShared *home = nullptr;                    // the device this pool belongs to (arena range)

    bool take_dev(size_t bytes, void **p)
    {
        auto it = dev_free.find(bytes);
        if (it == dev_free.end()) return false;
        *p = it->second;
        dev_free.erase(it);
        return true;
    }
    bool take_host(size_t bytes, void **p)
    {
        auto it = host_free.find(bytes);
        if (it == host_free.end()) return false;
        *p = it->second;
        host_free.erase(it);
        return true;
    }
    bool take_arr(const ArrKey &k, cudaArray_t *a)
    {
        auto it = arr_free.find(k);
        if (it == arr_free.end()) return false;
        *a = it->second;
        arr_free.erase(it);
        return true;
    }
    cudaError_t dmalloc(void **p, size_t bytes);
    cudaError_t hmalloc(void **p, size_t bytes);
    cudaError_t amalloc(cudaArray_t *a, const cudaChannelFormatDesc *desc, int w, int h, int layers);
    void dfree(void *p)
    {
        if (!p) return;
        auto it = dev_size.find(p);
        if (it == dev_size.end()) { if (!in_arena(p)) cudaFree(p); return; }
        dev_free.insert(std::make_pair(it->second, p));
    }
    void hfree(void *p)
    {
        if (!p) return;
        auto it = host_size.find(p);
        if (it == host_size.end()) { cudaFreeHost(p); return; }
        host_free.insert(std::make_pair(it->second, p));
    }
    void afree(cudaArray_t a)
    {
        if (!a) return;
        auto it = arr_key.find(a);
        if (it == arr_key.end()) { cudaFreeArray(a); return; }
        arr_free.insert(std::make_pair(it->second, a));
    }
    // hand one block (free or not) over to another pool's registry
    void disown_dev(void *p, MemPool &to)
    {
        auto it = dev_size.find(p);
        if (it == dev_size.end()) return;
        to.dev_size[p] = it->second;
        dev_size.erase(it);
    }
    // every FREE block changes owner (the caller has synchronised the stream that last used them)
    void give_free_to(MemPool &to)
    {
        for (auto &kv : dev_free) { to.dev_size[kv.second] = kv.first; to.dev_free.insert(kv); dev_size.erase(kv.second); }
        for (auto &kv : host_free) { to.host_size[kv.second] = kv.first; to.host_free.insert(kv); host_size.erase(kv.second); }
        for (auto &kv : arr_free) { to.arr_key[kv.second] = kv.first; to.arr_free.insert(kv); arr_key.erase(kv.second); }
        dev_free.clear(); host_free.clear(); arr_free.clear();
    }
    bool in_arena(const void *p) const;
    void release_all()
    {
        for (auto &kv : dev_size)
            if (!in_arena(kv.first)) cudaFree(kv.first);
        for (auto &kv : host_size) cudaFreeHost(kv.first);
        for (auto &kv : arr_key) cudaFreeArray(kv.first);
        dev_free.clear(); host_free.clear(); dev_size.clear(); host_size.clear(); arr_free.clear(); arr_key.clear();
    }
};

// per device: the shared pool, the curand_init states per (seed, W, H) (read-only once built; 164 MB at 3200x2130, one
// copy per device instead of one per view) and the number of live contexts (the last one out frees everything)
struct DeviceShared {
    std::mutex m;
    MemPool pool;
    std::map<std::tuple<uint64_t, int, int>, uint2 *> seeded_cache;
    int live = 0;                 // contexts + blocks handed out by acmmp_pool_alloc
    int cc_major = 0, cc_minor = 0, sm_count = 0;      // asked once (acmmp_create)
    // acmmp_reserve_device_memory: ONE cudaMalloc that new blocks are carved out of (bump pointer, 512-byte steps; a freed
    // block goes to the size-keyed free lists like any other).  cudaMalloc on a device that already holds hundreds of
    // allocations costs 1.5 - 12 ms per call (ACMMP_TRACE), and a resident scene needs three state blocks per view and level.
    char *arena = nullptr;
    size_t arena_size = 0, arena_used = 0;
    bool carve(size_t bytes, void **p)          // caller holds m
    {
        const size_t step = (bytes + 511) & ~(size_t)511;
        if (!arena || arena_used + step > arena_size) return false;
        *p = arena + arena_used;
        arena_used += step;
        return true;
    }
    void release_everything()                   // caller holds m
    {
        pool.release_all();
        seeded_cache.clear();
        if (arena) cudaFree(arena);
        arena = nullptr;
        arena_size = arena_used = 0;
    }
};
bool MemPool::in_arena(const void *p) const
{
    return home && home->arena && (const char *)p >= home->arena && (const char *)p < home->arena + home->arena_size;
}
static DeviceShared &device_shared(int device)
{
    static std::mutex m;
    static std::map<int, DeviceShared> table;          // node-based: references stay valid
    std::lock_guard<std::mutex> lock(m);
    DeviceShared &sh = table[device];
    sh.pool.home = &sh;
    return sh;
}

cudaError_t MemPool::dmalloc(void **p, size_t bytes)
{
    if (take_dev(bytes, p)) return cudaSuccess;
    if (shared) {
        std::lock_guard<std::mutex> lock(shared->m);
        if (shared->pool.take_dev(bytes, p)) { shared->pool.dev_size.erase(*p); dev_size[*p] = bytes; return cudaSuccess; }
        if (shared->carve(bytes, p)) { dev_size[*p] = bytes; return cudaSuccess; }
    }
    Trace tr("pool.cudaMalloc");
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) dev_size[*p] = bytes;
    return e;
}
cudaError_t MemPool::hmalloc(void **p, size_t bytes)
{
    if (take_host(bytes, p)) return cudaSuccess;
    if (shared) {
        std::lock_guard<std::mutex> lock(shared->m);
        if (shared->pool.take_host(bytes, p)) { shared->pool.host_size.erase(*p); host_size[*p] = bytes; return cudaSuccess; }
    }
    Trace tr("pool.cudaMallocHost");
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e == cudaSuccess) host_size[*p] = bytes;
    return e;
}
cudaError_t MemPool::amalloc(cudaArray_t *a, const cudaChannelFormatDesc *desc, int w, int h, int layers)
{
    const ArrKey k(w, h, layers);
    if (take_arr(k, a)) return cudaSuccess;
    if (shared) {
        std::lock_guard<std::mutex> lock(shared->m);
        if (shared->pool.take_arr(k, a)) { shared->pool.arr_key.erase(*a); arr_key[*a] = k; return cudaSuccess; }
    }
    Trace tr("pool.cudaMallocArray");
    cudaError_t e = layers > 0 ? cudaMalloc3DArray(a, desc, make_cudaExtent(w, h, layers), cudaArrayLayered)
                               : cudaMallocArray(a, desc, w, h);
    if (e == cudaSuccess) arr_key[*a] = k;
    return e;
}

struct acmmp_ctx {
    MemPool pool;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    acmmp_params params;
    int as_compiled = 1;
    int use_tma = 1;
    float tap_prune = 5.9604645e-08f;     // 2^-24, see acmmp_set_sphere_tap_pruning
    int num_sms = 148;
    uint64_t seed = 0;
    bool have_seeded = false;
    bool parked = false;                  // acmmp_park: only the stage state is held, the scratch went to the device pool
    uint64_t seeded_seed = 0;
    int seeded_w = 0, seeded_h = 0;

    int n = 0;                      // images (1 + sources)
    int W = 0, H = 0;
    std::vector<acmmp_camera> cams;
    std::vector<int> widths, heights;
    cudaArray_t src_array = nullptr;       // layered: one layer per source view
    cudaTextureObject_t src_tex = 0;
    std::vector<cudaArray_t> view_arrays;          // the same images, one 2-D array + texture per view
    std::vector<cudaTextureObject_t> view_tex;
    int src_w = 0, src_h = 0;              // layer size (largest source view)
    bool manual_clamp = false;             // source views differ in size -> smaller ones are edge-padded
    float *ref_dense = nullptr;     // W*H
    float *ref_padded = nullptr;
    int ref_pitch = 0;
    CUtensorMap tmap_pass, tmap_tp;
    NccTable ncc;
    std::vector<float *> depth_maps;        // owned copies (host-upload variant)
    std::vector<const float *> depth_ptrs;  // what the kernels read
    std::vector<int> depth_w, depth_h;
    std::vector<bool> depth_owned;

    ViewConst *views_dev = nullptr;
    float4 *planes = nullptr, *planes_alt = nullptr;
    float *costs = nullptr, *costs_alt = nullptr, *pre_costs = nullptr;
    uint32_t *selected_views = nullptr;
    uint2 *rng = nullptr, *rng_seeded = nullptr;
    float4 *prior_planes = nullptr;
    uint32_t *plane_masks = nullptr;
    float4 *coarse_planes = nullptr;
    int scaled_cols = 0, scaled_rows = 0;

    float4 *planes_host = nullptr;  // pinned
    float *costs_host = nullptr;    // pinned
    bool have_result = false;

    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pass_events;
    int pass_events_used = 0;
    float t_init = 0.f, t_pass_sum = 0.f, t_finalize = 0.f, t_last_pass = 0.f, t_jbu = 0.f;
    int n_pass = 0;
    bool timed_init = false, timed_final = false;
    int64_t launches = 0;
};

namespace {

template <typename T> cudaError_t pmalloc(acmmp_ctx *ctx, T **p, size_t bytes) { return ctx->pool.dmalloc(reinterpret_cast<void **>(p), bytes); }
template <typename T> cudaError_t phmalloc(acmmp_ctx *ctx, T **p, size_t bytes) { return ctx->pool.hmalloc(reinterpret_cast<void **>(p), bytes); }

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return ACMMP_E_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

int fail(acmmp_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

// Pooled temporaries of one call: returned to the pool on every exit path (CK returns early).
struct PoolTemps {
    acmmp_ctx *ctx;
    std::vector<void *> dev, host;
    explicit PoolTemps(acmmp_ctx *c) : ctx(c) {}
    ~PoolTemps()
    {
        for (void *p : dev) ctx->pool.dfree(p);
        for (void *p : host) ctx->pool.hfree(p);
    }
    template <typename T> cudaError_t d(T **p, size_t bytes)
    {
        cudaError_t e = pmalloc(ctx, p, bytes);
        if (e == cudaSuccess) dev.push_back(*p);
        return e;
    }
    template <typename T> cudaError_t h(T **p, size_t bytes)
    {
        cudaError_t e = phmalloc(ctx, p, bytes);
        if (e == cudaSuccess) host.push_back(*p);
        return e;
    }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int make_tmap(acmmp_ctx *ctx, CUtensorMap *tm, int box_w, int box_h)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ctx, ACMMP_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)ctx->ref_pitch, (cuuint64_t)(ctx->H + 2 * kRefPad)};
    const cuuint64_t strides[1] = {(cuuint64_t)ctx->ref_pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ctx->ref_padded, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, ACMMP_E_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return ACMMP_OK;
}

// everything a stage needs besides its state: textures, padded reference, ping-pong and per-stage buffers, RNG states
void free_scratch(acmmp_ctx *ctx, bool host_result)
{
    if (ctx->src_tex) cudaDestroyTextureObject(ctx->src_tex);
    if (ctx->src_array) ctx->pool.afree(ctx->src_array);
    for (cudaTextureObject_t t : ctx->view_tex) cudaDestroyTextureObject(t);
    for (cudaArray_t a : ctx->view_arrays) ctx->pool.afree(a);
    ctx->view_tex.clear();
    ctx->view_arrays.clear();
    ctx->src_tex = 0;
    ctx->src_array = nullptr;
    ctx->pool.dfree(ctx->ref_dense); ctx->ref_dense = nullptr;
    ctx->pool.dfree(ctx->ref_padded); ctx->ref_padded = nullptr;
    ctx->pool.dfree(ctx->views_dev); ctx->views_dev = nullptr;
    ctx->pool.dfree(ctx->planes_alt); ctx->pool.dfree(ctx->costs_alt);
    ctx->pool.dfree(ctx->selected_views); ctx->pool.dfree(ctx->rng);
    ctx->planes_alt = nullptr;
    ctx->costs_alt = nullptr;
    ctx->selected_views = nullptr;
    ctx->rng = ctx->rng_seeded = nullptr;
    ctx->have_seeded = false;
    if (host_result) {
        if (ctx->planes_host) ctx->pool.hfree(ctx->planes_host);
        if (ctx->costs_host) ctx->pool.hfree(ctx->costs_host);
        ctx->planes_host = nullptr; ctx->costs_host = nullptr;
        ctx->have_result = false;
    }
}

void free_prior(acmmp_ctx *ctx)
{
    ctx->pool.dfree(ctx->prior_planes); ctx->pool.dfree(ctx->plane_masks);
    ctx->prior_planes = nullptr; ctx->plane_masks = nullptr;
}

void free_views(acmmp_ctx *ctx)
{
    free_scratch(ctx, true);
    free_prior(ctx);
    // the stage state: planes + costs, and what a stage leaves for the NEXT stage of the same reference object -- the
    // costs of the uploaded planes that the hierarchy's acceptance test reads (pre_costs_cuda: written by the photometric
    // stage's initialisation, read by its passes AND by the prior stage's, ACMMP.cu:770-771, :1318-1322), the coarse planes
    ctx->pool.dfree(ctx->planes); ctx->pool.dfree(ctx->costs); ctx->pool.dfree(ctx->pre_costs); ctx->pool.dfree(ctx->coarse_planes);
    ctx->planes = nullptr;
    ctx->costs = ctx->pre_costs = nullptr;
    ctx->coarse_planes = nullptr;
    ctx->parked = false;
}

void free_depths(acmmp_ctx *ctx)
{
    for (size_t i = 0; i < ctx->depth_maps.size(); ++i)
        if (ctx->depth_owned[i]) ctx->pool.dfree(ctx->depth_maps[i]);
    ctx->depth_maps.clear();
    ctx->depth_ptrs.clear();
    ctx->depth_w.clear();
    ctx->depth_h.clear();
    ctx->depth_owned.clear();
}

// Fold (reference camera, source camera) into a ViewConst (see acmmp_types.cuh).
ViewConst fold_view(const acmmp_camera &r, const acmmp_camera &s)
{
    ViewConst c;
    std::memset(&c, 0, sizeof(c));
    const M3 Rr = from_f(r.R), Rs = from_f(s.R);
    const M3 Rrel = mul(Rs, transpose(Rr));          // ref cam -> src cam
    const M3 Rinv = transpose(Rrel);                 // src cam -> ref cam
    const double tr[3] = {r.t[0], r.t[1], r.t[2]}, ts[3] = {s.t[0], s.t[1], s.t[2]};
    double tmp[3], trel[3], tinv[3];
    mulv(Rrel, tr, tmp);
    for (int i = 0; i < 3; ++i) trel[i] = ts[i] - tmp[i];
    mulv(Rinv, ts, tmp);
    for (int i = 0; i < 3; ++i) tinv[i] = tr[i] - tmp[i];

    for (int i = 0; i < 9; ++i) { c.R[i] = (float)Rrel.m[i]; c.Ri[i] = (float)Rinv.m[i]; }
    for (int i = 0; i < 3; ++i) { c.t[i] = (float)trel[i]; c.ti[i] = (float)tinv[i]; }
    c.Wf = (float)s.width;
    c.Hf = (float)s.height;

    if (r.model == ACMMP_MODEL_PINHOLE) {
        const M3 M = mul(kproj(s.K), Rrel);
        double b[3];
        mulv(kproj(s.K), trel, b);
        const double ifx = 1.0 / r.K[0], ify = 1.0 / r.K[4];
        for (int i = 0; i < 3; ++i) {
            const double mx = M.m[3 * i + 0] * ifx, my = M.m[3 * i + 1] * ify, mz = M.m[3 * i + 2];
            c.Mx[i] = (float)mx; c.My[i] = (float)my; c.Mz[i] = (float)mz; c.b[i] = (float)b[i];
        }
        for (int i = 0; i < 3; ++i) {      // +0.5 texel-centre offset folded into rows 0 and 1
            const double h = (i < 2) ? 0.5 : 0.0;
            c.Fx[i] = (float)((M.m[3 * i + 0] + h * M.m[6]) * ifx);
            c.Fy[i] = (float)((M.m[3 * i + 1] + h * M.m[7]) * ify);
            c.Fz[i] = (float)(M.m[3 * i + 2] + h * M.m[8]);
            c.fb[i] = (float)(b[i] + h * b[2]);
        }
        const M3 Mi = mul(kproj(r.K), Rinv);
        double ib[3];
        mulv(kproj(r.K), tinv, ib);
        const double ifxs = 1.0 / s.K[0], ifys = 1.0 / s.K[4];
        for (int i = 0; i < 3; ++i) {
            c.Ix[i] = (float)(Mi.m[3 * i + 0] * ifxs);
            c.Iy[i] = (float)(Mi.m[3 * i + 1] * ifys);
            c.Iz[i] = (float)Mi.m[3 * i + 2];
            c.ib[i] = (float)ib[i];
        }
        c.cx = s.K[2];
        c.cy = s.K[5];
    } else {
        c.cx = s.params[1];
        c.cy = s.params[2];
    }
    return c;
}

int upload_view_consts(acmmp_ctx *ctx)
{
    const int nsrc = ctx->n - 1;
    std::vector<ViewConst> vc(nsrc > 0 ? nsrc : 1);
    for (int i = 0; i < nsrc; ++i) {
        vc[i] = fold_view(ctx->cams[0], ctx->cams[i + 1]);
        vc[i].tex = 0;
        if ((int)ctx->depth_ptrs.size() > i + 1) {
            vc[i].depth = ctx->depth_ptrs[i + 1];
            vc[i].dW = ctx->depth_w[i + 1];
            vc[i].dH = ctx->depth_h[i + 1];
        }
    }
    if (!ctx->views_dev) CK(pmalloc(ctx, &ctx->views_dev, sizeof(ViewConst) * kMaxSrc));
    CK(cudaMemcpyAsync(ctx->views_dev, vc.data(), sizeof(ViewConst) * (size_t)(nsrc > 0 ? nsrc : 1), cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));       // vc is a stack-lifetime staging buffer
    return ACMMP_OK;
}

FrameConst frame_const(const acmmp_ctx *ctx)
{
    FrameConst fc;
    std::memset(&fc, 0, sizeof(fc));
    const acmmp_camera &r = ctx->cams[0];
    fc.W = ctx->W; fc.H = ctx->H;
    fc.model = r.model;
    fc.nsrc = ctx->n - 1;
    if (r.model == ACMMP_MODEL_PINHOLE) {
        fc.cx = r.K[2]; fc.cy = r.K[5];
        fc.ifx = (float)(1.0 / r.K[0]); fc.ify = (float)(1.0 / r.K[4]);
    } else {
        fc.cx = r.params[1]; fc.cy = r.params[2];
    }
    fc.Wf = (float)ctx->W; fc.Hf = (float)ctx->H;
    for (int i = 0; i < 9; ++i) fc.R[i] = r.R[i];
    fc.depth_min = ctx->params.depth_min;
    fc.depth_max = ctx->params.depth_max;
    fc.geom = ctx->params.geom_consistency;
    fc.prior = ctx->params.planar_prior;
    fc.hierarchy = ctx->params.hierarchy;
    fc.upsample = ctx->params.upsample;
    fc.scaled_cols = ctx->scaled_cols;
    fc.scaled_rows = ctx->scaled_rows;
    fc.as_compiled = ctx->as_compiled;
    fc.ref_pitch = ctx->ref_pitch;
    fc.use_tma = ctx->use_tma;
    fc.tap_prune = (ctx->cams[0].model == ACMMP_MODEL_SPHERE) ? ctx->tap_prune : 0.0f;
    fc.tex_src = (unsigned long long)ctx->src_tex;
    fc.ref_padded = ctx->ref_padded;
    fc.views = ctx->views_dev;
    fc.planes = ctx->planes; fc.planes_alt = ctx->planes_alt;
    fc.costs = ctx->costs; fc.costs_alt = ctx->costs_alt;
    fc.pre_costs = ctx->pre_costs;
    fc.selected_views = ctx->selected_views;
    fc.rng = ctx->rng; fc.rng_seeded = ctx->rng_seeded;
    fc.prior_planes = ctx->prior_planes;
    fc.plane_masks = ctx->plane_masks;
    fc.coarse_planes = ctx->coarse_planes;
    return fc;
}

// the per-view NCC constants + texture handles that travel as a kernel parameter
NccTable ncc_table(const acmmp_ctx *ctx)
{
    NccTable nt;
    std::memset(&nt, 0, sizeof(nt));
    for (int i = 0; i + 1 < ctx->n && i < kMaxSrc; ++i) {
        const ViewConst c = fold_view(ctx->cams[0], ctx->cams[i + 1]);
        float *a = nt.c[i].a;
        if (ctx->cams[0].model == ACMMP_MODEL_PINHOLE) {
            for (int k = 0; k < 3; ++k) { a[k] = c.Fx[k]; a[3 + k] = c.Fy[k]; a[6 + k] = c.Fz[k]; a[9 + k] = c.fb[k]; }
            a[12] = c.Wf + 0.5f;
            a[13] = c.Hf + 0.5f;
        } else {
            for (int k = 0; k < 9; ++k) a[k] = c.R[k];
            for (int k = 0; k < 3; ++k) a[9 + k] = c.t[k];
            a[12] = c.cx; a[13] = c.cy; a[14] = c.Wf; a[15] = c.Hf;
        }
        nt.tex[i] = (unsigned long long)ctx->view_tex[i];
    }
    return nt;
}

template <int MODEL> size_t smem_quad(int nsrc) { return SmemLayout<MODEL, kTpTW, kTpTH, kPqWRS, kPqNT>(nsrc, kPqPix, 0, kTqPerHyp).total; }
template <int MODEL> size_t smem_pass(int nsrc) { return SmemLayout<MODEL, kPassTW, kPassTH, kPassWRS, kPassNT>(nsrc, kPassNT, kPassPix, kPassTq, 2).total; }

int configure_kernels(acmmp_ctx *ctx)
{
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: every device a context is created on
    // needs its own opt-in (one process may drive several GPUs, host/acmmp_main.cpp --gpus N).
    static std::mutex mu;
    static std::map<int, cudaError_t> done;
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find(ctx->device);
    if (it == done.end()) {
        cudaError_t result = cudaSetDevice(ctx->device);
        const int big = 227 * 1024;      // the per-CTA maximum of sm_100
        cudaError_t e;
#define SETATTR(k) if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) result = e;
        SETATTR((k_pass<kModelPinhole, kModePhoto>)) SETATTR((k_pass<kModelPinhole, kModePrior>)) SETATTR((k_pass<kModelPinhole, kModeGeom>))
        SETATTR((k_pass<kModelSphere, kModePhoto>)) SETATTR((k_pass<kModelSphere, kModePrior>)) SETATTR((k_pass<kModelSphere, kModeGeom>))
        SETATTR(k_random_init<kModelPinhole>) SETATTR(k_random_init<kModelSphere>)
        SETATTR(k_probe_quad<kModelPinhole>) SETATTR(k_probe_quad<kModelSphere>)
#undef SETATTR
        if (result != cudaSuccess) (void)cudaGetLastError();      // not sticky; a later context on this device retries
        else done[ctx->device] = result;
        if (result != cudaSuccess) return fail(ctx, ACMMP_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(result));
    }
    return ACMMP_OK;
}

int ensure_seeded(acmmp_ctx *ctx)
{
    if (ctx->have_seeded && ctx->seeded_seed == ctx->seed && ctx->seeded_w == ctx->W && ctx->seeded_h == ctx->H) return ACMMP_OK;
    const auto key = std::make_tuple(ctx->seed, ctx->W, ctx->H);
    DeviceShared &sh = *ctx->pool.shared;
    uint2 *cached = nullptr;
    {
        std::lock_guard<std::mutex> lock(sh.m);
        auto it = sh.seeded_cache.find(key);
        if (it != sh.seeded_cache.end()) cached = it->second;
    }
    if (!cached) {
        uint2 *buf = nullptr;
        CK(pmalloc(ctx, &buf, sizeof(uint2) * 3 * (size_t)ctx->W * ctx->H));
        std::vector<uint32_t> rows;
        xorwow_row_states(ctx->seed, ctx->H, rows);
        uint32_t *rows_dev = nullptr;
        CK(pmalloc(ctx, &rows_dev, rows.size() * sizeof(uint32_t)));
        CK(cudaMemcpyAsync(rows_dev, rows.data(), rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        k_rng_fill<<<(ctx->H + 63) / 64, 64, 0, ctx->stream>>>(rows_dev, ctx->W, ctx->H, buf);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->pool.dfree(rows_dev);
        // complete and read-only from here on: the device's table owns it (two contexts racing to build the same key:
        // the loser's copy goes back to the pool)
        std::lock_guard<std::mutex> lock(sh.m);
        auto ins = sh.seeded_cache.insert(std::make_pair(key, buf));
        if (ins.second) ctx->pool.disown_dev(buf, sh.pool);
        else ctx->pool.dfree(buf);
        cached = ins.first->second;
    }
    ctx->rng_seeded = cached;
    ctx->have_seeded = true;
    ctx->seeded_seed = ctx->seed;
    ctx->seeded_w = ctx->W;
    ctx->seeded_h = ctx->H;
    return ACMMP_OK;
}

int check_ready(acmmp_ctx *ctx)
{
    if (!ctx) return ACMMP_E_ARG;
    if (ctx->n < 2 || !ctx->planes) return fail(ctx, ACMMP_E_ARG, "acmmp_set_views has not been called (need >= 1 source view)");
    if (ctx->parked) return fail(ctx, ACMMP_E_ARG, "the context is parked: call acmmp_set_views[_device] with the same shapes first");
    if (ctx->params.geom_consistency && (int)ctx->depth_ptrs.size() != ctx->n)
        return fail(ctx, ACMMP_E_ARG, "geom_consistency is set but acmmp_set_depth_maps was not called with one map per view");
    if (ctx->params.planar_prior && !ctx->prior_planes)
        return fail(ctx, ACMMP_E_ARG, "planar_prior is set but acmmp_set_planar_prior_inputs was not called");
    if (ctx->params.hierarchy && !ctx->coarse_planes)
        return fail(ctx, ACMMP_E_ARG, "hierarchy is set but acmmp_set_hierarchy_inputs was not called");
    return ACMMP_OK;
}

int set_views_common(acmmp_ctx *ctx, int n, const float *const *images, bool on_device, const int32_t *widths,
                     const int32_t *heights, const acmmp_camera *cams)
{
    Trace tr("set_views");
    if (!ctx || n < 2 || n > kMaxSrc + 1 || !images || !widths || !heights || !cams)
        return fail(ctx, ACMMP_E_ARG, "acmmp_set_views: need 2..33 images");
    for (int i = 0; i < n; ++i) {
        if (cams[i].model != cams[0].model) return fail(ctx, ACMMP_E_UNSUPPORTED, "mixed camera models in one problem are not supported");
        if (cams[i].model != ACMMP_MODEL_PINHOLE && cams[i].model != ACMMP_MODEL_SPHERE)
            return fail(ctx, ACMMP_E_ARG, "unknown camera model");
        if (widths[i] <= 0 || heights[i] <= 0) return fail(ctx, ACMMP_E_ARG, "bad image size");
    }
    CK(cudaSetDevice(ctx->device));
    const bool same_shape = (ctx->n == n && ctx->W == widths[0] && ctx->H == heights[0]);
    bool same_all = same_shape;
    if (same_shape)
        for (int i = 0; i < n; ++i) same_all = same_all && ctx->widths[i] == widths[i] && ctx->heights[i] == heights[i];
    // a parked context re-activated with the shapes it was parked with keeps its stage state (planes, costs, prior,
    // coarse planes) and only takes scratch again; anything else starts the view from nothing, like a new context
    const bool resume = ctx->parked && same_all && ctx->planes;
    if (!same_all || ctx->parked) {
        if (!resume) free_views(ctx);
        ctx->parked = false;
        ctx->n = n;
        ctx->W = widths[0];
        ctx->H = heights[0];
        ctx->widths.assign(widths, widths + n);
        ctx->heights.assign(heights, heights + n);
        const size_t npx = (size_t)ctx->W * ctx->H;
        const cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        ctx->src_w = ctx->src_h = 0;
        ctx->manual_clamp = false;
        for (int i = 1; i < n; ++i) {
            ctx->src_w = std::max(ctx->src_w, (int)widths[i]);
            ctx->src_h = std::max(ctx->src_h, (int)heights[i]);
            if (widths[i] != widths[1] || heights[i] != heights[1]) ctx->manual_clamp = true;
        }
        CK(ctx->pool.amalloc(&ctx->src_array, &desc, ctx->src_w, ctx->src_h, n - 1));
        {
            cudaResourceDesc res;
            std::memset(&res, 0, sizeof(res));
            res.resType = cudaResourceTypeArray;
            res.res.array.array = ctx->src_array;
            cudaTextureDesc td;
            std::memset(&td, 0, sizeof(td));
            // The reference asks for Wrap with un-normalised coordinates (ACMMP.cpp:700-704), which
            // CUDA turns into Clamp; ask for what it gets.
            td.addressMode[0] = cudaAddressModeClamp;
            td.addressMode[1] = cudaAddressModeClamp;
            td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            CK(cudaCreateTextureObject(&ctx->src_tex, &res, &td, nullptr));
            for (int i = 1; i < n; ++i) {
                cudaArray_t arr = nullptr;
                CK(ctx->pool.amalloc(&arr, &desc, widths[i], heights[i], 0));
                ctx->view_arrays.push_back(arr);
                res.res.array.array = arr;
                cudaTextureObject_t t = 0;
                CK(cudaCreateTextureObject(&t, &res, &td, nullptr));
                ctx->view_tex.push_back(t);
            }
        }
        ctx->ref_pitch = (ctx->W + 2 * kRefPad + 3) & ~3;
        CK(pmalloc(ctx, &ctx->ref_dense, sizeof(float) * npx));
        CK(pmalloc(ctx, &ctx->ref_padded, sizeof(float) * (size_t)ctx->ref_pitch * (ctx->H + 2 * kRefPad)));
        if (!resume) {
            CK(pmalloc(ctx, &ctx->planes, sizeof(float4) * npx));
            CK(pmalloc(ctx, &ctx->costs, sizeof(float) * npx));
            CK(pmalloc(ctx, &ctx->pre_costs, sizeof(float) * npx));
            CK(cudaMemsetAsync(ctx->planes, 0, sizeof(float4) * npx, ctx->stream));
            CK(cudaMemsetAsync(ctx->costs, 0, sizeof(float) * npx, ctx->stream));
            CK(cudaMemsetAsync(ctx->pre_costs, 0, sizeof(float) * npx, ctx->stream));
        }
        CK(pmalloc(ctx, &ctx->planes_alt, sizeof(float4) * npx));
        CK(pmalloc(ctx, &ctx->costs_alt, sizeof(float) * npx));
        CK(pmalloc(ctx, &ctx->selected_views, sizeof(uint32_t) * npx));
        CK(pmalloc(ctx, &ctx->rng, sizeof(uint2) * 3 * npx));
        CK(cudaMemsetAsync(ctx->selected_views, 0, sizeof(uint32_t) * npx, ctx->stream));
        // the pinned result buffers are allocated on the first download (ensure_host_result): a resident chain that
        // never brings a level's result to the host does not pay 136 MB of cudaMallocHost per level
        typedef TileGeom<kPassTW, kPassTH> TGp;
        typedef TileGeom<kTpTW, kTpTH> TGt;
        int rc = make_tmap(ctx, &ctx->tmap_pass, TGp::PW, TGp::RH);
        if (rc) return rc;
        rc = make_tmap(ctx, &ctx->tmap_tp, TGt::PW, TGt::RH);
        if (rc) return rc;
    }
    tr.mark("alloc");
    ctx->cams.assign(cams, cams + n);
    for (int i = 0; i < n; ++i) {
        ctx->cams[i].width = widths[i];
        ctx->cams[i].height = heights[i];
    }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    float *pad_dst = nullptr;
    if (ctx->manual_clamp) CK(pmalloc(ctx, &pad_dst, sizeof(float) * (size_t)ctx->src_w * ctx->src_h));
    // host images cross PCIe once, into a linear staging buffer; both texture copies are device-to-device
    float *stage = nullptr;
    if (!on_device) CK(pmalloc(ctx, &stage, sizeof(float) * (size_t)ctx->src_w * ctx->src_h));
    for (int i = 1; i < n; ++i) {
        const float *src_dev = images[i];
        if (!on_device) {
            CK(cudaMemcpyAsync(stage, images[i], sizeof(float) * (size_t)widths[i] * heights[i], cudaMemcpyHostToDevice, ctx->stream));
            src_dev = stage;
        }
        cudaMemcpy3DParms cp;
        std::memset(&cp, 0, sizeof(cp));
        cp.dstArray = ctx->src_array;
        cp.dstPos = make_cudaPos(0, 0, i - 1);
        cp.kind = cudaMemcpyDeviceToDevice;
        if (widths[i] == ctx->src_w && heights[i] == ctx->src_h) {
            cp.srcPtr = make_cudaPitchedPtr(const_cast<float *>(src_dev), sizeof(float) * widths[i], widths[i], heights[i]);
            cp.extent = make_cudaExtent(widths[i], heights[i], 1);
        } else {
            // smaller than the layer: replicate the last column / row so that hardware clamping at the
            // layer edge equals clamping at the image edge
            dim3 grid((ctx->src_w + 255) / 256, ctx->src_h);
            k_replicate_pad<<<grid, 256, 0, ctx->stream>>>(src_dev, widths[i], heights[i], pad_dst, ctx->src_w, ctx->src_h);
            ctx->launches++;
            cp.srcPtr = make_cudaPitchedPtr(pad_dst, sizeof(float) * ctx->src_w, ctx->src_w, ctx->src_h);
            cp.extent = make_cudaExtent(ctx->src_w, ctx->src_h, 1);
        }
        CK(cudaMemcpy3DAsync(&cp, ctx->stream));
        CK(cudaMemcpy2DToArrayAsync(ctx->view_arrays[i - 1], 0, 0, src_dev, sizeof(float) * widths[i], sizeof(float) * widths[i],
                                    heights[i], cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (stage) ctx->pool.dfree(stage);      // stream-ordered reuse: later users enqueue on the same stream
    if (ctx->manual_clamp) ctx->pool.dfree(pad_dst);
    CK(cudaMemcpyAsync(ctx->ref_dense, images[0], sizeof(float) * (size_t)ctx->W * ctx->H, kind, ctx->stream));
    {
        dim3 grid((ctx->ref_pitch + 255) / 256, ctx->H + 2 * kRefPad);
        k_pad_reference<<<grid, 256, 0, ctx->stream>>>(ctx->ref_dense, ctx->W, ctx->H, ctx->ref_padded, ctx->ref_pitch);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    // what InuputInitialization derives (ACMMP.cpp:645-651)
    ctx->params.depth_min = ctx->cams[0].depth_min * 0.6f;
    ctx->params.depth_max = ctx->cams[0].depth_max * 1.2f;
    ctx->params.num_images = n;
    ctx->params.disparity_min = ctx->cams[0].K[0] * ctx->params.baseline / ctx->params.depth_max;
    ctx->params.disparity_max = ctx->cams[0].K[0] * ctx->params.baseline / ctx->params.depth_min;
    ctx->ncc = ncc_table(ctx);
    tr.mark("enqueue");
    int rc = upload_view_consts(ctx);
    if (rc) return rc;
    rc = configure_kernels(ctx);
    tr.mark("consts");
    return rc;
}

int set_depths_common(acmmp_ctx *ctx, int n, const float *const *maps, bool on_device, const int32_t *widths,
                      const int32_t *heights)
{
    Trace tr("set_depths");
    if (!ctx || n != ctx->n || !maps || !widths || !heights)
        return fail(ctx, ACMMP_E_ARG, "acmmp_set_depth_maps: one map per view (after acmmp_set_views)");
    // A null / empty neighbour map would make geom_address clamp to texel -1 of a null base (the file-chained host
    // path produced exactly that when a depths.dmb was missing): refuse it here.
    for (int i = 0; i < n; ++i) {
        if (widths[i] <= 0 || heights[i] <= 0 || (maps[i] == nullptr && i > 0))
            return fail(ctx, ACMMP_E_ARG, "acmmp_set_depth_maps: map " + std::to_string(i) + " is null or empty");
    }
    CK(cudaSetDevice(ctx->device));
    free_depths(ctx);
    for (int i = 0; i < n; ++i) {
        ctx->depth_w.push_back(widths[i]);
        ctx->depth_h.push_back(heights[i]);
        if (i == 0 && maps[0] == nullptr) {
            // the reference view's own depth map = depth of the device-resident state
            if (widths[0] != ctx->W || heights[0] != ctx->H) return fail(ctx, ACMMP_E_ARG, "own depth map must have the view's size");
            float *d = nullptr;
            const int npx = ctx->W * ctx->H;
            CK(pmalloc(ctx, &d, sizeof(float) * (size_t)npx));
            k_export_depth<<<(npx + 255) / 256, 256, 0, ctx->stream>>>(ctx->planes, npx, d);
            ctx->launches++;
            CK(cudaGetLastError());
            ctx->depth_maps.push_back(d);
            ctx->depth_ptrs.push_back(d);
            ctx->depth_owned.push_back(true);
        } else if (on_device) {
            ctx->depth_maps.push_back(nullptr);
            ctx->depth_ptrs.push_back(maps[i]);
            ctx->depth_owned.push_back(false);
        } else {
            float *d = nullptr;
            CK(pmalloc(ctx, &d, sizeof(float) * (size_t)widths[i] * heights[i]));
            CK(cudaMemcpyAsync(d, maps[i], sizeof(float) * (size_t)widths[i] * heights[i], cudaMemcpyHostToDevice, ctx->stream));
            ctx->depth_maps.push_back(d);
            ctx->depth_ptrs.push_back(d);
            ctx->depth_owned.push_back(true);
        }
    }
    return upload_view_consts(ctx);
}

void record_begin(acmmp_ctx *ctx, int which)
{
    cudaEventRecord(ctx->ev[which], ctx->stream);
}

template <int MODEL>
int launch_init(acmmp_ctx *ctx)
{
    const FrameConst fc = frame_const(ctx);
    dim3 grid((ctx->W + kTpTW - 1) / kTpTW, (ctx->H + kTpTH - 1) / kTpTH);
    const size_t smem = smem_quad<MODEL>(fc.nsrc);
    if (smem > 227 * 1024) return fail(ctx, ACMMP_E_UNSUPPORTED, "k_random_init: too many source views for one SM's shared memory");
    k_random_init<MODEL><<<grid, kPqNT, smem, ctx->stream>>>(fc, ctx->ncc, ctx->tmap_tp);
    ctx->launches++;
    CK(cudaGetLastError());
    return ACMMP_OK;
}

template <int MODEL>
int launch_pass(acmmp_ctx *ctx, int colour, int iter)
{
    const FrameConst fc = frame_const(ctx);
    // persistent kernel: one CTA per SM walks the tile grid (row-major) with stride gridDim.x
    const int tiles_x = (ctx->W + kPassTW - 1) / kPassTW, ntiles = tiles_x * ((ctx->H + kPassTH - 1) / kPassTH);
    const dim3 grid(std::min(ntiles, ctx->num_sms * ACMMP_PASS_MIN_CTAS));
    // the reference's flag combinations are exclusive per stage; should a caller set both, geometric wins for the
    // cost terms and the prior term is dropped -- refuse instead
    if (fc.geom && fc.prior) return fail(ctx, ACMMP_E_UNSUPPORTED, "geom_consistency and planar_prior in the same stage are not supported");
    const size_t smem = smem_pass<MODEL>(fc.nsrc);
    if (smem > 227 * 1024) return fail(ctx, ACMMP_E_UNSUPPORTED, "k_pass needs more shared memory than an SM has for this many source views");
    if (fc.geom) k_pass<MODEL, kModeGeom><<<grid, kPassNT, smem, ctx->stream>>>(fc, ctx->ncc, ctx->tmap_pass, colour, iter, tiles_x, ntiles);
    else if (fc.prior) k_pass<MODEL, kModePrior><<<grid, kPassNT, smem, ctx->stream>>>(fc, ctx->ncc, ctx->tmap_pass, colour, iter, tiles_x, ntiles);
    else k_pass<MODEL, kModePhoto><<<grid, kPassNT, smem, ctx->stream>>>(fc, ctx->ncc, ctx->tmap_pass, colour, iter, tiles_x, ntiles);
    ctx->launches++;
    CK(cudaGetLastError());
    std::swap(ctx->planes, ctx->planes_alt);
    std::swap(ctx->costs, ctx->costs_alt);
    return ACMMP_OK;
}

template <int MODEL>
int launch_finalize(acmmp_ctx *ctx)
{
    const FrameConst fc = frame_const(ctx);
    const int npx = ctx->W * ctx->H;
    k_depth_normal<MODEL><<<(npx + 255) / 256, 256, 0, ctx->stream>>>(fc);
    const dim3 mgrid((ctx->W + kMedTW - 1) / kMedTW, (ctx->H + kMedTH - 1) / kMedTH);
    k_median_filter<<<mgrid, 256, 0, ctx->stream>>>(fc, 0);
    k_median_filter<<<mgrid, 256, 0, ctx->stream>>>(fc, 1);
    ctx->launches += 3;
    CK(cudaGetLastError());
    return ACMMP_OK;
}

int do_init(acmmp_ctx *ctx)
{
    int rc = check_ready(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if (ctx->params.upsample) {
        const int Imagescale = (int)std::fmax((float)ctx->W / (float)ctx->scaled_cols, (float)ctx->H / (float)ctx->scaled_rows);
        if ((Imagescale * Imagescale + 1) / 2 > kHalo) return fail(ctx, ACMMP_E_UNSUPPORTED, "upsample factor > 3 is not supported");
    }
    rc = ensure_seeded(ctx);
    if (rc) return rc;
    record_begin(ctx, 0);
    rc = (ctx->cams[0].model == ACMMP_MODEL_PINHOLE) ? launch_init<kModelPinhole>(ctx) : launch_init<kModelSphere>(ctx);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->timed_init = true;
    ctx->pass_events_used = 0;
    ctx->timed_final = false;
    return rc;
}

int do_pass(acmmp_ctx *ctx, int colour, int iter)
{
    int rc = check_ready(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((int)ctx->pass_events.size() <= ctx->pass_events_used) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        ctx->pass_events.push_back(std::make_pair(a, b));
    }
    auto &pe = ctx->pass_events[ctx->pass_events_used++];
    cudaEventRecord(pe.first, ctx->stream);
    rc = (ctx->cams[0].model == ACMMP_MODEL_PINHOLE) ? launch_pass<kModelPinhole>(ctx, colour, iter)
                                                     : launch_pass<kModelSphere>(ctx, colour, iter);
    cudaEventRecord(pe.second, ctx->stream);
    return rc;
}

int do_finalize(acmmp_ctx *ctx)
{
    int rc = check_ready(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    record_begin(ctx, 2);
    rc = (ctx->cams[0].model == ACMMP_MODEL_PINHOLE) ? launch_finalize<kModelPinhole>(ctx) : launch_finalize<kModelSphere>(ctx);
    cudaEventRecord(ctx->ev[3], ctx->stream);
    ctx->timed_final = true;
    return rc;
}

int collect_timings(acmmp_ctx *ctx)
{
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->timed_init) cudaEventElapsedTime(&ctx->t_init, ctx->ev[0], ctx->ev[1]);
    if (ctx->timed_final) cudaEventElapsedTime(&ctx->t_finalize, ctx->ev[2], ctx->ev[3]);
    ctx->t_pass_sum = 0.f;
    ctx->n_pass = ctx->pass_events_used;
    for (int i = 0; i < ctx->pass_events_used; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->pass_events[i].first, ctx->pass_events[i].second);
        ctx->t_pass_sum += ms;
        ctx->t_last_pass = ms;
    }
    return ACMMP_OK;
}

template <int MODEL>
int launch_probe(acmmp_ctx *ctx, int mode, int view, const float4 *planes_dev, float *out, float4 *out4)
{
    const FrameConst fc = frame_const(ctx);
    dim3 grid((ctx->W + kTpTW - 1) / kTpTW, (ctx->H + kTpTH - 1) / kTpTH);
    k_probe<MODEL><<<grid, 128, 0, ctx->stream>>>(fc, mode, view, planes_dev, out, out4);
    ctx->launches++;
    CK(cudaGetLastError());
    return ACMMP_OK;
}

// view >= 1: NCC against that view; view == 0: initial cost + selected views (all views)
template <int MODEL>
int launch_probe_quad(acmmp_ctx *ctx, int view, const float4 *planes_dev, float *out, uint32_t *out_views)
{
    const FrameConst fc = frame_const(ctx);
    dim3 grid((ctx->W + kTpTW - 1) / kTpTW, (ctx->H + kTpTH - 1) / kTpTH);
    const size_t smem = smem_quad<MODEL>(fc.nsrc);
    if (smem > 227 * 1024) return fail(ctx, ACMMP_E_UNSUPPORTED, "k_probe_quad: too many source views for one SM's shared memory");
    k_probe_quad<MODEL><<<grid, kPqNT, smem, ctx->stream>>>(fc, ctx->ncc, ctx->tmap_tp, view, planes_dev, out, out_views);
    ctx->launches++;
    CK(cudaGetLastError());
    return ACMMP_OK;
}

int run_probe(acmmp_ctx *ctx, int mode, int view, const float *planes4, float *out, float *out4, uint32_t *out_views)
{
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!planes4) return fail(ctx, ACMMP_E_ARG, "probe: planes is null");
    if (mode != 3 && (view < 1 || view >= ctx->n)) return fail(ctx, ACMMP_E_ARG, "probe: view must be 1..n-1");
    if (mode == 1 && (int)ctx->depth_ptrs.size() != ctx->n) return fail(ctx, ACMMP_E_ARG, "probe geom: no depth maps");
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    float4 *dp = nullptr, *do4 = nullptr;
    float *dout = nullptr;
    uint32_t *dv = nullptr;
    PoolTemps tmp(ctx);
    CK(tmp.d(&dp, sizeof(float4) * npx));
    CK(tmp.d(&do4, sizeof(float4) * npx));
    CK(tmp.d(&dout, sizeof(float) * npx));
    CK(tmp.d(&dv, sizeof(uint32_t) * npx));
    CK(cudaMemcpyAsync(dp, planes4, sizeof(float4) * npx, cudaMemcpyHostToDevice, ctx->stream));
    const bool pinhole = ctx->cams[0].model == ACMMP_MODEL_PINHOLE;
    if (mode == 0 || mode == 3) {          // the NCC forms: through quad_ncc, like every kernel of the library
        const int v = (mode == 3) ? 0 : view;
        rc = pinhole ? launch_probe_quad<kModelPinhole>(ctx, v, dp, dout, dv) : launch_probe_quad<kModelSphere>(ctx, v, dp, dout, dv);
    } else {
        rc = pinhole ? launch_probe<kModelPinhole>(ctx, mode, view, dp, dout, do4) : launch_probe<kModelSphere>(ctx, mode, view, dp, dout, do4);
    }
    if (rc == ACMMP_OK) {
        if (out) CK(cudaMemcpyAsync(out, dout, sizeof(float) * npx, cudaMemcpyDeviceToHost, ctx->stream));
        if (out4) CK(cudaMemcpyAsync(out4, do4, sizeof(float4) * npx, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_views) CK(cudaMemcpyAsync(out_views, dv, sizeof(uint32_t) * npx, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return rc;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

void acmmp_default_params(acmmp_params *p)
{
    std::memset(p, 0, sizeof(*p));
    p->max_iterations = 3; p->patch_size = 11; p->num_images = 5; p->max_image_size = 3200;
    p->radius_increment = 2; p->sigma_spatial = 5.0f; p->sigma_color = 3.0f; p->top_k = 4;
    p->baseline = 0.54f; p->depth_min = 0.0f; p->depth_max = 1.0f; p->disparity_min = 0.0f; p->disparity_max = 1.0f;
}

const char *acmmp_version(void) { return "acmmp_b200 0.1 (sm_100a)"; }
int acmmp_abi_sizeof_camera(void) { return (int)sizeof(acmmp_camera); }
int acmmp_abi_sizeof_params(void) { return (int)sizeof(acmmp_params); }

int acmmp_create(acmmp_ctx **out, int device)
{
    Trace tr("create");
    if (!out) return ACMMP_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ACMMP_E_CUDA;
    DeviceShared &dsh = device_shared(device);
    int cc_major = 0, cc_minor = 0, sm_count = 0;
    {
        // asked once per device: cudaGetDeviceProperties takes 5 - 40 ms, and a resident scene creates a context per view
        std::lock_guard<std::mutex> lock(dsh.m);
        if (dsh.sm_count == 0) {
            if (cudaDeviceGetAttribute(&dsh.cc_major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
                cudaDeviceGetAttribute(&dsh.cc_minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess ||
                cudaDeviceGetAttribute(&dsh.sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
                dsh.sm_count = 0;
                return ACMMP_E_CUDA;
            }
        }
        cc_major = dsh.cc_major; cc_minor = dsh.cc_minor; sm_count = dsh.sm_count;
    }
    if (cc_major != 10) {
        std::fprintf(stderr, "acmmp_b200: device %d is sm_%d%d; this library only carries sm_100a code\n", device, cc_major, cc_minor);
        return ACMMP_E_UNSUPPORTED;
    }
    if (cudaSetDevice(device) != cudaSuccess) return ACMMP_E_CUDA;
    acmmp_ctx *ctx = new acmmp_ctx();
    ctx->device = device;
    ctx->pool.shared = &dsh;
    ctx->pool.home = ctx->pool.shared;
    {
        std::lock_guard<std::mutex> lock(ctx->pool.shared->m);
        ctx->pool.shared->live++;
    }
    ctx->num_sms = sm_count;
    acmmp_default_params(&ctx->params);
    if (const char *e = std::getenv("ACMMP_NO_TMA")) ctx->use_tma = (e[0] == '1') ? 0 : 1;   // debug aid
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        {
            std::lock_guard<std::mutex> lock(ctx->pool.shared->m);
            ctx->pool.shared->live--;
        }
        delete ctx;
        return ACMMP_E_CUDA;
    }
    for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
    *out = ctx;
    return ACMMP_OK;
}

int acmmp_destroy(acmmp_ctx *ctx)
{
    if (!ctx) return ACMMP_E_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_views(ctx);
    free_depths(ctx);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ctx->ev[i]);
    for (auto &pe : ctx->pass_events) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    {
        // the stream is idle: the free blocks go to the device's pool for the other contexts; the last context of a
        // device returns everything to the driver
        DeviceShared &sh = *ctx->pool.shared;
        std::lock_guard<std::mutex> lock(sh.m);
        ctx->pool.give_free_to(sh.pool);
        if (--sh.live == 0) sh.release_everything();
    }
    ctx->pool.release_all();               // anything still registered here was not returned by its owner
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ACMMP_OK;
}

const char *acmmp_last_error(const acmmp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int acmmp_set_views(acmmp_ctx *ctx, int n, const float *const *images, const int32_t *widths, const int32_t *heights,
                    const acmmp_camera *cams)
{
    return set_views_common(ctx, n, images, false, widths, heights, cams);
}

int acmmp_set_views_device(acmmp_ctx *ctx, int n, const float *const *images_dev, const int32_t *widths,
                           const int32_t *heights, const acmmp_camera *cams)
{
    return set_views_common(ctx, n, images_dev, true, widths, heights, cams);
}

int acmmp_set_geom_consistency(acmmp_ctx *ctx, int multi_geometry)
{
    if (!ctx) return ACMMP_E_ARG;
    ctx->params.geom_consistency = 1;             // ACMMP.cpp:548-555
    ctx->params.max_iterations = 2;
    if (multi_geometry) ctx->params.multi_geometry = 1;
    return ACMMP_OK;
}

int acmmp_set_hierarchy(acmmp_ctx *ctx)
{
    if (!ctx) return ACMMP_E_ARG;
    ctx->params.hierarchy = 1;                    // ACMMP.cpp:557-560
    return ACMMP_OK;
}

int acmmp_set_planar_prior(acmmp_ctx *ctx)
{
    if (!ctx) return ACMMP_E_ARG;
    ctx->params.planar_prior = 1;                 // ACMMP.cpp:562-565
    return ACMMP_OK;
}

int acmmp_set_max_iterations(acmmp_ctx *ctx, int n)
{
    if (!ctx || n < 0) return ACMMP_E_ARG;
    ctx->params.max_iterations = n;
    return ACMMP_OK;
}

int acmmp_get_params(const acmmp_ctx *ctx, acmmp_params *out)
{
    if (!ctx || !out) return ACMMP_E_ARG;
    *out = ctx->params;
    return ACMMP_OK;
}

int acmmp_reset_modes(acmmp_ctx *ctx)
{
    if (!ctx) return ACMMP_E_ARG;
    ctx->params.geom_consistency = 0;
    ctx->params.multi_geometry = 0;
    ctx->params.planar_prior = 0;
    ctx->params.hierarchy = 0;
    ctx->params.upsample = 0;
    ctx->params.max_iterations = 3;
    return ACMMP_OK;
}

int acmmp_park(acmmp_ctx *ctx, int keep_prior, int keep_host_result)
{
    if (!ctx) return ACMMP_E_ARG;
    if (!ctx->planes) return fail(ctx, ACMMP_E_ARG, "acmmp_park: no views on this context");
    Trace tr("park");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    free_scratch(ctx, !keep_host_result);
    if (!keep_prior) {
        free_prior(ctx);
        ctx->params.planar_prior = 0;
    }
    free_depths(ctx);
    ctx->params.geom_consistency = ctx->params.multi_geometry = 0;
    ctx->parked = true;
    {
        std::lock_guard<std::mutex> lock(ctx->pool.shared->m);
        ctx->pool.give_free_to(ctx->pool.shared->pool);
    }
    return ACMMP_OK;
}

int acmmp_reserve_device_memory(int device, size_t bytes)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ACMMP_E_CUDA;
    DeviceShared &sh = device_shared(device);
    std::lock_guard<std::mutex> lock(sh.m);
    if (sh.arena) return ACMMP_E_ARG;                         // one reservation per device (until everything is released)
    if (cudaSetDevice(device) != cudaSuccess) return ACMMP_E_CUDA;
    Trace tr("reserve_device_memory");
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { (void)cudaGetLastError(); return ACMMP_E_CUDA; }
    sh.arena = (char *)p;
    sh.arena_size = bytes;
    sh.arena_used = 0;
    return ACMMP_OK;
}

int acmmp_reserve_pinned(int device, size_t bytes)
{
    if (bytes == 0) return ACMMP_E_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ACMMP_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ACMMP_E_CUDA;
    void *p = nullptr;
    {
        Trace tr("pool.cudaMallocHost");
        if (cudaMallocHost(&p, bytes) != cudaSuccess) { (void)cudaGetLastError(); return ACMMP_E_CUDA; }
    }
    DeviceShared &sh = device_shared(device);
    std::lock_guard<std::mutex> lock(sh.m);
    sh.pool.host_size[p] = bytes;
    sh.pool.host_free.insert(std::make_pair(bytes, p));
    return ACMMP_OK;
}

int acmmp_pool_alloc(int device, size_t bytes, void **out)
{
    if (!out || bytes == 0) return ACMMP_E_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ACMMP_E_CUDA;
    DeviceShared &sh = device_shared(device);
    std::lock_guard<std::mutex> lock(sh.m);
    void *p = nullptr;
    if (!sh.pool.take_dev(bytes, &p) && !sh.carve(bytes, &p)) {
        if (cudaSetDevice(device) != cudaSuccess) return ACMMP_E_CUDA;
        Trace tr("pool.cudaMalloc");
        if (cudaMalloc(&p, bytes) != cudaSuccess) { (void)cudaGetLastError(); return ACMMP_E_CUDA; }
    }
    sh.pool.dev_size[p] = bytes;
    sh.live++;
    *out = p;
    return ACMMP_OK;
}

int acmmp_pool_free(int device, void *p)
{
    if (!p) return ACMMP_OK;
    DeviceShared &sh = device_shared(device);
    std::lock_guard<std::mutex> lock(sh.m);
    if (sh.pool.dev_size.find(p) == sh.pool.dev_size.end()) return ACMMP_E_ARG;
    sh.pool.dfree(p);
    if (--sh.live == 0) sh.release_everything();
    return ACMMP_OK;
}

int acmmp_set_depth_maps(acmmp_ctx *ctx, int n, const float *const *maps, const int32_t *widths, const int32_t *heights)
{
    return set_depths_common(ctx, n, maps, false, widths, heights);
}

int acmmp_set_depth_maps_device(acmmp_ctx *ctx, int n, const float *const *maps_dev, const int32_t *widths,
                                const int32_t *heights)
{
    return set_depths_common(ctx, n, maps_dev, true, widths, heights);
}

int acmmp_set_planes(acmmp_ctx *ctx, const float *planes4, const float *costs)
{
    if (!ctx || !ctx->planes || !planes4 || !costs) return fail(ctx, ACMMP_E_ARG, "acmmp_set_planes: call acmmp_set_views first");
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    CK(cudaMemcpyAsync(ctx->planes, planes4, sizeof(float4) * npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->costs, costs, sizeof(float) * npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ACMMP_OK;
}

int acmmp_set_hierarchy_inputs(acmmp_ctx *ctx, const float *coarse_planes4, int sw, int sh, const float *fine_depth)
{
    if (!ctx || !ctx->planes || !coarse_planes4 || !fine_depth || sw <= 0 || sh <= 0)
        return fail(ctx, ACMMP_E_ARG, "acmmp_set_hierarchy_inputs: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    ctx->pool.dfree(ctx->coarse_planes);
    ctx->coarse_planes = nullptr;
    CK(pmalloc(ctx, &ctx->coarse_planes, sizeof(float4) * (size_t)sw * sh));
    CK(cudaMemcpyAsync(ctx->coarse_planes, coarse_planes4, sizeof(float4) * (size_t)sw * sh, cudaMemcpyHostToDevice, ctx->stream));
    ctx->scaled_cols = sw;
    ctx->scaled_rows = sh;
    // ACMMP.cpp:808-815
    ctx->params.upsample = (sw != ctx->W || sh != ctx->H) ? 1 : 0;
    ctx->params.scaled_cols = (float)sw;
    ctx->params.scaled_rows = (float)sh;
    // plane = (0, 0, 0, fine depth): ACMMP.cpp:833-840 with the unwritten normal defined as 0
    std::vector<float> tmp(4 * npx, 0.f);
    for (size_t i = 0; i < npx; ++i) tmp[4 * i + 3] = fine_depth[i];
    CK(cudaMemcpyAsync(ctx->planes, tmp.data(), sizeof(float4) * npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ACMMP_OK;
}

int acmmp_set_planar_prior_inputs(acmmp_ctx *ctx, const float *plane_params4, int n_planes, const float *masks)
{
    if (!ctx || !ctx->planes || !masks || (n_planes > 0 && !plane_params4))
        return fail(ctx, ACMMP_E_ARG, "acmmp_set_planar_prior_inputs: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int npx = ctx->W * ctx->H;
    float *masks_dev = nullptr;
    float4 *params_dev = nullptr;
    CK(pmalloc(ctx, &masks_dev, sizeof(float) * (size_t)npx));
    CK(pmalloc(ctx, &params_dev, sizeof(float4) * (size_t)std::max(n_planes, 1)));
    CK(cudaMemcpyAsync(masks_dev, masks, sizeof(float) * (size_t)npx, cudaMemcpyHostToDevice, ctx->stream));
    if (n_planes > 0)
        CK(cudaMemcpyAsync(params_dev, plane_params4, sizeof(float4) * (size_t)n_planes, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->prior_planes) CK(pmalloc(ctx, &ctx->prior_planes, sizeof(float4) * (size_t)npx));
    if (!ctx->plane_masks) CK(pmalloc(ctx, &ctx->plane_masks, sizeof(uint32_t) * (size_t)npx));
    k_expand_prior<<<(npx + 255) / 256, 256, 0, ctx->stream>>>(masks_dev, params_dev, std::max(n_planes, 1), npx, ctx->prior_planes,
                                                              ctx->plane_masks);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->pool.dfree(masks_dev);
    ctx->pool.dfree(params_dev);
    ctx->params.planar_prior = 1;
    return ACMMP_OK;
}

namespace {
PriorCam prior_cam(const acmmp_ctx *ctx)
{
    PriorCam c;
    const acmmp_camera &cam = ctx->cams[0];
    c.model = cam.model;
    c.W = ctx->W;
    c.H = ctx->H;
    c.K0 = cam.K[0]; c.K2 = cam.K[2]; c.K4 = cam.K[4]; c.K5 = cam.K[5];
    c.cx = cam.params[1]; c.cy = cam.params[2];
    c.depth_min = ctx->params.depth_min;
    c.depth_max = ctx->params.depth_max;
    return c;
}
} // namespace

int acmmp_support_points(acmmp_ctx *ctx, int32_t *xy, int capacity, int *n)
{
    Trace tr("support_points");
    if (!ctx || !ctx->planes || !xy || !n) return fail(ctx, ACMMP_E_ARG, "acmmp_support_points: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int cells_x = (ctx->W + 4) / 5, cells_y = (ctx->H + 4) / 5, ncell = cells_x * cells_y;
    int2 *cells_dev = nullptr, *cells_host = nullptr;
    PoolTemps tmp(ctx);
    CK(tmp.d(&cells_dev, sizeof(int2) * (size_t)ncell));
    CK(tmp.h(&cells_host, sizeof(int2) * (size_t)ncell));
    k_support_cells<<<(ncell + 255) / 256, 256, 0, ctx->stream>>>(ctx->costs, ctx->W, ctx->H, cells_x, cells_y, cells_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cells_host, cells_dev, sizeof(int2) * (size_t)ncell, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int count = 0;
    for (int i = 0; i < ncell; ++i) {
        if (cells_host[i].x < 0) continue;
        if (count < capacity) {
            xy[2 * count] = cells_host[i].x;
            xy[2 * count + 1] = cells_host[i].y;
        }
        ++count;
    }
    *n = count;
    if (count > capacity) return fail(ctx, ACMMP_E_ARG, "acmmp_support_points: output array too small");
    return ACMMP_OK;
}

int acmmp_download_prior(acmmp_ctx *ctx, float *prior_planes4, uint32_t *plane_masks)
{
    if (!ctx || !ctx->prior_planes || !ctx->plane_masks || !prior_planes4 || !plane_masks)
        return fail(ctx, ACMMP_E_ARG, "acmmp_download_prior: no prior on this context");
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    CK(cudaMemcpyAsync(prior_planes4, ctx->prior_planes, sizeof(float4) * npx, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(plane_masks, ctx->plane_masks, sizeof(uint32_t) * npx, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ACMMP_OK;
}

int acmmp_planar_prior_from_triangles(acmmp_ctx *ctx, const int32_t *tri_xy, int n_tri)
{
    Trace tr("prior_from_triangles");
    if (!ctx || !ctx->planes || n_tri < 0 || (n_tri > 0 && !tri_xy))
        return fail(ctx, ACMMP_E_ARG, "acmmp_planar_prior_from_triangles: bad arguments");
    for (int i = 0; i < 6 * n_tri; i += 2) {
        if (tri_xy[i] < 0 || tri_xy[i] >= ctx->W || tri_xy[i + 1] < 0 || tri_xy[i + 1] >= ctx->H)
            return fail(ctx, ACMMP_E_ARG, "acmmp_planar_prior_from_triangles: triangle vertex outside the image");
    }
    CK(cudaSetDevice(ctx->device));
    tr.mark("check");
    const int npx = ctx->W * ctx->H;
    const PriorCam cam = prior_cam(ctx);
    // temporaries sized by the shape, not by this call's triangle count (support points sit one per 5x5 cell, a
    // triangulation of n points has < 2 n triangles): the pool hands the same blocks back for every view of a level
    const size_t cap = std::max((size_t)n_tri, 2 * (size_t)((ctx->W + 4) / 5) * (size_t)((ctx->H + 4) / 5));
    int *tri_dev = nullptr, *long_list = nullptr, *long_items = nullptr;
    float4 *params_dev = nullptr;
    uint32_t *mask_dev = nullptr;
    PoolTemps tmp(ctx);
    CK(tmp.d(&tri_dev, sizeof(int) * 6 * cap));
    CK(tmp.d(&params_dev, sizeof(float4) * cap));
    CK(tmp.d(&mask_dev, sizeof(uint32_t) * (size_t)npx));
    const int list_capacity = (int)cap + 65536;
    CK(tmp.d(&long_list, sizeof(int) * (size_t)list_capacity));
    CK(tmp.d(&long_items, sizeof(int)));
    if (!ctx->prior_planes) CK(pmalloc(ctx, &ctx->prior_planes, sizeof(float4) * (size_t)npx));
    if (!ctx->plane_masks) CK(pmalloc(ctx, &ctx->plane_masks, sizeof(uint32_t) * (size_t)npx));
    tr.mark("alloc");
    CK(cudaMemsetAsync(mask_dev, 0, sizeof(uint32_t) * (size_t)npx, ctx->stream));
    CK(cudaMemsetAsync(long_items, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(long_list, 0xFF, sizeof(int) * (size_t)list_capacity, ctx->stream));
    if (n_tri > 0) {
        CK(cudaMemcpyAsync(tri_dev, tri_xy, sizeof(int) * 6 * (size_t)n_tri, cudaMemcpyHostToDevice, ctx->stream));
        k_tri_planes<<<(n_tri + 255) / 256, 256, 0, ctx->stream>>>(tri_dev, n_tri, ctx->planes, cam, params_dev);
        k_tri_raster<<<(n_tri + 127) / 128, 128, 0, ctx->stream>>>(tri_dev, n_tri, ctx->W, mask_dev, long_list, list_capacity, long_items);
        k_tri_raster_long<<<148 * 8, kRasterChunk, 0, ctx->stream>>>(tri_dev, ctx->W, mask_dev, long_list, list_capacity, long_items);
        ctx->launches += 3;
    }
    k_prior_finish<<<(npx + 255) / 256, 256, 0, ctx->stream>>>(mask_dev, params_dev, cam, ctx->prior_planes, ctx->plane_masks);
    ctx->launches++;
    CK(cudaGetLastError());
    tr.mark("enqueue");
    CK(cudaStreamSynchronize(ctx->stream));
    tr.mark("wait");
    ctx->params.planar_prior = 1;
    return ACMMP_OK;
}

static int next_level_common(acmmp_ctx *ctx, int n, const float *const *images, bool on_device, const int32_t *widths,
                             const int32_t *heights, const acmmp_camera *cams);

int acmmp_next_level(acmmp_ctx *ctx, int n, const float *const *images, const int32_t *widths, const int32_t *heights,
                     const acmmp_camera *cams)
{
    return next_level_common(ctx, n, images, false, widths, heights, cams);
}

int acmmp_next_level_device(acmmp_ctx *ctx, int n, const float *const *images_dev, const int32_t *widths, const int32_t *heights,
                            const acmmp_camera *cams)
{
    return next_level_common(ctx, n, images_dev, true, widths, heights, cams);
}

static int next_level_common(acmmp_ctx *ctx, int n, const float *const *images, bool on_device, const int32_t *widths,
                             const int32_t *heights, const acmmp_camera *cams)
{
    Trace tr("next_level");
    if (!ctx || !ctx->planes) return fail(ctx, ACMMP_E_ARG, "acmmp_next_level: no previous level on this context");
    CK(cudaSetDevice(ctx->device));
    const int sw = ctx->W, sh = ctx->H, snpx = sw * sh;
    float4 *coarse = nullptr;
    float *coarse_depth = nullptr;
    CK(pmalloc(ctx, &coarse, sizeof(float4) * (size_t)snpx));
    CK(pmalloc(ctx, &coarse_depth, sizeof(float) * (size_t)snpx));
    k_make_coarse<<<(snpx + 255) / 256, 256, 0, ctx->stream>>>(ctx->planes, ctx->costs, snpx, coarse, coarse_depth);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    tr.mark("coarse");
    acmmp_reset_modes(ctx);
    int rc = set_views_common(ctx, n, images, on_device, widths, heights, cams);      // re-allocates at the new size
    tr.mark("set_views");
    if (rc) { ctx->pool.dfree(coarse); ctx->pool.dfree(coarse_depth); return rc; }
    const int npx = ctx->W * ctx->H;
    const int Imagescale = std::max(ctx->H / sh, ctx->W / sw);
    float *fine_depth = nullptr;
    CK(pmalloc(ctx, &fine_depth, sizeof(float) * (size_t)npx));
    if (Imagescale == 1) {
        // RunJBU produces nothing in this case (ACMMP.cpp:1077-1080) and the reference would re-read the old
        // depths.dmb; same size here means the coarse depth is the depth
        if (npx != snpx) { ctx->pool.dfree(coarse); ctx->pool.dfree(coarse_depth); ctx->pool.dfree(fine_depth); return fail(ctx, ACMMP_E_UNSUPPORTED, "level size ratio < 2"); }
        CK(cudaMemcpyAsync(fine_depth, coarse_depth, sizeof(float) * (size_t)npx, cudaMemcpyDeviceToDevice, ctx->stream));
        k_coarse_w_from_depth<<<(snpx + 255) / 256, 256, 0, ctx->stream>>>(coarse_depth, snpx, coarse);      // ACMMP.cpp:826-828
        ctx->launches++;
        CK(cudaGetLastError());
    } else {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
        rc = acmmp_jbu_device(ctx->device, ctx->ref_dense, ctx->W, ctx->H, coarse_depth, sw, sh, fine_depth, ctx->stream);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ctx->t_jbu, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        tr.mark("jbu_wait");
        ctx->launches++;
        if (rc) { ctx->pool.dfree(coarse); ctx->pool.dfree(coarse_depth); ctx->pool.dfree(fine_depth); return fail(ctx, rc, "JBU failed"); }
    }
    k_seed_planes_from_depth<<<(npx + 255) / 256, 256, 0, ctx->stream>>>(fine_depth, npx, ctx->planes);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->pool.dfree(ctx->coarse_planes);
    ctx->coarse_planes = coarse;
    ctx->scaled_cols = sw;
    ctx->scaled_rows = sh;
    ctx->params.hierarchy = 1;
    ctx->params.upsample = (sw != ctx->W || sh != ctx->H) ? 1 : 0;
    ctx->params.scaled_cols = (float)sw;
    ctx->params.scaled_rows = (float)sh;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->pool.dfree(coarse_depth);
    ctx->pool.dfree(fine_depth);
    return ACMMP_OK;
}

int acmmp_result_host(acmmp_ctx *ctx, const float **planes4, const float **costs)
{
    if (!ctx || !ctx->have_result) return fail(ctx, ACMMP_E_ARG, "acmmp_result_host: no completed acmmp_run_patch_match");
    if (planes4) *planes4 = reinterpret_cast<const float *>(ctx->planes_host);
    if (costs) *costs = ctx->costs_host;
    return ACMMP_OK;
}

int acmmp_set_seed(acmmp_ctx *ctx, uint64_t seed)
{
    if (!ctx) return ACMMP_E_ARG;
    ctx->seed = seed;
    return ACMMP_OK;
}

int acmmp_set_sphere_tap_pruning(acmmp_ctx *ctx, float relative_weight)
{
    if (!ctx || !(relative_weight >= 0.0f) || relative_weight > 1e-3f) return fail(ctx, ACMMP_E_ARG, "acmmp_set_sphere_tap_pruning: threshold must be in [0, 1e-3]");
    ctx->tap_prune = relative_weight;
    return ACMMP_OK;
}

int acmmp_set_plane_now_semantics(acmmp_ctx *ctx, int as_compiled)
{
    if (!ctx) return ACMMP_E_ARG;
    ctx->as_compiled = as_compiled ? 1 : 0;
    return ACMMP_OK;
}

int acmmp_random_init(acmmp_ctx *ctx) { return do_init(ctx); }

int acmmp_checkerboard_pass(acmmp_ctx *ctx, int colour, int iter)
{
    if (!ctx || (colour != 0 && colour != 1)) return ACMMP_E_ARG;
    return do_pass(ctx, colour, iter);
}

int acmmp_finalize(acmmp_ctx *ctx) { return do_finalize(ctx); }

int acmmp_synchronize(acmmp_ctx *ctx)
{
    Trace tr("synchronize");
    if (!ctx) return ACMMP_E_ARG;
    CK(cudaSetDevice(ctx->device));
    return collect_timings(ctx);
}

static int ensure_host_result(acmmp_ctx *ctx)
{
    const size_t npx = (size_t)ctx->W * ctx->H;
    if (!ctx->planes_host) CK(phmalloc(ctx, &ctx->planes_host, sizeof(float4) * npx));
    if (!ctx->costs_host) CK(phmalloc(ctx, &ctx->costs_host, sizeof(float) * npx));
    return ACMMP_OK;
}

static int run_stage(acmmp_ctx *ctx, bool download)
{
    int rc = do_init(ctx);
    if (rc) return rc;
    for (int i = 0; i < ctx->params.max_iterations; ++i) {
        if ((rc = do_pass(ctx, 0, i))) return rc;
        if ((rc = do_pass(ctx, 1, i))) return rc;
    }
    if ((rc = do_finalize(ctx))) return rc;
    ctx->have_result = false;
    if (download) {
        const size_t npx = (size_t)ctx->W * ctx->H;
        { const int rc_h = ensure_host_result(ctx); if (rc_h) return rc_h; }
        CK(cudaMemcpyAsync(ctx->planes_host, ctx->planes, sizeof(float4) * npx, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->costs_host, ctx->costs, sizeof(float) * npx, cudaMemcpyDeviceToHost, ctx->stream));
    }
    rc = collect_timings(ctx);
    if (rc) return rc;
    CK(cudaGetLastError());
    ctx->have_result = download;
    return ACMMP_OK;
}

int acmmp_run_patch_match(acmmp_ctx *ctx) { return ctx ? run_stage(ctx, true) : ACMMP_E_ARG; }

int acmmp_run_patch_match_resident(acmmp_ctx *ctx) { return ctx ? run_stage(ctx, false) : ACMMP_E_ARG; }

int acmmp_download_result(acmmp_ctx *ctx)
{
    Trace tr("download_result");
    if (!ctx || !ctx->planes) return ACMMP_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    { const int rc_h = ensure_host_result(ctx); if (rc_h) return rc_h; }
    CK(cudaMemcpyAsync(ctx->planes_host, ctx->planes, sizeof(float4) * npx, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->costs_host, ctx->costs, sizeof(float) * npx, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_result = true;
    return ACMMP_OK;
}

int acmmp_get_result(acmmp_ctx *ctx, float *planes4, float *costs)
{
    if (!ctx || !ctx->have_result) return fail(ctx, ACMMP_E_ARG, "acmmp_get_result: no completed acmmp_run_patch_match");
    const size_t npx = (size_t)ctx->W * ctx->H;
    if (planes4) std::memcpy(planes4, ctx->planes_host, sizeof(float4) * npx);
    if (costs) std::memcpy(costs, ctx->costs_host, sizeof(float) * npx);
    return ACMMP_OK;
}

int acmmp_width(const acmmp_ctx *ctx) { return ctx ? ctx->W : 0; }
int acmmp_height(const acmmp_ctx *ctx) { return ctx ? ctx->H : 0; }

int acmmp_device_buffers(acmmp_ctx *ctx, void **planes4_dev, void **costs_dev)
{
    if (!ctx || !ctx->planes) return ACMMP_E_ARG;
    if (planes4_dev) *planes4_dev = ctx->planes;
    if (costs_dev) *costs_dev = ctx->costs;
    return ACMMP_OK;
}

int acmmp_export_depth_device(acmmp_ctx *ctx, float *depth_dev)
{
    Trace tr("export_depth");
    if (!ctx || !ctx->planes || !depth_dev) return ACMMP_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int npx = ctx->W * ctx->H;
    k_export_depth<<<(npx + 255) / 256, 256, 0, ctx->stream>>>(ctx->planes, npx, depth_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    return ACMMP_OK;
}

int acmmp_export_depth_device_sync(acmmp_ctx *ctx, float *depth_dev)
{
    const int rc = acmmp_export_depth_device(ctx, depth_dev);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return ACMMP_OK;
}

int acmmp_download_state(acmmp_ctx *ctx, float *planes4, float *costs, uint32_t *selected_views, uint32_t *rand6,
                         float *pre_costs)
{
    if (ctx && ctx->parked) return fail(ctx, ACMMP_E_ARG, "the context is parked");
    if (!ctx || !ctx->planes) return ACMMP_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    CK(cudaStreamSynchronize(ctx->stream));
    if (planes4) CK(cudaMemcpy(planes4, ctx->planes, sizeof(float4) * npx, cudaMemcpyDeviceToHost));
    if (costs) CK(cudaMemcpy(costs, ctx->costs, sizeof(float) * npx, cudaMemcpyDeviceToHost));
    if (selected_views) CK(cudaMemcpy(selected_views, ctx->selected_views, sizeof(uint32_t) * npx, cudaMemcpyDeviceToHost));
    if (rand6) CK(cudaMemcpy(rand6, ctx->rng, sizeof(uint32_t) * 6 * npx, cudaMemcpyDeviceToHost));
    if (pre_costs) CK(cudaMemcpy(pre_costs, ctx->pre_costs, sizeof(float) * npx, cudaMemcpyDeviceToHost));
    return ACMMP_OK;
}

int acmmp_upload_state(acmmp_ctx *ctx, const float *planes4, const float *costs, const uint32_t *selected_views,
                       const uint32_t *rand6, const float *pre_costs)
{
    if (ctx && ctx->parked) return fail(ctx, ACMMP_E_ARG, "the context is parked");
    if (!ctx || !ctx->planes) return ACMMP_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    CK(cudaStreamSynchronize(ctx->stream));
    if (planes4) CK(cudaMemcpy(ctx->planes, planes4, sizeof(float4) * npx, cudaMemcpyHostToDevice));
    if (costs) CK(cudaMemcpy(ctx->costs, costs, sizeof(float) * npx, cudaMemcpyHostToDevice));
    if (selected_views) CK(cudaMemcpy(ctx->selected_views, selected_views, sizeof(uint32_t) * npx, cudaMemcpyHostToDevice));
    if (rand6) CK(cudaMemcpy(ctx->rng, rand6, sizeof(uint32_t) * 6 * npx, cudaMemcpyHostToDevice));
    if (pre_costs) CK(cudaMemcpy(ctx->pre_costs, pre_costs, sizeof(float) * npx, cudaMemcpyHostToDevice));
    return ACMMP_OK;
}

int acmmp_jbu_device(int device, const float *image_dev, int w, int h, const float *coarse_depth_dev, int sw, int sh,
                     float *out_depth_dev, void *cuda_stream)
{
    if (!image_dev || !coarse_depth_dev || !out_depth_dev || w <= 0 || h <= 0 || sw <= 0 || sh <= 0) return ACMMP_E_ARG;
    const int Imagescale = std::max(h / sh, w / sw);      // ACMMP.cpp:1075 (integer division)
    if (Imagescale == 1) return ACMMP_E_ARG;              // ACMMP.cpp:1077-1080: nothing is produced
    if (cudaSetDevice(device) != cudaSuccess) return ACMMP_E_CUDA;
    dim3 grid((w + 15) / 16, (h + 15) / 16);
    k_jbu<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(image_dev, w, h, coarse_depth_dev, sw, sh, Imagescale, out_depth_dev);
    return cudaGetLastError() == cudaSuccess ? ACMMP_OK : ACMMP_E_CUDA;
}

static thread_local float g_last_jbu_ms = 0.f;
float acmmp_last_jbu_ms(void) { return g_last_jbu_ms; }

int acmmp_jbu(int device, const float *image, int w, int h, const float *coarse_depth, int sw, int sh, float *out_depth)
{
    if (!image || !coarse_depth || !out_depth || w <= 0 || h <= 0 || sw <= 0 || sh <= 0) return ACMMP_E_ARG;
    if (std::max(h / sh, w / sw) == 1) return ACMMP_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return ACMMP_E_CUDA;
    float *di = nullptr, *dd = nullptr, *dout = nullptr;
    int rc = ACMMP_E_CUDA;
    if (cudaMalloc(&di, sizeof(float) * (size_t)w * h) == cudaSuccess && cudaMalloc(&dd, sizeof(float) * (size_t)sw * sh) == cudaSuccess &&
        cudaMalloc(&dout, sizeof(float) * (size_t)w * h) == cudaSuccess &&
        cudaMemcpy(di, image, sizeof(float) * (size_t)w * h, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(dd, coarse_depth, sizeof(float) * (size_t)sw * sh, cudaMemcpyHostToDevice) == cudaSuccess) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, nullptr);
        rc = acmmp_jbu_device(device, di, w, h, dd, sw, sh, dout, nullptr);
        cudaEventRecord(e1, nullptr);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&g_last_jbu_ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (rc == ACMMP_OK && cudaMemcpy(out_depth, dout, sizeof(float) * (size_t)w * h, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = ACMMP_E_CUDA;
    }
    cudaFree(di); cudaFree(dd); cudaFree(dout);
    return rc;
}

int acmmp_probe_coords(acmmp_ctx *ctx, const float *planes4, int view, int variant, float *out72)
{
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!planes4 || !out72 || view < 1 || view >= ctx->n) return fail(ctx, ACMMP_E_ARG, "acmmp_probe_coords: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)ctx->W * ctx->H;
    PoolTemps tmp(ctx);
    float4 *dp = nullptr;
    float *dout = nullptr;
    CK(tmp.d(&dp, sizeof(float4) * npx));
    CK(tmp.d(&dout, sizeof(float) * 72 * npx));
    CK(cudaMemcpyAsync(dp, planes4, sizeof(float4) * npx, cudaMemcpyHostToDevice, ctx->stream));
    const FrameConst fc = frame_const(ctx);
    dim3 grid((ctx->W + 15) / 16, (ctx->H + 7) / 8);
    if (ctx->cams[0].model == ACMMP_MODEL_PINHOLE) k_probe_coords<kModelPinhole><<<grid, 128, 0, ctx->stream>>>(fc, ctx->ncc, view, variant, dp, dout);
    else k_probe_coords<kModelSphere><<<grid, 128, 0, ctx->stream>>>(fc, ctx->ncc, view, variant, dp, dout);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out72, dout, sizeof(float) * 72 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ACMMP_OK;
}

int acmmp_probe_ncc(acmmp_ctx *ctx, const float *planes4, int view, float *out) { return run_probe(ctx, 0, view, planes4, out, nullptr, nullptr); }
int acmmp_probe_geom(acmmp_ctx *ctx, const float *planes4, int view, float *out) { return run_probe(ctx, 1, view, planes4, out, nullptr, nullptr); }
int acmmp_probe_warp(acmmp_ctx *ctx, const float *planes4, int view, float *out4) { return run_probe(ctx, 2, view, planes4, nullptr, out4, nullptr); }
int acmmp_probe_initcost(acmmp_ctx *ctx, const float *planes4, float *out, uint32_t *selected_views)
{
    return run_probe(ctx, 3, 1, planes4, out, nullptr, selected_views);
}

int acmmp_last_timings(acmmp_ctx *ctx, float what[8])
{
    if (!ctx || !what) return ACMMP_E_ARG;
    for (int i = 0; i < 8; ++i) what[i] = 0.f;
    what[0] = ctx->t_init;
    what[1] = ctx->t_pass_sum;
    what[2] = ctx->t_finalize;
    what[3] = (float)ctx->n_pass;
    what[4] = ctx->t_last_pass;
    what[5] = ctx->t_jbu;
    return ACMMP_OK;
}

int64_t acmmp_launch_count(const acmmp_ctx *ctx) { return ctx ? ctx->launches : 0; }


// ---------------------------------------------------------------------------------------------
// fusion
// ---------------------------------------------------------------------------------------------
struct acmmp_fusion {
    int device = 0;
    int n = 0;
    cudaStream_t stream = nullptr;
    std::vector<acmmp::FusionViewDev> views;          // host copy of the table
    std::vector<std::vector<void *>> owned;           // device buffers this object allocated, per view
    acmmp::FusionViewDev *views_dev = nullptr;
    bool table_dirty = true;
    acmmp_point *dense = nullptr, *out = nullptr;
    unsigned char *flags = nullptr;
    int *counts = nullptr, *total = nullptr;
    size_t dense_cap = 0, out_cap = 0, count_cap = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    std::string err;
};

#define FCK(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            f->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
            return ACMMP_E_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

static int fusion_fail(acmmp_fusion *f, int code, const std::string &msg)
{
    if (f) f->err = msg;
    return code;
}

static void fusion_free_view(acmmp_fusion *f, int i)
{
    for (void *p : f->owned[i]) cudaFree(p);
    f->owned[i].clear();
}

int acmmp_fusion_create(int device, int n_views, acmmp_fusion **out)
{
    if (!out || n_views < 1) return ACMMP_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ACMMP_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return ACMMP_E_CUDA;      // sm_100a code only
    acmmp_fusion *f = new acmmp_fusion();
    f->device = device;
    f->n = n_views;
    f->views.resize(n_views);
    f->owned.resize(n_views);
    for (auto &v : f->views) std::memset(&v, 0, sizeof(v));
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&f->views_dev, sizeof(acmmp::FusionViewDev) * n_views) != cudaSuccess || cudaMalloc(&f->total, sizeof(int)) != cudaSuccess ||
        cudaEventCreate(&f->ev[0]) != cudaSuccess || cudaEventCreate(&f->ev[1]) != cudaSuccess) {
        acmmp_fusion_destroy(f);
        return ACMMP_E_CUDA;
    }
    *out = f;
    return ACMMP_OK;
}

void acmmp_fusion_destroy(acmmp_fusion *f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    for (int i = 0; i < f->n; ++i) fusion_free_view(f, i);
    cudaFree(f->views_dev); cudaFree(f->dense); cudaFree(f->out); cudaFree(f->flags); cudaFree(f->counts); cudaFree(f->total);
    if (f->ev[0]) cudaEventDestroy(f->ev[0]);
    if (f->ev[1]) cudaEventDestroy(f->ev[1]);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

const char *acmmp_fusion_last_error(const acmmp_fusion *f) { return f ? f->err.c_str() : "null fusion object"; }

int acmmp_fusion_set_view_device(acmmp_fusion *f, int index, const acmmp_camera *cam, int w, int h, const float *depth_dev,
                                 const void *normals4_dev, const float *gray_dev)
{
    if (!f || index < 0 || index >= f->n || !cam || w <= 0 || h <= 0 || !depth_dev || !normals4_dev || !gray_dev)
        return fusion_fail(f, ACMMP_E_ARG, "acmmp_fusion_set_view_device: bad arguments");
    FCK(cudaSetDevice(f->device));
    fusion_free_view(f, index);
    acmmp::FusionViewDev &v = f->views[index];
    v.cam = *cam;
    v.cam.width = w;
    v.cam.height = h;
    v.depth = depth_dev;
    v.normal = static_cast<const float4 *>(normals4_dev);
    v.gray = gray_dev;
    v.bgr = nullptr;
    f->table_dirty = true;
    return ACMMP_OK;
}

int acmmp_fusion_set_view(acmmp_fusion *f, int index, const acmmp_camera *cam, int w, int h, const float *depth,
                          const float *normals3, const float *gray)
{
    if (!f || index < 0 || index >= f->n || !cam || w <= 0 || h <= 0 || !depth || !normals3 || !gray)
        return fusion_fail(f, ACMMP_E_ARG, "acmmp_fusion_set_view: bad arguments");
    FCK(cudaSetDevice(f->device));
    fusion_free_view(f, index);
    const size_t npx = (size_t)w * h;
    float *d = nullptr, *g = nullptr, *n3 = nullptr;
    float4 *n4 = nullptr;
    FCK(cudaMalloc(&d, sizeof(float) * npx));
    f->owned[index].push_back(d);
    FCK(cudaMalloc(&g, sizeof(float) * npx));
    f->owned[index].push_back(g);
    FCK(cudaMalloc(&n4, sizeof(float4) * npx));
    f->owned[index].push_back(n4);
    FCK(cudaMalloc(&n3, sizeof(float) * 3 * npx));
    FCK(cudaMemcpyAsync(d, depth, sizeof(float) * npx, cudaMemcpyHostToDevice, f->stream));
    FCK(cudaMemcpyAsync(g, gray, sizeof(float) * npx, cudaMemcpyHostToDevice, f->stream));
    FCK(cudaMemcpyAsync(n3, normals3, sizeof(float) * 3 * npx, cudaMemcpyHostToDevice, f->stream));
    acmmp::k_pack_normals<<<(unsigned)((npx + 255) / 256), 256, 0, f->stream>>>(n3, (int)npx, n4);
    cudaError_t e = cudaStreamSynchronize(f->stream);
    cudaFree(n3);
    FCK(e);
    acmmp::FusionViewDev &v = f->views[index];
    v.cam = *cam;
    v.cam.width = w;
    v.cam.height = h;
    v.depth = d;
    v.normal = n4;
    v.gray = g;
    v.bgr = nullptr;
    f->table_dirty = true;
    return ACMMP_OK;
}

int acmmp_fusion_set_view_colour(acmmp_fusion *f, int index, const uint8_t *bgr, int w, int h)
{
    if (!f || index < 0 || index >= f->n || !bgr) return fusion_fail(f, ACMMP_E_ARG, "acmmp_fusion_set_view_colour: bad arguments");
    acmmp::FusionViewDev &v = f->views[index];
    if (!v.depth || v.cam.width != w || v.cam.height != h)
        return fusion_fail(f, ACMMP_E_ARG, "acmmp_fusion_set_view_colour: set the view first; the colour image must have the depth map's size");
    FCK(cudaSetDevice(f->device));
    const size_t npx = (size_t)w * h;
    std::vector<uchar4> packed(npx);
    for (size_t k = 0; k < npx; ++k) packed[k] = make_uchar4(bgr[3 * k], bgr[3 * k + 1], bgr[3 * k + 2], 255);
    uchar4 *dev = nullptr;
    FCK(cudaMalloc(&dev, sizeof(uchar4) * npx));
    f->owned[index].push_back(dev);
    FCK(cudaMemcpy(dev, packed.data(), sizeof(uchar4) * npx, cudaMemcpyHostToDevice));
    v.bgr = dev;
    f->table_dirty = true;
    return ACMMP_OK;
}

// fuse one reference view and bring its points to the host: PointList records (36 bytes) or PLY vertex records (27 bytes)
static int fusion_run_common(acmmp_fusion *f, int ref, int n_src, const int32_t *src, void *host_out, int capacity, int *n_points,
                             float *kernel_ms, bool ply, const char *what)
{
    if (!f || ref < 0 || ref >= f->n || n_src < 0 || (n_src > 0 && !src) || !n_points || capacity < 0 || (capacity > 0 && !host_out))
        return fusion_fail(f, ACMMP_E_ARG, std::string(what) + ": bad arguments");
    if (!f->views[ref].depth) return fusion_fail(f, ACMMP_E_ARG, std::string(what) + ": the reference view was not set");
    FCK(cudaSetDevice(f->device));
    acmmp::FusionProblemDev prob;
    prob.num_src = std::min(n_src, acmmp::kFuseMaxSrc);
    for (int j = 0; j < acmmp::kFuseMaxSrc; ++j) {
        int s = (j < prob.num_src) ? src[j] : -1;
        if (s >= f->n || (s >= 0 && !f->views[s].depth)) s = -1;
        prob.src[j] = s;
    }
    if (f->table_dirty) {
        FCK(cudaMemcpyAsync(f->views_dev, f->views.data(), sizeof(acmmp::FusionViewDev) * f->n, cudaMemcpyHostToDevice, f->stream));
        FCK(cudaStreamSynchronize(f->stream));      // the host table may change right after
        f->table_dirty = false;
    }
    const int npx = f->views[ref].cam.width * f->views[ref].cam.height;
    const int nblocks = (npx + acmmp::kFuseBlock - 1) / acmmp::kFuseBlock;
    if ((size_t)npx > f->dense_cap) {
        cudaFree(f->dense); cudaFree(f->flags);
        f->dense = nullptr; f->flags = nullptr; f->dense_cap = 0;
        FCK(cudaMalloc(&f->dense, sizeof(acmmp_point) * (size_t)npx));
        FCK(cudaMalloc(&f->flags, (size_t)npx));
        f->dense_cap = (size_t)npx;
    }
    if ((size_t)nblocks > f->count_cap) {
        cudaFree(f->counts);
        f->counts = nullptr; f->count_cap = 0;
        FCK(cudaMalloc(&f->counts, sizeof(int) * (size_t)nblocks));
        f->count_cap = (size_t)nblocks;
    }
    // the compacted output never exceeds one record per pixel: sized by the view, so that a host buffer that turns out too
    // small costs a second copy, not a second run
    const int dev_capacity = std::min(capacity, npx);
    if ((size_t)dev_capacity > f->out_cap) {
        cudaFree(f->out);
        f->out = nullptr; f->out_cap = 0;
        FCK(cudaMalloc(&f->out, sizeof(acmmp_point) * (size_t)dev_capacity));
        f->out_cap = (size_t)dev_capacity;
    }
    cudaEventRecord(f->ev[0], f->stream);
    acmmp::k_fuse_view<<<nblocks, acmmp::kFuseBlock, 0, f->stream>>>(f->views_dev, ref, prob, f->dense, f->flags, f->counts);
    acmmp::k_scan_blocks<<<1, 1024, 0, f->stream>>>(f->counts, nblocks, f->total);
    if (dev_capacity > 0) {
        if (ply) acmmp::k_compact_ply<<<nblocks, acmmp::kFuseBlock, 0, f->stream>>>(f->dense, f->flags, f->counts, npx, reinterpret_cast<unsigned char *>(f->out), dev_capacity);
        else acmmp::k_compact_points<<<nblocks, acmmp::kFuseBlock, 0, f->stream>>>(f->dense, f->flags, f->counts, npx, f->out, dev_capacity);
    }
    cudaEventRecord(f->ev[1], f->stream);
    FCK(cudaGetLastError());
    int total = 0;
    FCK(cudaMemcpyAsync(&total, f->total, sizeof(int), cudaMemcpyDeviceToHost, f->stream));
    FCK(cudaStreamSynchronize(f->stream));
    *n_points = total;
    if (kernel_ms) cudaEventElapsedTime(kernel_ms, f->ev[0], f->ev[1]);
    if (total > capacity) return fusion_fail(f, ACMMP_E_ARG, std::string(what) + ": capacity too small (n_points holds the need)");
    if (total > 0) FCK(cudaMemcpy(host_out, f->out, (ply ? (size_t)27 : sizeof(acmmp_point)) * (size_t)total, cudaMemcpyDeviceToHost));
    return ACMMP_OK;
}

int acmmp_fusion_run(acmmp_fusion *f, int ref, int n_src, const int32_t *src, acmmp_point *points, int capacity,
                     int *n_points, float *kernel_ms)
{
    return fusion_run_common(f, ref, n_src, src, points, capacity, n_points, kernel_ms, false, "acmmp_fusion_run");
}

int acmmp_fusion_run_ply(acmmp_fusion *f, int ref, int n_src, const int32_t *src, uint8_t *records27, int capacity,
                         int *n_points, float *kernel_ms)
{
    return fusion_run_common(f, ref, n_src, src, records27, capacity, n_points, kernel_ms, true, "acmmp_fusion_run_ply");
}

/* test / inspection hook: which pixels of the last fused reference view produced a point (w*h bytes, host) */
int acmmp_fusion_last_flags(acmmp_fusion *f, int ref, unsigned char *flags)
{
    if (!f || ref < 0 || ref >= f->n || !flags || !f->flags) return fusion_fail(f, ACMMP_E_ARG, "acmmp_fusion_last_flags: nothing fused yet");
    FCK(cudaSetDevice(f->device));
    const size_t npx = (size_t)f->views[ref].cam.width * f->views[ref].cam.height;
    if (npx > f->dense_cap) return fusion_fail(f, ACMMP_E_ARG, "acmmp_fusion_last_flags: not the view of the last run");
    FCK(cudaMemcpy(flags, f->flags, npx, cudaMemcpyDeviceToHost));
    return ACMMP_OK;
}

} // extern "C"
