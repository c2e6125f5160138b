/* include/acmmp_b200.h -- C ABI of libacmmp_b200.so
 *
 * B200-native (sm_100a) PatchMatch depth/normal estimation, drop-in for the device side of the
 * reference's `ACMMP` host class (reference ACMMP.h:57-111).  Plain pointers and sizes only; every
 * entry point returns 0 on success or a negative ACMMP_E_* code (the reference prints and calls
 * exit(), ACMMP.cpp:64-97; a library must not).  There is NO CPU fallback: every compute entry
 * point fails with ACMMP_E_CUDA when no sm_100 device is usable.
 *
 * Which reference interface each entry point replaces is cited per declaration
 * (file:line into the reference tree).
 */
#ifndef ACMMP_B200_H_
#define ACMMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACMMP_OK 0
#define ACMMP_E_ARG (-1)       /* bad argument / call order */
#define ACMMP_E_CUDA (-2)      /* CUDA runtime or driver failure (message: acmmp_last_error) */
#define ACMMP_E_UNSUPPORTED (-3)

#define ACMMP_MODEL_PINHOLE 0  /* reference main.h:35-38 */
#define ACMMP_MODEL_SPHERE 11
#define ACMMP_MAX_SRC 32       /* reference ACMMP.cu:522 (cost_vector[32]) and the 32-bit view mask */

/* Byte-for-byte the reference `struct Camera` (main.h:40-54), 120 bytes. */
typedef struct acmmp_camera {
    int32_t model;       /* ACMMP_MODEL_* */
    float params[4];     /* SPHERE: f, cx, cy, unused */
    float R[9];          /* world -> camera, row major */
    float t[3];
    float K[9];          /* PINHOLE intrinsics, row major */
    int32_t width, height;
    float depth_min, depth_max;
} acmmp_camera;

/* Byte-for-byte the reference `struct PatchMatchParams` (ACMMP.h:32-55), 68 bytes.
 * patch_size / radius_increment / sigma_* / top_k are accepted only at the reference defaults
 * (11 / 2 / 5 / 3 / 4): the kernels are specialised for the 6x6-tap window. */
typedef struct acmmp_params {
    int32_t max_iterations;
    int32_t patch_size;
    int32_t num_images;
    int32_t max_image_size;
    int32_t radius_increment;
    float sigma_spatial;
    float sigma_color;
    int32_t top_k;
    float baseline;
    float depth_min;
    float depth_max;
    float disparity_min;
    float disparity_max;
    float scaled_cols;
    float scaled_rows;
    uint8_t geom_consistency;
    uint8_t planar_prior;
    uint8_t multi_geometry;
    uint8_t hierarchy;
    uint8_t upsample;
    uint8_t pad_[3];
} acmmp_params;

typedef struct acmmp_ctx acmmp_ctx;

/* Fill *p with the reference defaults (ACMMP.h:33-54). */
void acmmp_default_params(acmmp_params *p);

/* Library / build information (never touches the GPU). */
const char *acmmp_version(void);
int acmmp_abi_sizeof_camera(void);
int acmmp_abi_sizeof_params(void);

/* One context per reference view being processed == one reference `ACMMP` object
 * (constructor ACMMP.cpp:99, destructor :101-143).  `device` is the CUDA ordinal
 * (the reference hard-codes cudaSetDevice(0), main.cpp:77). */
int acmmp_create(acmmp_ctx **out, int device);
int acmmp_destroy(acmmp_ctx *ctx);
const char *acmmp_last_error(const acmmp_ctx *ctx);

/* Replaces the upload half of ACMMP::CudaSpaceInitialization (ACMMP.cpp:685-724): image 0 is the
 * reference view, images 1..n-1 the source views; images[i] is a dense row-major float32 array
 * widths[i] x heights[i] in HOST memory (grey levels 0..255 as produced by InuputInitialization,
 * ACMMP.cpp:578-581, :626-628).  cams[i].width/height are overwritten from widths/heights.
 * Source views become R32F bilinear textures (clamp addressing, un-normalised coordinates: what
 * ACMMP.cpp:698-704 effectively configures); the reference view additionally becomes a
 * border-replicated pitch-linear image that the kernels stage into shared memory with TMA.
 * params->depth_min/max and num_images are derived as ACMMP.cpp:645-648 does. */
int acmmp_set_views(acmmp_ctx *ctx, int n, const float *const *images, const int32_t *widths,
                    const int32_t *heights, const acmmp_camera *cams);

/* Same, but images[i] are DEVICE pointers (dense row-major float32) on the context's device:
 * the GPU-resident pipeline and bench.py's device-resident arm use this. */
int acmmp_set_views_device(acmmp_ctx *ctx, int n, const float *const *images_dev, const int32_t *widths,
                           const int32_t *heights, const acmmp_camera *cams);

/* Mode flags, the reference setters ACMMP.cpp:548-565 (SetGeomConsistencyParams,
 * SetHierarchyParams, SetPlanarPriorParams) plus max_iterations. */
int acmmp_set_geom_consistency(acmmp_ctx *ctx, int multi_geometry);
int acmmp_set_hierarchy(acmmp_ctx *ctx);
int acmmp_set_planar_prior(acmmp_ctx *ctx);
int acmmp_set_max_iterations(acmmp_ctx *ctx, int n);
int acmmp_get_params(const acmmp_ctx *ctx, acmmp_params *out);
/* Back to a freshly constructed object's flags (geom / prior / hierarchy off, max_iterations 3) while
 * keeping the uploaded views: lets one context serve consecutive stages of the same view and level
 * (the reference constructs a new ACMMP object per stage and re-uploads everything, main.cpp:83-93). */
int acmmp_reset_modes(acmmp_ctx *ctx);
/* Between two stages of a view that stays resident: give every buffer except the stage state back to the device's pool
 * (textures, padded reference, ping-pong / per-stage buffers, RNG states, neighbour depth-map table -- ~1 GB of the
 * ~1.15 GB a 3200x2130 view with 10 source views holds) so that the NEXT view's stage takes the same blocks instead of
 * allocating its own.  What stays: planes + costs, the coarse planes of the hierarchy hand-over, with keep_prior != 0 the
 * planar prior (planes + mask) and with keep_host_result != 0 the pinned result buffers acmmp_result_host points into.
 * Waits for the context's stream.  The reference's counterpart is ~ACMMP() between stages (ACMMP.cpp:101-143), which
 * frees everything because the state travels through .dmb files.  To go on: acmmp_set_views[_device] with the same shapes
 * (re-uploads the images, keeps the state; other shapes start the view afresh), then the mode setters, neighbour depth
 * maps and acmmp_run_patch_match* as usual.  acmmp_support_points, acmmp_planar_prior_from_triangles,
 * acmmp_export_depth_device, acmmp_download_result and acmmp_next_level* work on a parked context. */
int acmmp_park(acmmp_ctx *ctx, int keep_prior, int keep_host_result);
/* The pool of a device, for hosts that keep many views resident (no reference counterpart: the reference allocates per
 * ProcessProblem call, ACMMP.cpp:685-724).  acmmp_reserve_device_memory: one allocation of `bytes` that the pool carves
 * new blocks out of before it falls back to cudaMalloc (cudaMalloc on a device that already holds hundreds of allocations
 * costs milliseconds per call); one reservation per device, released with everything else when the last context of the
 * device is destroyed.  acmmp_pool_alloc / acmmp_pool_free: device buffers of the host (image pool, depth-map tables) out
 * of the same pool; a block handed out keeps the pool alive like a context does.  Thread-safe. */
int acmmp_reserve_device_memory(int device, size_t bytes);
/* One page-locked host block of exactly `bytes` into the device's pool, ahead of time: the result buffers of
 * acmmp_download_result / acmmp_run_patch_match are 16 and 4 bytes per pixel, and a page-locked allocation in the middle
 * of a multi-threaded run costs 20 - 100 ms. */
int acmmp_reserve_pinned(int device, size_t bytes);
int acmmp_pool_alloc(int device, size_t bytes, void **out);
int acmmp_pool_free(int device, void *p);

/* Geometric consistency inputs: the n depth maps (index 0 = reference view, 1.. = source views)
 * that InuputInitialization reads from depths.dmb / depths_geom.dmb (ACMMP.cpp:653-678) and
 * CudaSpaceInitialization uploads (ACMMP.cpp:726-751).  Host pointers, dense float32. */
int acmmp_set_depth_maps(acmmp_ctx *ctx, int n, const float *const *maps, const int32_t *widths,
                         const int32_t *heights);
int acmmp_set_depth_maps_device(acmmp_ctx *ctx, int n, const float *const *maps_dev, const int32_t *widths,
                                const int32_t *heights);

/* Previous-stage state of the reference view, W*H float4 (world normal xyz, depth w) + W*H costs:
 * the reload at ACMMP.cpp:753-785 (geom mode).  Host pointers. */
int acmmp_set_planes(acmmp_ctx *ctx, const float *planes4, const float *costs);

/* Hierarchy inputs (ACMMP.cpp:788-844): coarse-level (normal xyz, w) map of size sw x sh -- w is the
 * coarse COST when sw x sh differs from the image size (upsample mode, ACMMP.cpp:823-825), else the
 * depth -- and the fine-level depth map (JBU output) that seeds plane.w (ACMMP.cpp:833-840; the
 * normal part the reference leaves uninitialised is defined as 0 here). */
int acmmp_set_hierarchy_inputs(acmmp_ctx *ctx, const float *coarse_planes4, int sw, int sh,
                               const float *fine_depth);

/* GPU-resident stage chaining: what the reference does through .dmb files and a
 * fresh ACMMP object per stage (main.cpp:199-208 -> ACMMP.cpp:753-801), without leaving the device.
 *
 * acmmp_set_depth_maps / _device accept maps[0] == NULL: "the reference view's depth map is the depth of
 * the state currently on the device" (the planes left by the previous stage, which are also exactly what
 * the geometric stage reloads, ACMMP.cpp:772-785, so no acmmp_set_planes is needed).
 *
 * acmmp_next_level: move the context to the next (finer) pyramid level.  Takes the new level's views like
 * acmmp_set_views; the previous level's result is joint-bilaterally upsampled on the device (RunJBU) and
 * becomes the hierarchy input (SetHierarchyParams + ACMMP.cpp:788-844).  Leaves the context in hierarchy
 * mode, ready for acmmp_run_patch_match. */
int acmmp_next_level(acmmp_ctx *ctx, int n, const float *const *images, const int32_t *widths,
                     const int32_t *heights, const acmmp_camera *cams);
/* The same with the images already on the context's device (dense float32, as acmmp_set_views_device): a driver that
 * keeps every view's level image on the device once instead of uploading it with every context that uses it. */
int acmmp_next_level_device(acmmp_ctx *ctx, int n, const float *const *images_dev, const int32_t *widths,
                            const int32_t *heights, const acmmp_camera *cams);
/* Pinned host buffers holding the result of the last acmmp_run_patch_match (W*H float4 planes, W*H float
 * costs); valid until the next run or re-configuration.  Saves the copy acmmp_get_result makes. */
int acmmp_result_host(acmmp_ctx *ctx, const float **planes4, const float **costs);

/* Replaces ACMMP::CudaPlanarPriorInitialization (ACMMP.cpp:847-867): plane_params = n_planes x
 * float4 (camera-frame normal, d); masks = W*H float, 1-based triangle id or 0.  Also sets
 * planar_prior like SetPlanarPriorParams. */
int acmmp_set_planar_prior_inputs(acmmp_ctx *ctx, const float *plane_params4, int n_planes, const float *masks);

/* The planar-prior stage on the device.  The reference runs it on the CPU between the
 * photometric and the prior stage; with these two calls only the Delaunay triangulation stays on the host.
 *
 * acmmp_support_points replaces ACMMP::GetSupportPoints (ACMMP.cpp:904-930) on the costs of the state on the device
 * (after acmmp_run_patch_match[_resident]): per 5x5 cell the pixel of least cost if that cost is below 0.1.  xy
 * receives (x, y) int32 pairs in the reference's order (cell columns outer, rows inner), *n their number;
 * capacity is in points.
 *
 * acmmp_planar_prior_from_triangles replaces the rest of main.cpp:113-185 + CudaPlanarPriorInitialization
 * (ACMMP.cpp:847-867): tri_xy holds n_tri triangles, six int32 each (x1 y1 x2 y2 x3 y3, every vertex inside the
 * image), in the order their 1-based ids are assigned.  Per triangle the plane through the three lifted vertices
 * (GetPriorPlaneParams, ACMMP.cpp:956-989; depths = the state on the device) and the reference's barycentric
 * stepping rasteriser (main.cpp:153-159; a later triangle overwrites an earlier one); per pixel the depth-range
 * test of main.cpp:168-181 (GetDepthFromPlaneParam, ACMMP.cpp:991-1011).  Leaves the prior inputs on the device and
 * sets planar_prior like SetPlanarPriorParams.  PINHOLE results are bit-identical to the CPU stage. */
int acmmp_support_points(acmmp_ctx *ctx, int32_t *xy, int capacity, int *n);
int acmmp_planar_prior_from_triangles(acmmp_ctx *ctx, const int32_t *tri_xy, int n_tri);
/* What the device holds after acmmp_set_planar_prior_inputs / acmmp_planar_prior_from_triangles: per pixel the prior
 * plane (n_cam, d) and the 1-based triangle id (0 = none) -- prior_planes_host / plane_masks_host of ACMMP.cpp:847-867. */
int acmmp_download_prior(acmmp_ctx *ctx, float *prior_planes4, uint32_t *plane_masks);

/* The reference seeds cuRAND XORWOW with clock64() per thread (ACMMP.cu:684); here the seed is
 * explicit: state(pixel) = curand_init(seed, subsequence = y, offset = x). */
int acmmp_set_seed(acmmp_ctx *ctx, uint64_t seed);

/* SPHERE camera model only.  The fork's angular bilateral weight (ACMMP.cu:436-442, :479-486) shrinks with the image
 * height (sigma_eff = 5 pi / H): at 3200x1600 the four nearest window taps weigh 5.6e-7, the next ring 1e-14, the corners
 * 5e-32 -- most of the 36 samples ComputeBilateralNCC (ACMMP.cu:450-495) takes per (hypothesis, view) add terms far below
 * the float32 resolution of its sums.  Tap steps (2x2 blocks of taps) whose weights are all below
 * relative_weight x (sum of the 36 weights) are not sampled.  Default 2^-24; 0 samples everything (bit-identical to the
 * unpruned evaluation).  The reference-side sums and the `sum_bw < 1e-6` exit (ACMMP.cu:497) always use all 36 taps.
 * Measured effect on the cost: see tests/test_gpu_parity.py::test_sphere_tap_pruning_*. */
int acmmp_set_sphere_tap_pruning(acmmp_ctx *ctx, float relative_weight);

/* 0: plane_now is the current plane unless a neighbour is accepted (what the source intends);
 * 1 (default): what the reference BINARY does -- `float4 plane_hypotheses_now` is uninitialised
 * (ACMMP.cu:1301) and nvcc 12.9 keeps the best neighbour's plane in it whenever that neighbour
 * exists, accepted or not.  See DESIGN.md "undefined behaviour". */
int acmmp_set_plane_now_semantics(acmmp_ctx *ctx, int as_compiled);

/* Replaces ACMMP::RunPatchMatch (ACMMP.cu:1506-1556): init, max_iterations x (black, red),
 * depth/normal extraction, two median passes, device->host copy of planes and costs.
 * Asynchronous on the context's stream up to the final copy, which it waits for. */
int acmmp_run_patch_match(acmmp_ctx *ctx);
/* The same stage without the device->host copy: the result stays on the device for the next stage of the
 * GPU-resident chain (acmmp_set_depth_maps with maps[0] == NULL, acmmp_next_level, acmmp_export_depth_device).
 * acmmp_download_result fetches it into the pinned host buffers when the host does need it (what
 * RunPatchMatch's cudaMemcpy at ACMMP.cu:1553-1554 does unconditionally). */
int acmmp_run_patch_match_resident(acmmp_ctx *ctx);
int acmmp_download_result(acmmp_ctx *ctx);

/* The same stages one launch at a time (what RunPatchMatch launches at ACMMP.cu:1534, :1538/:1540,
 * :1545-:1551).  colour 0 = black, 1 = red.  No host synchronisation. */
int acmmp_random_init(acmmp_ctx *ctx);
int acmmp_checkerboard_pass(acmmp_ctx *ctx, int colour, int iter);
int acmmp_finalize(acmmp_ctx *ctx);
int acmmp_synchronize(acmmp_ctx *ctx);

/* Replaces ACMMP::GetPlaneHypothesis / GetCost (ACMMP.cpp:884-892) in bulk: copies the host
 * result of the last acmmp_run_patch_match. planes4: W*H*4 floats (world normal, depth). */
int acmmp_get_result(acmmp_ctx *ctx, float *planes4, float *costs);
int acmmp_width(const acmmp_ctx *ctx);
int acmmp_height(const acmmp_ctx *ctx);

/* Device-resident results for chaining stages without host round trips (valid until the context
 * is destroyed or re-configured): float4 planes, float costs. */
int acmmp_device_buffers(acmmp_ctx *ctx, void **planes4_dev, void **costs_dev);
/* Pack plane.w (depth) of the current state into a dense float32 device map (what a neighbour
 * needs for geometric consistency; what depths*.dmb holds).
 * ASYNCHRONOUS: the copy kernel is only enqueued on the context's private (non-blocking) stream.  A consumer on
 * any other stream or context -- acmmp_set_depth_maps_device of another view, a collective -- must call
 * acmmp_synchronize(ctx) (or acmmp_export_depth_device_sync) first. */
int acmmp_export_depth_device(acmmp_ctx *ctx, float *depth_dev);
/* The same, returning after the map is complete (stream-synchronised). */
int acmmp_export_depth_device_sync(acmmp_ctx *ctx, float *depth_dev);

/* Raw device state <-> host (tests, stage chaining).  Any pointer may be NULL.
 * rand6: 6 x uint32 per pixel = XORWOW {d, v[0..4]}. */
int acmmp_download_state(acmmp_ctx *ctx, float *planes4, float *costs, uint32_t *selected_views,
                         uint32_t *rand6, float *pre_costs);
int acmmp_upload_state(acmmp_ctx *ctx, const float *planes4, const float *costs, const uint32_t *selected_views,
                       const uint32_t *rand6, const float *pre_costs);

/* Replaces RunJBU / JBU::CudaRun / JBU_cu (ACMMP.cpp:1071-1122, ACMMP.cu:1558-1649) minus the file
 * write: joint-bilateral upsampling of a coarse depth map (sw x sh) guided by the fine grey image
 * (w x h).  Host pointers.  Returns ACMMP_E_ARG when max(h/sh, w/sw) == 1 (the reference returns
 * without output, ACMMP.cpp:1077-1080). */
int acmmp_jbu(int device, const float *image, int w, int h, const float *coarse_depth, int sw, int sh,
              float *out_depth);
int acmmp_jbu_device(int device, const float *image_dev, int w, int h, const float *coarse_depth_dev, int sw,
                     int sh, float *out_depth_dev, void *cuda_stream);
/* CUDA-event time (ms) of the JBU kernel launched by the last acmmp_jbu call of this thread. */
float acmmp_last_jbu_ms(void);

/* Deterministic sub-kernel probes (parity tests): evaluate, for every pixel p of the reference
 * view and a caller-supplied per-pixel plane (camera-frame normal, d), with the SAME device code
 * the checkerboard and initialisation kernels run (ncc / initcost: quad_ncc, four lanes per pixel):
 *   ncc      : ComputeBilateralNCC against source view `view` (1-based)      (ACMMP.cu:405-516)
 *   geom     : ComputeGeomConsistencyCost against depth map `view`            (ACMMP.cu:646-671)
 *   warp     : (src x, src y, src depth, ref depth) of p under the plane       (ACMMP.cu:187, :565, :602)
 *   initcost : ComputeMultiViewInitialCostandSelectedViews                     (ACMMP.cu:519-556)
 * Host pointers; planes4 and out4 are W*H*4 floats. */
int acmmp_probe_ncc(acmmp_ctx *ctx, const float *planes4, int view, float *out);
/* the 36 fetch coordinates of that form (texel-centre shift included): out72 = W*H*72 floats, (u, v) of tap
 * k = ii*6 + jj, i = 2 ii - 5, j = 2 jj - 5 -- the reference's sample loop order, ACMMP.cu:450-476.
 * variant 0: one hypothesis at a time; 1 / 2 (SPHERE): the packed two-hypothesis form of the pass with this plane
 * as its first / second hypothesis (same bits expected) */
int acmmp_probe_coords(acmmp_ctx *ctx, const float *planes4, int view, int variant, float *out72);
int acmmp_probe_geom(acmmp_ctx *ctx, const float *planes4, int view, float *out);
int acmmp_probe_warp(acmmp_ctx *ctx, const float *planes4, int view, float *out4);
int acmmp_probe_initcost(acmmp_ctx *ctx, const float *planes4, float *out, uint32_t *selected_views);

/* Timing of the last launches on the context's stream, CUDA events, milliseconds:
 * what[0]=random_init, [1]=sum of checkerboard passes, [2]=finalize, [3]=number of passes,
 * [4]=last pass, [5]=JBU kernel of the last acmmp_next_level.  Valid after acmmp_synchronize /
 * acmmp_run_patch_match. */
int acmmp_last_timings(acmmp_ctx *ctx, float what[8]);

/* How many kernels of this library were launched on the context since creation. */
int64_t acmmp_launch_count(const acmmp_ctx *ctx);

/* ---------------------------------------------------------------------------------------------
 * Depth-map fusion.  Replaces the device part of RunFusionCuda
 * (ACMMP.cu:1817-2105): SimpleFusionKernel (:1664-1814) per reference view, the valid points
 * compacted on the device in pixel order (the reference copies a 36-byte PointList + a flag for every
 * pixel to the host and filters there, :2056-2076).
 * --------------------------------------------------------------------------------------------- */
/* reference `struct PointList`, main.h:71-75: world point, world normal, colour (written as b, g, r by the PLY writer) */
typedef struct {
    float coord[3];
    float normal[3];
    float color[3];
} acmmp_point;

typedef struct acmmp_fusion acmmp_fusion;

int acmmp_fusion_create(int device, int n_views, acmmp_fusion **out);
void acmmp_fusion_destroy(acmmp_fusion *f);
const char *acmmp_fusion_last_error(const acmmp_fusion *f);
/* One view of the scene: camera ALREADY scaled to the depth map's size (RescaleImageAndCamera,
 * ACMMP.cpp:213-245), depth (w*h, depths_geom.dmb / depths.dmb), world-frame normals (w*h*3, normals.dmb)
 * and grey levels 0..255 (w*h) at that size.  Host pointers; copied to the device. */
int acmmp_fusion_set_view(acmmp_fusion *f, int index, const acmmp_camera *cam, int w, int h, const float *depth,
                          const float *normals3, const float *gray);
/* Optional, after acmmp_fusion_set_view[_device] of that view: its colour image at the depth map's size, w*h*3 bytes in
 * OpenCV's order (B, G, R per pixel: cv::imread(IMREAD_COLOR) + RescaleImageAndCamera, ACMMP.cu:1862-1872).  Without it
 * the fused points carry the grey level in all three channels.  acmmp_point::color is (B, G, R) like PointList::color
 * (ACMMP.cu:1704-1708; the PLY writer swaps back, ACMMP.cpp:509-511). */
int acmmp_fusion_set_view_colour(acmmp_fusion *f, int index, const uint8_t *bgr, int w, int h);
/* The same with DEVICE pointers that stay valid until the object is destroyed (resident chain: the depth /
 * normal maps a PatchMatch context holds, acmmp_device_buffers): normals4 = float4 per pixel. */
int acmmp_fusion_set_view_device(acmmp_fusion *f, int index, const acmmp_camera *cam, int w, int h, const float *depth_dev,
                                 const void *normals4_dev, const float *gray_dev);
/* Fuse reference view `ref` against the views src[0 .. n_src) (indices into the view table, -1 = not
 * available; at most 32 like FusionProblem, ACMMP.cu:1656-1661).  points: room for `capacity` points (host);
 * *n_points: how many the view produced (ACMMP_E_ARG when capacity was too small: call again).
 * kernel_ms (optional): CUDA-event time of the three kernels. */
int acmmp_fusion_run(acmmp_fusion *f, int ref, int n_src, const int32_t *src, acmmp_point *points, int capacity,
                     int *n_points, float *kernel_ms);
/* The same, with the points already in the PLY file's 27-byte vertex records (x y z nx ny nz as little-endian floats, then
 * red green blue; what StoreColorPlyFileBinaryPointCloud writes per point, ACMMP.cpp:481-534, including the non-finite
 * coordinate -> origin rule): the host appends `records27` to the file as it is.  capacity in points. */
int acmmp_fusion_run_ply(acmmp_fusion *f, int ref, int n_src, const int32_t *src, uint8_t *records27, int capacity,
                         int *n_points, float *kernel_ms);
/* which pixels of the reference view fused LAST produced a point: w*h bytes (host), row-major */
int acmmp_fusion_last_flags(acmmp_fusion *f, int ref, unsigned char *flags);

#ifdef __cplusplus
}
#endif
#endif /* ACMMP_B200_H_ */
