// acmmp_main.cpp -- `acmmp_b200 dense_folder [--seed S] [--device D] [--gpus N] [--max-views N] [--resident 1]`: the reference's pipeline
// schedule (main.cpp:392-482) on top of the B200 library, stage by stage through the same .dmb files:
//   per pyramid level (coarsest first):
//     [level > 0] JBU of depths_geom.dmb -> depths.dmb, then photometric stage with hierarchy
//     photometric stage -> CPU planar prior -> prior stage (same object)            -> depths.dmb
//     2 x geometric-consistency stage (the second one with multi_geometry)          -> depths_geom.dmb
// `--resident 1` runs the same stages GPU-resident (RunResident below): no .dmb round trips between stages, every
// image read once per level; it writes the same final maps.  `--gpus N` deals the reference views to N devices (one host
// thread each) and exchanges the depth maps with peer copies at the two exchange points of a level.
// Fusion (RunFusionCuda, main.cpp:478-479) follows the last level (`--fusion 0` skips it): ACMMP/ACMM_model_cuda_5.ply.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <sstream>
#include <thread>
#include <stdexcept>
#include <sys/stat.h>
#include <sys/types.h>

#include <cuda_runtime.h>

#include "acmmp_host.h"

namespace {

uint64_t g_seed = 0;
int g_device = 0;
int g_gpu_prior = 0;
double g_gpu_ms = 0.0, g_prior_s = 0.0;
// wall-clock attribution of the resident schedule (printed in the summary): image load + scaling, view upload / level
// change (incl. context creation), stage runs (kernels + waits + downloads), depth-map export, result output
double g_t_barrier = 0.0, g_t_setup = 0.0, g_t_ctx = 0.0, g_t_upload = 0.0, g_t_support = 0.0, g_t_prior_dev = 0.0;
double g_t_load = 0.0, g_t_views = 0.0, g_t_run = 0.0, g_t_export = 0.0, g_t_output = 0.0, g_t_join = 0.0;
double g_t_sweep1 = 0.0, g_t_geom = 0.0;   // totals: first sweep, geometric sweeps
double g_t_exchange = 0.0;                 // --gpus N: peer copies of the depth maps
int g_gpus = 1, g_devices_used = 1;
int g_fusion = 1;                          // RunFusionCuda after the last level, like main.cpp:478-479

std::mutex g_prior_mutex;
void add_prior_s(double dt)      // the CPU prior stages of two views may overlap on worker threads
{
    std::lock_guard<std::mutex> lock(g_prior_mutex);
    g_prior_s += dt;
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

std::string result_folder_of(const std::string &dense_folder, int ref_id)
{
    std::stringstream s;
    s << dense_folder << "/ACMMP/2333_" << std::setw(8) << std::setfill('0') << ref_id;
    return s.str();
}

// main.cpp:35-71: per view max_size = min(max(rows, cols), 3200), k = number of halvings until <= 1000
int ComputeMultiScaleSettings(const std::string &dense_folder, std::vector<Problem> &problems)
{
    int max_num_downscale = -1;
    const int size_bound = 1000;
    PatchMatchParams pmp;
    acmmp_default_params(&pmp);
    for (auto &problem : problems) {
        int cols = 0, rows = 0;
        if (!ImageSize(dense_folder, problem.ref_image_id, cols, rows)) std::cerr << "Error: no image for view " << problem.ref_image_id << std::endl;
        int max_size = std::max(rows, cols);
        if (max_size > pmp.max_image_size) max_size = pmp.max_image_size;
        problem.max_image_size = max_size;
        int k = 0;
        while (max_size > size_bound) {
            max_size /= 2;
            k++;
        }
        if (k > max_num_downscale) max_num_downscale = k;
        problem.num_downscale = k;
    }
    return max_num_downscale;
}

void collect(ACMMP &acmmp, cv::Mat_<float> &depths, cv::Mat_<cv::Vec3f> &normals, cv::Mat_<float> &costs)
{
    const int width = acmmp.GetReferenceImageWidth(), height = acmmp.GetReferenceImageHeight();
    for (int row = 0; row < height; ++row) {
        for (int col = 0; col < width; ++col) {
            const int center = row * width + col;
            const float4 ph = acmmp.GetPlaneHypothesis(center);
            depths(row, col) = ph.w;
            normals(row, col) = cv::Vec3f(ph.x, ph.y, ph.z);
            costs(row, col) = acmmp.GetCost(center);
        }
    }
    float t[8];
    acmmp.GetTimings(t);
    g_gpu_ms += t[0] + t[1] + t[2];
}

// The CPU planar-prior stage, main.cpp:113-185: support points -> Delaunay triangles -> triangle-id mask
// (the reference's barycentric stepping rasteriser) + one plane per triangle -> pixels whose prior depth leaves the
// depth range dropped.  `depths` is the photometric stage's depth map.
void PlanarPriorStage(ACMMP &acmmp, const cv::Mat_<float> &depths, cv::Mat_<float> &mask_tri, std::vector<float4> &planeParams_tri)
{
    const double t0 = now_s();
    acmmp.SetPlanarPriorParams();
    const int npx = acmmp.GetReferenceImageWidth() * acmmp.GetReferenceImageHeight();
    std::vector<float> costs((size_t)npx);
    for (int k = 0; k < npx; ++k) costs[k] = acmmp.GetCost(k);
    PlanarPriorCpu(acmmp.GetReferenceCamera(), depths, costs.data(), acmmp.GetMinDepth(), acmmp.GetMaxDepth(), mask_tri, planeParams_tri);
    add_prior_s(now_s() - t0);
}

// The same stage with everything but the triangulation on the device: support points
// from the costs on the device, Delaunay on the host, then plane fit / rasteriser / depth-range test / prior upload in
// one library call.  Ends where PlanarPriorStage + CudaPlanarPriorInitialization end.
void PlanarPriorStageGpu(ACMMP &acmmp)
{
    const double t0 = now_s();
    const int width = acmmp.GetReferenceImageWidth(), height = acmmp.GetReferenceImageHeight();
    acmmp.SetPlanarPriorParams();
    const cv::Rect imageRC(0, 0, width, height);
    std::vector<cv::Point> support2DPoints;
    acmmp.GetSupportPointsDevice(support2DPoints);
    const auto triangles = acmmp.DelaunayTriangulation(imageRC, support2DPoints);
    std::vector<Triangle> inside;
    inside.reserve(triangles.size());
    for (const auto &triangle : triangles)
        if (imageRC.contains(triangle.pt1) && imageRC.contains(triangle.pt2) && imageRC.contains(triangle.pt3)) inside.push_back(triangle);
    acmmp.CudaPlanarPriorFromTriangles(inside);
    add_prior_s(now_s() - t0);
}

// main.cpp:73-210
void ProcessProblem(const std::string &dense_folder, const std::vector<Problem> &problems, const int idx, bool geom_consistency,
                    bool planar_prior, bool hierarchy, bool multi_geometry = false)
{
    const Problem &problem = problems[idx];
    std::cout << "Processing image " << std::setw(8) << std::setfill('0') << problem.ref_image_id << "..." << std::endl;
    const std::string result_folder = result_folder_of(dense_folder, problem.ref_image_id);
    mkdir(result_folder.c_str(), 0777);

    ACMMP acmmp(g_device);
    acmmp.SetSeed(g_seed);
    if (geom_consistency) acmmp.SetGeomConsistencyParams(multi_geometry);
    if (hierarchy) acmmp.SetHierarchyParams();
    acmmp.InuputInitialization(dense_folder, problems, idx);
    acmmp.CudaSpaceInitialization(dense_folder, problem);
    acmmp.RunPatchMatch();

    const int width = acmmp.GetReferenceImageWidth(), height = acmmp.GetReferenceImageHeight();
    cv::Mat_<float> depths = cv::Mat_<float>::zeros(height, width);
    cv::Mat_<cv::Vec3f> normals = cv::Mat_<cv::Vec3f>::zeros(height, width);
    cv::Mat_<float> costs = cv::Mat_<float>::zeros(height, width);
    collect(acmmp, depths, normals, costs);

    if (planar_prior) {                                     // main.cpp:113-197
        std::cout << "Run Planar Prior Assisted PatchMatch MVS ..." << std::endl;
        if (g_gpu_prior) {
            PlanarPriorStageGpu(acmmp);
        } else {
            cv::Mat_<float> mask_tri;
            std::vector<float4> planeParams_tri;
            PlanarPriorStage(acmmp, depths, mask_tri, planeParams_tri);
            acmmp.CudaPlanarPriorInitialization(planeParams_tri, mask_tri);
        }
        acmmp.RunPatchMatch();
        collect(acmmp, depths, normals, costs);
    }

    const std::string suffix = geom_consistency ? "/depths_geom.dmb" : "/depths.dmb";
    writeDepthDmb(result_folder + suffix, depths);
    writeNormalDmb(result_folder + "/normals.dmb", normals);
    writeDepthDmb(result_folder + "/costs.dmb", costs);
    std::cout << "Processing image " << std::setw(8) << std::setfill('0') << problem.ref_image_id << " done!" << std::endl;
}

// main.cpp:212-238
void JointBilateralUpsampling(const std::string &dense_folder, const Problem &problem, int acmmp_size)
{
    cv::Mat_<float> ref_depth;
    readDepthDmb(result_folder_of(dense_folder, problem.ref_image_id) + "/depths_geom.dmb", ref_depth);
    cv::Mat_<float> image_float;
    if (!LoadGreyImage(dense_folder, problem.ref_image_id, image_float)) return;
    const float factor_x = static_cast<float>(acmmp_size) / image_float.cols;
    const float factor_y = static_cast<float>(acmmp_size) / image_float.rows;
    const float factor = std::min(factor_x, factor_y);
    const int new_cols = (int)std::round(image_float.cols * factor);
    const int new_rows = (int)std::round(image_float.rows * factor);
    cv::Mat_<float> scaled;
    ResizeLinear(image_float, scaled, new_cols, new_rows);
    std::cout << "Run JBU for image " << problem.ref_image_id << ".jpg" << std::endl;
    RunJBU(scaled, ref_depth, dense_folder, problem, g_device);
}

// ---- GPU-resident schedule ---------------------------------------------------------
// The same stages in the same order as the file-chained loop in main() -- per level: photometric (+ hierarchy) and
// prior stage for every view, then two geometric sweeps over the views -- but one ACMMP object per view lives through
// the whole run: a stage finds the previous stage's planes / costs on the device, the next level is reached through
// the on-device JBU, the neighbours' depth maps are device buffers (`dmap` = what depths.dmb would hold, `gmap` =
// depths_geom.dmb, overwritten view by view exactly like the file: the second geometric sweep is Gauss-Seidel across
// views in the reference, main.cpp:443-445, and here).  Every image is read and scaled once per level instead of
// once per (view that uses it, stage).  Only the finest level's maps are written, at the end.
// Returns false (nothing done) when the scene does not fit the scheme; the caller then runs the file-chained schedule.
struct DeviceMap {
    float *ptr = nullptr;
    int w = 0, h = 0;
    size_t cap = 0;
    // `reserve`: pixels of the finest level, so that the buffer is allocated once (a cudaFree + cudaMalloc per level
    // cost 20 ms per export on a device that holds a few GB of pooled buffers)
    // The memory comes out of the library's pool of the device (acmmp_pool_alloc): with the reservation made at the start
    // of the resident schedule it is a pointer bump.
    void fit(int nw, int nh, size_t reserve, int device)
    {
        const size_t need = std::max((size_t)nw * nh, reserve);
        if (need > cap) {
            if (ptr) acmmp_pool_free(device, ptr);
            void *p = nullptr;
            if (acmmp_pool_alloc(device, need * sizeof(float), &p) != ACMMP_OK) throw std::runtime_error("device memory for a depth map / level image");
            ptr = (float *)p;
            cap = need;
        }
        w = nw;
        h = nh;
    }
};

// Rendezvous of the per-device host threads; abort() releases everybody when one of them failed.
class PhaseBarrier {
public:
    explicit PhaseBarrier(int n) : n_(n) {}
    bool wait()                                           // false: the run was aborted
    {
        std::unique_lock<std::mutex> lock(m_);
        if (aborted_) return false;
        const int gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lock, [&] { return gen != gen_ || aborted_; });
        }
        return !aborted_;
    }
    void abort()
    {
        std::lock_guard<std::mutex> lock(m_);
        aborted_ = true;
        cv_.notify_all();
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
    bool aborted_ = false;
};

// Multi-GPU: the reference views are dealt round-robin to `ndev` devices, one host thread per
// device issues that device's work.  Views are independent inside a stage (main.cpp:431-446); the only cross-view input
// is the neighbours' depth maps of the geometric stages, so every device keeps a table with the current map of EVERY
// view and pulls the maps other devices own at the two exchange points of a level (after the prior stage, after the first
// geometric round) with peer copies over NVLink -- what the reference passes through depths.dmb / depths_geom.dmb.
// Inside a device the second geometric round stays Gauss-Seidel (a view reads the maps its own device rewrote earlier in
// the round), across devices it is Jacobi (the maps of the first round); one device = the reference's order exactly.
bool RunResident(const std::string &dense_folder, std::vector<Problem> &problems, const size_t num_images, int max_num_downscale, int ndev,
                 std::vector<ResidentView> *resident_out)
{
    std::map<int, int> index_of;                      // image id -> position in `problems`
    for (size_t i = 0; i < problems.size(); ++i) index_of[problems[i].ref_image_id] = (int)i;
    for (size_t i = 0; i < num_images; ++i) {
        if (problems[i].num_downscale != max_num_downscale) return false;       // views with fewer levels: JBU no-op case
        for (int id : problems[i].src_image_ids) {
            const auto it = index_of.find(id);
            if (it == index_of.end() || (size_t)it->second >= num_images) return false;    // a neighbour that is not processed
        }
    }
    const double t_setup0 = now_s();
    // the coarsest level's images of all views: read, decoded and scaled on a host thread while CUDA comes up below
    std::vector<cv::Mat_<float>> level_image(num_images), next_image(num_images);
    std::vector<Camera> level_camera(num_images), next_camera(num_images);
    std::shared_future<void> first_load = std::async(std::launch::async, [&]() {
        for (size_t i = 0; i < num_images; ++i)
            LoadScaledView(dense_folder, problems[i].ref_image_id, problems[i].max_image_size / (1 << max_num_downscale), next_image[i], next_camera[i]);
    }).share();
    int avail = 0;
    cudaGetDeviceCount(&avail);
    if (ndev < 1 || g_device + ndev > avail) {
        std::cout << "resident schedule: " << ndev << " devices from device " << g_device << " requested, " << avail << " present" << std::endl;
        return false;
    }
    ndev = (int)std::min<size_t>((size_t)ndev, std::max<size_t>(num_images, 1));
    const auto device_of = [&](size_t view) { return (int)(view % (size_t)ndev); };
    // image sizes: once (headers only), not per device
    std::vector<double> level_px(num_images, 0.0);
    std::vector<size_t> full_px(num_images, 0);         // upper bound of a view's pixel count at any level
    std::vector<size_t> finest_px(num_images, 0);       // exact pixel count of the finest level (LoadScaledView's arithmetic)
    for (size_t i = 0; i < num_images; ++i) {
        int cols = 0, rows = 0;
        if (!ImageSize(dense_folder, problems[i].ref_image_id, cols, rows)) continue;
        const double scale = std::min(1.0, (double)problems[i].max_image_size / std::max(cols, rows));
        level_px[i] = (double)cols * rows * scale * scale;
        full_px[i] = (size_t)cols * rows;
        const int m = problems[i].max_image_size;
        if (cols <= m && rows <= m) {
            finest_px[i] = (size_t)cols * rows;
        } else {
            const float factor = std::min(static_cast<float>(m) / cols, static_cast<float>(m) / rows);
            finest_px[i] = (size_t)std::round(cols * factor) * (size_t)std::round(rows * factor);
        }
    }
    // What a device holds: per owned view the stage state of the current level (planes, costs, pre-costs: 24 bytes
    // per pixel -- parked views keep nothing else); the level images of the views it needs (4 B/px); two depth-map tables
    // of all views (8 B/px); and the scratch of the views in flight, which all views share through the device's pool
    // (images twice -- layered + per-view textures --, padded reference, ping-pong planes and costs, view masks, two RNG
    // states, prior planes + masks, pinned-result staging: 8 (n_src + 1) + 160 B/px, two sets in rotation plus the
    // smaller sets of the coarser levels that stay in the pool).  Scenes that do not fit fall back to the file-chained
    // schedule.
    std::vector<size_t> reserve_bytes(ndev, 0);
    for (int d = 0; d < ndev; ++d) {
        double need = 0.0, scratch = 0.0;
        std::vector<char> uses(num_images, 0);
        for (size_t i = 0; i < num_images; ++i)
            if (device_of(i) == d) {
                uses[i] = 1;
                for (int id : problems[i].src_image_ids) uses[index_of.at(id)] = 1;
            }
        for (size_t i = 0; i < num_images; ++i) {
            const double px = level_px[i];
            if (device_of(i) == d) {
                need += px * 24.0;
                scratch = std::max(scratch, px * (8.0 * (problems[i].src_image_ids.size() + 1) + 160.0));
            }
            if (uses[i]) need += px * 4.0;
            need += px * 8.0;
        }
        need += 2.7 * scratch;
        reserve_bytes[d] = (size_t)(need * 1.1) + ((size_t)256 << 20);
    }
    // The devices come up side by side (a CUDA context takes 1 - 2.5 s on these machines; one after the other that was 9.4 s
    // of a 4-device run): context, memory check, and the one allocation per device that the library's pool carves the state
    // blocks, scratch sets, level images and depth-map tables out of (texture arrays and pinned buffers are separate
    // allocations; what does not fit falls back to cudaMalloc).
    {
        std::vector<int> fits(ndev, 1);
        std::vector<std::thread> starters;
        for (int d = 0; d < ndev; ++d)
            starters.emplace_back([&, d]() {
                size_t free_b = 0, total_b = 0;
                if (cudaSetDevice(g_device + d) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { fits[d] = 0; return; }
                if ((double)reserve_bytes[d] > 0.85 * (double)free_b) {
                    std::cout << "resident schedule: device " << g_device + d << " needs about " << reserve_bytes[d] / 1e9 << " GB of device memory, "
                              << free_b / 1e9 << " GB are free" << std::endl;
                    fits[d] = 0;
                    return;
                }
                if (acmmp_reserve_device_memory(g_device + d, reserve_bytes[d]) != ACMMP_OK)
                    std::cout << "resident schedule: no " << reserve_bytes[d] / 1e9 << " GB reservation on device " << g_device + d << " (allocating block by block)" << std::endl;
                // the page-locked result buffers of the final downloads (one pair per size: a writer thread hands them back
                // after one copy pass, long before the device finishes its next view)
                std::vector<size_t> sizes;
                for (size_t i = 0; i < num_images; ++i)
                    if (device_of(i) == d && finest_px[i] && std::find(sizes.begin(), sizes.end(), finest_px[i]) == sizes.end()) sizes.push_back(finest_px[i]);
                for (size_t npx : sizes) {
                    acmmp_reserve_pinned(g_device + d, 16 * npx);
                    acmmp_reserve_pinned(g_device + d, 4 * npx);
                }
            });
        for (auto &t : starters) t.join();
        for (int d = 0; d < ndev; ++d)
            if (!fits[d]) return false;
    }
    if (ndev > 1) {
        for (int a = 0; a < ndev; ++a)
            for (int b = 0; b < ndev; ++b)
                if (a != b) {
                    cudaSetDevice(g_device + a);
                    if (cudaDeviceEnablePeerAccess(g_device + b, 0) != cudaSuccess) (void)cudaGetLastError();      // copies then stage through the host
                }
        std::cout << "resident schedule on " << ndev << " devices, views dealt round-robin" << std::endl;
    }

    g_t_setup = now_s() - t_setup0;                    // CUDA initialisation (all GPUs of the box are enumerated), contexts, peer access
    std::vector<std::unique_ptr<ACMMP>> objs(num_images);
    // tab[d][v]: the map of view v on device d ("d" = what depths.dmb would hold, "g" = depths_geom.dmb)
    std::vector<std::vector<DeviceMap>> dtab(ndev, std::vector<DeviceMap>(num_images)), gtab(ndev, std::vector<DeviceMap>(num_images));
    std::vector<cv::Mat_<float>> final_prior_depth(num_images);
    // which views does device d need the maps of (its own views' source views that another device owns)
    std::vector<std::vector<size_t>> remote(ndev);
    for (int d = 0; d < ndev; ++d) {
        std::vector<char> need(num_images, 0);
        for (size_t i = 0; i < num_images; ++i)
            if (device_of(i) == d)
                for (int id : problems[i].src_image_ids) need[index_of.at(id)] = 1;
        for (size_t v = 0; v < num_images; ++v)
            if (need[v] && device_of(v) != d) remote[d].push_back(v);
    }
    std::vector<int> level_w(num_images, 0), level_h(num_images, 0);
    std::vector<std::vector<DeviceMap>> pools(ndev, std::vector<DeviceMap>(num_images));
    PhaseBarrier barrier(ndev);
    std::mutex stats_mutex;
    std::string failure;
    const int levels = max_num_downscale + 1;

    auto device_main = [&](const int d) {
        // per-thread wall-clock attribution, merged into the globals at the end
        double t_ctx = 0, t_upload = 0, t_support = 0, t_prior_dev = 0, t_barrier = 0;
        auto timed_wait = [&]() {                           // time this thread waits for the other devices' threads
            const double t0 = now_s();
            const bool ok = barrier.wait();
            t_barrier += now_s() - t0;
            return ok;
        };
        double t_load = 0, t_views = 0, t_run = 0, t_export = 0, t_output = 0, t_join = 0, t_sweep1 = 0, t_geom = 0, t_exchange = 0, gpu_ms = 0;
        cudaSetDevice(g_device + d);
        std::vector<size_t> mine;
        for (size_t i = 0; i < num_images; ++i)
            if (device_of(i) == d) mine.push_back(i);
        auto pull = [&](std::vector<std::vector<DeviceMap>> &tab) {
            // the maps other devices own, straight from their tables into mine (NVLink peer copies)
            const double t0 = now_s();
            for (size_t v : remote[d]) {
                const int o = device_of(v);
                tab[d][v].fit(level_w[v], level_h[v], full_px[v], g_device + d);
                if (cudaMemcpyPeerAsync(tab[d][v].ptr, g_device + d, tab[o][v].ptr, g_device + o, sizeof(float) * (size_t)level_w[v] * level_h[v], 0) != cudaSuccess)
                    throw std::runtime_error("peer copy of a depth map failed");
            }
            if (cudaStreamSynchronize(0) != cudaSuccess) throw std::runtime_error("peer copies of the depth maps failed");
            t_exchange += now_s() - t0;
        };
        std::vector<std::future<void>> writers;             // .dmb output of finished views
        std::future<void> prefetch;                         // the next level's images of this thread's views
        std::vector<DeviceMap> &pool = pools[d];            // level images on this device
        std::vector<size_t> needed;                         // my views and their source views
        {
            std::vector<char> need(num_images, 0);
            for (size_t i : mine) {
                need[i] = 1;
                for (int id : problems[i].src_image_ids) need[index_of.at(id)] = 1;
            }
            for (size_t v = 0; v < num_images; ++v)
                if (need[v]) needed.push_back(v);
        }
        bool first_level = true;
        for (int level = 0; level < levels; ++level) {
            const int scale = max_num_downscale - level;
            const bool finest = scale == 0;
            if (d == 0) {
                std::cout << "Scale: " << scale << std::endl;
                for (auto &problem : problems) {
                    if (problem.num_downscale >= 0) {
                        problem.cur_image_size = problem.max_image_size / (int)std::pow(2, problem.num_downscale);
                        problem.num_downscale--;
                    }
                }
            }
            if (!timed_wait()) return;
            // every view of this level, read and scaled once (each thread its share, all of them shared afterwards)
            double tp = now_s();
            if (level == 0 || prefetch.valid()) {
                if (level == 0) first_load.get();              // started before CUDA was initialised
                else prefetch.get();                           // read and scaled beside the previous level's kernels
                for (size_t i : mine) {
                    level_image[i] = std::move(next_image[i]);
                    level_camera[i] = next_camera[i];
                }
            } else {
                for (size_t i : mine) LoadScaledView(dense_folder, problems[i].ref_image_id, problems[i].cur_image_size, level_image[i], level_camera[i]);
            }
            for (size_t i : mine) {
                level_w[i] = level_image[i].cols;
                level_h[i] = level_image[i].rows;
            }
            t_load += now_s() - tp;
            if (!timed_wait()) return;
            // Every image this device's views use (their own and their source views'), on the device ONCE per level: a view's
            // image is a source image of ~10 other views, and uploading it with each of them was 0.3 s per 3200x2130 view
            tp = now_s();
            for (size_t v : needed) {
                pool[v].fit(level_w[v], level_h[v], full_px[v], g_device + d);
                if (cudaMemcpyAsync(pool[v].ptr, level_image[v].ptr(), sizeof(float) * (size_t)level_w[v] * level_h[v], cudaMemcpyHostToDevice, 0) != cudaSuccess)
                    throw std::runtime_error("upload of a level image failed");
            }
            if (cudaStreamSynchronize(0) != cudaSuccess) throw std::runtime_error("upload of the level images failed");
            t_views += now_s() - tp;
            t_upload += now_s() - tp;
            if (level + 1 < levels) {
                // the next level's images of this thread's views: file read / decode / bilinear scaling on a host thread
                // while this one drives the device (every view has the same number of levels here, see the top)
                const int next_shift = max_num_downscale - level - 1;
                prefetch = std::async(std::launch::async, [&, next_shift]() {
                    for (size_t i : mine)
                        LoadScaledView(dense_folder, problems[i].ref_image_id, problems[i].max_image_size / (1 << next_shift), next_image[i], next_camera[i]);
                });
            }

            // (re-)activate view i on its context: its image and its source views' images out of the level pool
            auto activate = [&](const size_t i, const bool next_level, const bool fresh = false) {
                const Problem &problem = problems[i];
                std::vector<const float *> images{pool[i].ptr};
                std::vector<int> ws{level_w[i]}, hs{level_h[i]};
                std::vector<Camera> cameras{level_camera[i]};
                for (int id : problem.src_image_ids) {
                    const size_t s = index_of.at(id);
                    images.push_back(pool[s].ptr);
                    ws.push_back(level_w[s]);
                    hs.push_back(level_h[s]);
                    cameras.push_back(level_camera[s]);
                }
                objs[i]->SetViewsDevice(images, ws, hs, cameras, next_level, fresh ? &level_image[i] : nullptr);
            };
            struct PriorJob {
                std::future<void> done;
                std::vector<cv::Point> support;
                std::vector<Triangle> inside;                  // --gpu-prior 1: triangles inside the image, id order
                cv::Mat_<float> mask_tri;                      // --gpu-prior 0: the CPU stage's outputs
                std::vector<float4> planeParams_tri;
                double host_s = 0.0;
            };
            std::map<size_t, std::unique_ptr<PriorJob>> pending;
            auto finish_view = [&](const size_t v) {
                ACMMP &a = *objs[v];
                double tw = now_s();
                pending[v]->done.get();                        // rethrows a worker exception
                t_join += now_s() - tw;
                tw = now_s();
                activate(v, false);                            // parked after its photometric stage
                t_views += now_s() - tw;
                tw = now_s();
                if (g_gpu_prior) {
                    a.CudaPlanarPriorFromTriangles(pending[v]->inside);
                    t_prior_dev += now_s() - tw;
                    add_prior_s(pending[v]->host_s + (now_s() - tw));
                } else {
                    a.CudaPlanarPriorInitialization(pending[v]->planeParams_tri, pending[v]->mask_tri);
                }
                pending.erase(v);
                const int width = a.GetReferenceImageWidth(), height = a.GetReferenceImageHeight();
                tw = now_s();
                a.RunPatchMatchResident(false);
                t_run += now_s() - tw;
                float tt[8];
                a.GetTimings(tt);
                gpu_ms += tt[0] + tt[1] + tt[2];
                tw = now_s();
                dtab[d][v].fit(width, height, full_px[v], g_device + d);
                a.ExportDepthDevice(dtab[d][v].ptr);
                a.Park();                                      // the next view's stage runs in this one's scratch
                t_export += now_s() - tw;
                tw = now_s();
                if (finest) {
                    // depths.dmb is an output at the finest level: the exported map (27 MB) instead of the whole result (136 MB)
                    final_prior_depth[v] = cv::Mat_<float>(height, width);
                    if (cudaMemcpy(final_prior_depth[v].ptr(), dtab[d][v].ptr, sizeof(float) * (size_t)width * height, cudaMemcpyDeviceToHost) != cudaSuccess)
                        throw std::runtime_error("download of a prior-stage depth map failed");
                }
                t_output += now_s() - tw;
            };
            const double t_s1 = now_s();
            size_t previous = (size_t)-1;
            for (size_t i : mine) {                                                  // photometric + prior stage
                const Problem &problem = problems[i];
                std::cout << "Processing image " << std::setw(8) << std::setfill('0') << problem.ref_image_id << "..." << std::endl;
                tp = now_s();
                if (first_level) {
                    objs[i].reset(new ACMMP(g_device + d));
                    objs[i]->SetSeed(g_seed);
                    t_ctx += now_s() - tp;
                }
                ACMMP &acmmp = *objs[i];
                activate(i, !first_level, true);
                t_views += now_s() - tp;
                tp = now_s();
                acmmp.RunPatchMatchResident(!g_gpu_prior);                           // the CPU prior stage reads the result
                t_run += now_s() - tp;
                float t[8];
                acmmp.GetTimings(t);
                gpu_ms += t[0] + t[1] + t[2];
                // The host part of the planar-prior stage of view i (the Delaunay triangulation; with --gpu-prior 0 the whole
                // CPU stage) runs on a worker thread while this thread -- which issues ALL work of its device, in a fixed
                // order -- goes on with the photometric stage of its next view; view i is finished (prior upload, prior-stage
                // PatchMatch, depth export) one iteration later.  (Device work from two threads would queue behind each
                // other's persistent kernels: a k_pass launch holds every SM for its ~20 ms.)
                pending[i].reset(new PriorJob());
                PriorJob *job = pending[i].get();
                ACMMP *obj = objs[i].get();
                if (g_gpu_prior) {
                    const double t0 = now_s();
                    acmmp.SetPlanarPriorParams();
                    acmmp.GetSupportPointsDevice(job->support);
                    t_support += now_s() - t0;
                    add_prior_s(now_s() - t0);
                    job->done = std::async(std::launch::async, [job, obj]() {
                        const double t1 = now_s();
                        const int width = obj->GetReferenceImageWidth(), height = obj->GetReferenceImageHeight();
                        const cv::Rect imageRC(0, 0, width, height);
                        const auto triangles = obj->DelaunayTriangulation(imageRC, job->support);
                        job->inside.reserve(triangles.size());
                        for (const auto &tr : triangles)
                            if (imageRC.contains(tr.pt1) && imageRC.contains(tr.pt2) && imageRC.contains(tr.pt3)) job->inside.push_back(tr);
                        job->host_s = now_s() - t1;
                    });
                } else {
                    // (the worker only READS the object -- its host copy of the result, camera, depth range -- while this
                    // thread parks it; CudaPlanarPriorInitialization sets the mode flag when the view is taken up again)
                    job->done = std::async(std::launch::async, [job, obj]() {
                        const double t0 = now_s();
                        const int width = obj->GetReferenceImageWidth(), height = obj->GetReferenceImageHeight();
                        cv::Mat_<float> depths(height, width);
                        std::vector<float> costs((size_t)width * height);
                        for (int k = 0; k < width * height; ++k) {
                            depths.ptr()[k] = obj->GetPlaneHypothesis(k).w;
                            costs[k] = obj->GetCost(k);
                        }
                        PlanarPriorCpu(obj->GetReferenceCamera(), depths, costs.data(), obj->GetMinDepth(), obj->GetMaxDepth(), job->mask_tri,
                                       job->planeParams_tri);
                        add_prior_s(now_s() - t0);
                    });
                }
                // only the stage state stays with the view (CPU prior stage: and the host copy its worker reads)
                acmmp.Park(false, !g_gpu_prior);
                if (previous != (size_t)-1) finish_view(previous);
                previous = i;
            }
            if (previous != (size_t)-1) finish_view(previous);
            t_sweep1 += now_s() - t_s1;
            if (!timed_wait()) return;                                             // every owner's map is in its table
            if (ndev > 1) pull(dtab);
            if (!timed_wait()) return;
            const double t_g = now_s();
            for (int geom_iter = 0; geom_iter < 2; ++geom_iter) {                    // geometric sweeps
                const bool multi_geometry = geom_iter > 0;
                for (size_t i : mine) {
                    const Problem &problem = problems[i];
                    ACMMP &acmmp = *objs[i];
                    double ta = now_s();
                    activate(i, false);
                    t_views += now_s() - ta;
                    acmmp.ResetModes();
                    acmmp.SetGeomConsistencyParams(multi_geometry);
                    std::vector<const float *> maps;
                    std::vector<int> ws, hs;
                    for (int id : problem.src_image_ids) {
                        const DeviceMap &m = multi_geometry ? gtab[d][index_of.at(id)] : dtab[d][index_of.at(id)];
                        maps.push_back(m.ptr);
                        ws.push_back(m.w);
                        hs.push_back(m.h);
                    }
                    acmmp.SetNeighbourDepthMapsDevice(maps, ws, hs);
                    const bool last = finest && multi_geometry;
                    double tg = now_s();
                    acmmp.RunPatchMatchResident(last);
                    t_run += now_s() - tg;
                    float t[8];
                    acmmp.GetTimings(t);
                    gpu_ms += t[0] + t[1] + t[2];
                    tg = now_s();
                    gtab[d][i].fit(acmmp.GetReferenceImageWidth(), acmmp.GetReferenceImageHeight(), full_px[i], g_device + d);
                    acmmp.ExportDepthDevice(gtab[d][i].ptr);
                    acmmp.Park(false, last);                   // the writer below reads the host copy, then lets it go too
                    t_export += now_s() - tg;
                    tg = now_s();
                    if (last) {
                        // the view's four .dmb files on a writer thread: the result sits in the object's host buffers, which
                        // nothing touches any more (this was the view's last stage), and this thread goes on with the next view
                        ACMMP *obj = &acmmp;
                        const int ref_id = problem.ref_image_id;
                        const cv::Mat_<float> *prior_depth = &final_prior_depth[i];
                        writers.push_back(std::async(std::launch::async, [obj, ref_id, prior_depth, &dense_folder]() {
                            const int width = obj->GetReferenceImageWidth(), height = obj->GetReferenceImageHeight();
                            const size_t npx = (size_t)width * height;
                            cv::Mat_<float> depths(height, width), costs(height, width);
                            cv::Mat_<cv::Vec3f> normals(height, width);
                            {
                                // one pass over the pinned result, then the pinned buffers go back to the pool for the next
                                // view's download (a page-locked allocation costs 20 - 100 ms): the files are written from the copies
                                const float *ph = obj->PlanesHost();
                                float *dd = depths.ptr();
                                cv::Vec3f *nn = normals.ptr();
                                for (size_t k = 0; k < npx; ++k) {
                                    nn[k] = cv::Vec3f(ph[4 * k], ph[4 * k + 1], ph[4 * k + 2]);
                                    dd[k] = ph[4 * k + 3];
                                }
                                std::memcpy(costs.ptr(), obj->CostsHost(), sizeof(float) * npx);
                            }
                            obj->Park();
                            const std::string result_folder = result_folder_of(dense_folder, ref_id);
                            mkdir(result_folder.c_str(), 0777);
                            writeDepthDmb(result_folder + "/depths.dmb", *prior_depth);
                            writeDepthDmb(result_folder + "/depths_geom.dmb", depths);
                            writeNormalDmb(result_folder + "/normals.dmb", normals);
                            writeDepthDmb(result_folder + "/costs.dmb", costs);
                        }));
                        std::cout << "Processing image " << std::setw(8) << std::setfill('0') << problem.ref_image_id << " done!" << std::endl;
                    }
                    t_output += now_s() - tg;
                }
                if (!multi_geometry && ndev > 1) {
                    if (!timed_wait()) return;                                     // ExportDepthDevice waits for its copy
                    pull(gtab);
                    if (!timed_wait()) return;
                }
            }
            t_geom += now_s() - t_g;
            first_level = false;
            if (!timed_wait()) return;                                             // nobody overwrites a table somebody still reads
        }
        {
            const double tw = now_s();
            for (auto &w : writers) w.get();                 // rethrows a writer's exception
            t_output += now_s() - tw;
        }
        std::lock_guard<std::mutex> lock(stats_mutex);
        g_gpu_ms += gpu_ms;
        g_t_ctx = std::max(g_t_ctx, t_ctx); g_t_upload = std::max(g_t_upload, t_upload); g_t_support = std::max(g_t_support, t_support);
        g_t_prior_dev = std::max(g_t_prior_dev, t_prior_dev);
        g_t_barrier = std::max(g_t_barrier, t_barrier);
        g_t_load = std::max(g_t_load, t_load); g_t_views = std::max(g_t_views, t_views); g_t_run = std::max(g_t_run, t_run);
        g_t_export = std::max(g_t_export, t_export); g_t_output = std::max(g_t_output, t_output); g_t_join = std::max(g_t_join, t_join);
        g_t_sweep1 = std::max(g_t_sweep1, t_sweep1); g_t_geom = std::max(g_t_geom, t_geom); g_t_exchange = std::max(g_t_exchange, t_exchange);
    };
    auto guarded = [&](const int d) {
        try {
            device_main(d);
        } catch (const std::exception &e) {
            {
                std::lock_guard<std::mutex> lock(stats_mutex);
                if (failure.empty()) failure = e.what();
            }
            barrier.abort();
        }
    };
    std::vector<std::thread> threads;
    for (int d = 1; d < ndev; ++d) threads.emplace_back(guarded, d);
    guarded(0);
    for (auto &t : threads) t.join();
    if (resident_out && failure.empty()) {
        // What the fusion would read back from the .dmb files and the image folder is still on the devices: final depth maps
        // (the tables the last sweep exported into), the contexts' planes (world normal + depth), the level images.  The
        // fusion runs on the first device; the maps of views another device owns come over with three peer copies each.
        const double t0 = now_s();
        cudaSetDevice(g_device);
        bool ok = true;
        for (size_t i = 0; i < num_images && ok; ++i) {
            const int d = device_of(i);
            const size_t npx = (size_t)level_w[i] * level_h[i];
            ResidentView r;
            r.ref_image_id = problems[i].ref_image_id;
            r.cam = level_camera[i];
            r.width = level_w[i];
            r.height = level_h[i];
            const float *src[3] = {gtab[d][i].ptr, objs[i]->GetPlanesDevice(), pools[d][i].ptr};
            const size_t bytes[3] = {sizeof(float) * npx, 4 * sizeof(float) * npx, sizeof(float) * npx};
            const float *dst[3] = {src[0], src[1], src[2]};
            if (d != 0) {
                for (int k = 0; k < 3 && ok; ++k) {
                    void *p = nullptr;
                    ok = acmmp_pool_alloc(g_device, bytes[k], &p) == ACMMP_OK &&
                         cudaMemcpyPeer(p, g_device, src[k], g_device + d, bytes[k]) == cudaSuccess;
                    dst[k] = (const float *)p;
                }
            }
            r.depth_dev = dst[0];
            r.planes4_dev = dst[1];
            r.gray_dev = dst[2];
            resident_out->push_back(r);
        }
        if (!ok) {
            (void)cudaGetLastError();
            resident_out->clear();                      // the fusion reads the files instead
        }
        g_t_exchange += now_s() - t0;
    }
    // The process ends right after this schedule: the contexts (a few GB of pooled device and pinned buffers each)
    // are left to the driver's process teardown, which reclaims them much faster than hundreds of cudaFree /
    // cudaFreeHost calls would (measured: ~0.4 s per view at C2 size).
    for (auto &o : objs) o.release();
    if (!failure.empty()) throw std::runtime_error(failure);
    g_devices_used = ndev;
    return true;
}

} // namespace

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::cout << "USAGE: acmmp_b200 dense_folder [--seed S] [--device D] [--gpus N] [--max-views N] [--resident 0|1] [--gpu-prior 0|1] [--fusion 0|1] [--fusion-only 0|1]" << std::endl;
        return -1;
    }
    const std::string dense_folder = argv[1];
    size_t max_views = 0;
    int resident = 0, fusion_only = 0;
    for (int i = 2; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--resident")) resident = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--gpu-prior")) g_gpu_prior = std::atoi(argv[i + 1]);
        if (!std::strcmp(argv[i], "--seed")) g_seed = std::strtoull(argv[i + 1], nullptr, 10);
        else if (!std::strcmp(argv[i], "--device")) g_device = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--max-views")) max_views = (size_t)std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--gpus")) { g_gpus = std::atoi(argv[i + 1]); resident = 1; }      // multi-GPU is a resident schedule
        else if (!std::strcmp(argv[i], "--fusion")) g_fusion = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--fusion-only")) fusion_only = std::atoi(argv[i + 1]);       // re-fuse the maps of an earlier run
    }
    std::vector<Problem> problems;
    GenerateSampleList(dense_folder, problems);
    mkdir((dense_folder + "/ACMMP").c_str(), 0777);
    const size_t num_images = max_views ? std::min(max_views, problems.size()) : problems.size();
    std::cout << "There are " << num_images << " problems needed to be processed!" << std::endl;
    if (num_images < problems.size()) {
        // --max-views: a source view that is not processed has no depth map for the geometric stages (the file-chained
        // schedule would read a depths.dmb nobody wrote).  The sub-scene keeps only neighbours that are processed.
        std::map<int, size_t> pos;
        for (size_t i = 0; i < problems.size(); ++i) pos[problems[i].ref_image_id] = i;
        size_t dropped = 0;
        for (size_t i = 0; i < num_images; ++i) {
            auto &src = problems[i].src_image_ids;
            const size_t before = src.size();
            src.erase(std::remove_if(src.begin(), src.end(), [&](int id) { const auto it = pos.find(id); return it == pos.end() || it->second >= num_images; }), src.end());
            dropped += before - src.size();
            if (src.empty()) {
                std::cerr << "acmmp_b200: --max-views " << num_images << " leaves view " << problems[i].ref_image_id << " without source views" << std::endl;
                return 1;
            }
        }
        if (dropped) std::cout << "--max-views: dropped " << dropped << " source-view references to views that are not processed" << std::endl;
    }
    const double t_start = now_s();
    std::vector<ResidentView> resident_views;          // --resident 1 on one device: the final maps stay there for the fusion
    try {
        int max_num_downscale = fusion_only ? -1 : ComputeMultiScaleSettings(dense_folder, problems);
        if (resident && !fusion_only) {
            std::vector<Problem> work = problems;
            if (RunResident(dense_folder, work, num_images, max_num_downscale, g_gpus, &resident_views)) max_num_downscale = -1;     // done
            else { std::cout << "resident schedule not applicable to this scene; running the file-chained schedule" << std::endl; resident = 0; }
        }
        int flag = 0;
        const int geom_iterations = 2;
        while (max_num_downscale >= 0) {                       // main.cpp:417-476
            std::cout << "Scale: " << max_num_downscale << std::endl;
            for (auto &problem : problems) {
                if (problem.num_downscale >= 0) {
                    problem.cur_image_size = problem.max_image_size / (int)std::pow(2, problem.num_downscale);
                    problem.num_downscale--;
                }
            }
            if (flag == 0) {
                flag = 1;
                for (size_t i = 0; i < num_images; ++i) ProcessProblem(dense_folder, problems, (int)i, false, true, false);
            } else {
                for (size_t i = 0; i < num_images; ++i) JointBilateralUpsampling(dense_folder, problems[i], problems[i].cur_image_size);
                for (size_t i = 0; i < num_images; ++i) ProcessProblem(dense_folder, problems, (int)i, false, true, true);
            }
            for (int geom_iter = 0; geom_iter < geom_iterations; ++geom_iter) {
                const bool multi_geometry = geom_iter > 0;
                for (size_t i = 0; i < num_images; ++i) ProcessProblem(dense_folder, problems, (int)i, true, false, false, multi_geometry);
            }
            max_num_downscale--;
        }
    } catch (const std::exception &e) {
        std::cerr << "acmmp_b200: " << e.what() << std::endl;
        return 1;
    }
    const double wall_patchmatch = now_s() - t_start;
    double fusion_s = 0.0, fusion_kernel_ms = 0.0;
    size_t fusion_points = 0;
    if (g_fusion) {                                            // main.cpp:478-479
        try {
            const double t0 = now_s();
            std::vector<Problem> fused(problems.begin(), problems.begin() + num_images);
            fusion_points = RunFusionCuda(dense_folder, fused, true, g_device, &fusion_kernel_ms, resident_views.empty() ? nullptr : &resident_views);
            fusion_s = now_s() - t0;
        } catch (const std::exception &e) {
            std::cerr << "acmmp_b200: " << e.what() << std::endl;
            return 1;
        }
    }
    std::cout << "{\"mode\": \"" << (resident ? "resident" : "files") << "\", \"gpu_prior\": " << g_gpu_prior << ", \"views\": " << num_images << ", \"wall_s\": " << wall_patchmatch << ", \"kernel_ms\": " << g_gpu_ms
              << ", \"prior_cpu_s\": " << g_prior_s << ", \"load_s\": " << g_t_load << ", \"views_s\": " << g_t_views << ", \"run_s\": " << g_t_run
              << ", \"export_s\": " << g_t_export << ", \"output_s\": " << g_t_output << ", \"join_s\": " << g_t_join << ", \"barrier_s\": " << g_t_barrier << ", \"setup_s\": " << g_t_setup << ", \"ctx_s\": " << g_t_ctx << ", \"upload_s\": " << g_t_upload
              << ", \"support_s\": " << g_t_support << ", \"prior_dev_s\": " << g_t_prior_dev << ", \"sweep1_s\": " << g_t_sweep1 << ", \"geom_s\": " << g_t_geom
              << ", \"gpus\": " << (resident ? g_devices_used : 1) << ", \"exchange_s\": " << g_t_exchange
              << ", \"fusion_s\": " << fusion_s << ", \"fusion_kernel_ms\": " << fusion_kernel_ms << ", \"fusion_points\": " << fusion_points << "}" << std::endl;
    return 0;
}
