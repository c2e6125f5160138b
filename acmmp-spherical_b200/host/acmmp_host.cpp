// acmmp_host.cpp -- `class ACMMP` over the C ABI (see acmmp_host.h).  Which reference lines each method
// replaces is cited per method; the reference is /root/reference (read-only), nothing is copied from it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <chrono>
#include <future>
#include <map>
#include <memory>
#include <sstream>
#include <unordered_map>
#include <stdexcept>

#include "acmmp_host.h"

namespace {

std::string result_folder_of(const std::string &dense_folder, int ref_id)
{
    std::stringstream s;
    s << dense_folder << "/ACMMP/2333_" << std::setw(8) << std::setfill('0') << ref_id;
    return s.str();
}

} // namespace

ACMMP::ACMMP(int device) : device_(device)
{
    acmmp_default_params(&params_);
    const int rc = acmmp_create(&ctx_, device);
    if (rc != ACMMP_OK) throw std::runtime_error("acmmp_create failed: no usable sm_100 CUDA device (there is no CPU fallback)");
}

ACMMP::~ACMMP()
{
    if (ctx_) acmmp_destroy(ctx_);
}

void ACMMP::check(int rc, const char *what)
{
    if (rc != ACMMP_OK) throw std::runtime_error(std::string(what) + ": " + acmmp_last_error(ctx_));
}

// reference ACMMP.cpp:548-565
void ACMMP::SetGeomConsistencyParams(bool multi_geometry)
{
    params_.geom_consistency = 1;
    params_.max_iterations = 2;
    if (multi_geometry) params_.multi_geometry = 1;
    check(acmmp_set_geom_consistency(ctx_, multi_geometry ? 1 : 0), "SetGeomConsistencyParams");
}

void ACMMP::SetHierarchyParams()
{
    params_.hierarchy = 1;
    check(acmmp_set_hierarchy(ctx_), "SetHierarchyParams");
}

void ACMMP::SetPlanarPriorParams()
{
    params_.planar_prior = 1;
    check(acmmp_set_planar_prior(ctx_), "SetPlanarPriorParams");
}

void ACMMP::SetSeed(uint64_t seed) { check(acmmp_set_seed(ctx_, seed), "SetSeed"); }

// reference ACMMP.cpp:567-679: disk -> host.  Reference + source images as float grey levels, each scaled to its
// own view's cur_image_size with bilinear resampling, intrinsics scaled with it; in geometric mode also the
// depth maps of the reference and source views (depths.dmb, or depths_geom.dmb in the second geometric round).
void ACMMP::InuputInitialization(const std::string &dense_folder, const std::vector<Problem> &problems, const int idx)
{
    images_.clear();
    cameras_.clear();
    const Problem &problem = problems[idx];
    std::vector<int> ids;
    ids.push_back(problem.ref_image_id);
    ids.insert(ids.end(), problem.src_image_ids.begin(), problem.src_image_ids.end());
    // The reference indexes the problem vector BY IMAGE ID here (ACMMP.cpp:609), which reads out of bounds -- or another
    // view's size -- unless pair.txt lists the ids 0..N-1 in order.  Same result for well-formed scenes, defined for the
    // others: look the source view's problem up by id; a view that is nobody's reference view takes this view's size.
    std::unordered_map<int, size_t> index_of;
    for (size_t k = 0; k < problems.size(); ++k) index_of.emplace(problems[k].ref_image_id, k);
    for (size_t i = 0; i < ids.size(); ++i) {
        int max_image_size = problems[idx].cur_image_size;
        if (i > 0) {
            const auto it = index_of.find(ids[i]);
            if (it != index_of.end()) max_image_size = problems[it->second].cur_image_size;
        }
        cv::Mat_<float> img;
        Camera cam;
        LoadScaledView(dense_folder, ids[i], max_image_size, img, cam);
        images_.push_back(img);
        cameras_.push_back(cam);
    }
    params_.depth_min = cameras_[0].depth_min * 0.6f;
    params_.depth_max = cameras_[0].depth_max * 1.2f;
    params_.num_images = (int)images_.size();
    std::cout << "depthe range: " << params_.depth_min << " " << params_.depth_max << std::endl;
    std::cout << "num images: " << params_.num_images << std::endl;

    if (params_.geom_consistency) {
        depths_.clear();
        const std::string suffix = params_.multi_geometry ? "/depths_geom.dmb" : "/depths.dmb";
        for (int id : ids) {
            cv::Mat_<float> d;
            const std::string path = result_folder_of(dense_folder, id) + suffix;
            if (readDepthDmb(path, d) != 0 || d.rows <= 0 || d.cols <= 0)
                throw std::runtime_error("geometric stage: cannot read the depth map " + path +
                                         " (is every source view among the processed views?)");
            depths_.push_back(d);
        }
    }
}

// One view of InuputInitialization (ACMMP.cpp:576-643): grey image as float, camera file, both scaled so that the
// longer side is at most max_image_size (bilinear resampling, intrinsics scaled by the achieved ratios).
bool LoadScaledView(const std::string &dense_folder, int id, int max_image_size, cv::Mat_<float> &image, Camera &cam)
{
    bool ok = LoadGreyImage(dense_folder, id, image);
    if (!ok) std::cerr << "Error: could not read image " << id << std::endl;
    std::stringstream cam_path;
    cam_path << dense_folder << "/cams/" << std::setw(8) << std::setfill('0') << id << "_cam.txt";
    cam = ReadCamera(cam_path.str());
    cam.height = image.rows;
    cam.width = image.cols;
    if (image.cols <= max_image_size && image.rows <= max_image_size) return ok;
    const float factor_x = static_cast<float>(max_image_size) / image.cols;
    const float factor_y = static_cast<float>(max_image_size) / image.rows;
    const float factor = std::min(factor_x, factor_y);
    const int new_cols = (int)std::round(image.cols * factor);
    const int new_rows = (int)std::round(image.rows * factor);
    const float scale_x = new_cols / static_cast<float>(image.cols);
    const float scale_y = new_rows / static_cast<float>(image.rows);
    cv::Mat_<float> scaled;
    ResizeLinear(image, scaled, new_cols, new_rows);
    image = scaled;
    if (cam.model == SPHERE) {
        cam.params[1] *= scale_x;
        cam.params[2] *= scale_y;
    } else {
        cam.K[0] *= scale_x; cam.K[2] *= scale_x;
        cam.K[4] *= scale_y; cam.K[5] *= scale_y;
    }
    cam.height = new_rows;
    cam.width = new_cols;
    return ok;
}

// ---- GPU-resident stage chaining (see acmmp_host.h) ------------------------------------------------------------
void ACMMP::SetViewsHost(const std::vector<cv::Mat_<float>> &images, const std::vector<Camera> &cameras, bool next_level)
{
    images_ = images;
    cameras_ = cameras;
    const int n = (int)images_.size();
    std::vector<const float *> ptrs(n);
    std::vector<int32_t> ws(n), hs(n);
    for (int i = 0; i < n; ++i) {
        ptrs[i] = images_[i].ptr();
        ws[i] = images_[i].cols;
        hs[i] = images_[i].rows;
    }
    params_.depth_min = cameras_[0].depth_min * 0.6f;
    params_.depth_max = cameras_[0].depth_max * 1.2f;
    params_.num_images = n;
    if (next_level) {
        check(acmmp_next_level(ctx_, n, ptrs.data(), ws.data(), hs.data(), cameras_.data()), "SetViewsHost (next level)");
        params_.geom_consistency = params_.multi_geometry = params_.planar_prior = 0;
        params_.max_iterations = 3;
        params_.hierarchy = 1;
    } else {
        check(acmmp_set_views(ctx_, n, ptrs.data(), ws.data(), hs.data(), cameras_.data()), "SetViewsHost");
    }
}

void ACMMP::SetViewsDevice(const std::vector<const float *> &images_dev, const std::vector<int> &widths, const std::vector<int> &heights,
                           const std::vector<Camera> &cameras, bool next_level, const cv::Mat_<float> *ref_host)
{
    const int n = (int)images_dev.size();
    // (ref_host == nullptr: the host copy of the reference image stays what it was -- a re-activation after Park())
    cv::Mat_<float> ref_image;
    if (ref_host) ref_image = *ref_host;
    else if (!images_.empty()) ref_image = std::move(images_[0]);
    images_.assign((size_t)n, cv::Mat_<float>());
    images_[0] = std::move(ref_image);
    cameras_ = cameras;
    std::vector<int32_t> ws(widths.begin(), widths.end()), hs(heights.begin(), heights.end());
    params_.depth_min = cameras_[0].depth_min * 0.6f;
    params_.depth_max = cameras_[0].depth_max * 1.2f;
    params_.num_images = n;
    if (next_level) {
        check(acmmp_next_level_device(ctx_, n, images_dev.data(), ws.data(), hs.data(), cameras_.data()), "SetViewsDevice (next level)");
        params_.geom_consistency = params_.multi_geometry = params_.planar_prior = 0;
        params_.max_iterations = 3;
        params_.hierarchy = 1;
    } else {
        check(acmmp_set_views_device(ctx_, n, images_dev.data(), ws.data(), hs.data(), cameras_.data()), "SetViewsDevice");
    }
}

void ACMMP::ResetModes()
{
    check(acmmp_reset_modes(ctx_), "ResetModes");
    params_.geom_consistency = params_.multi_geometry = params_.planar_prior = params_.hierarchy = 0;
    params_.max_iterations = 3;
}

void ACMMP::Park(bool keep_prior, bool keep_host_result)
{
    check(acmmp_park(ctx_, keep_prior ? 1 : 0, keep_host_result ? 1 : 0), "Park");
    params_.geom_consistency = params_.multi_geometry = 0;
    if (!keep_prior) params_.planar_prior = 0;
    if (!keep_host_result) planes_host_ = costs_host_ = nullptr;
}

const float *ACMMP::GetPlanesDevice()
{
    void *planes = nullptr, *costs = nullptr;
    check(acmmp_device_buffers(ctx_, &planes, &costs), "GetPlanesDevice");
    return static_cast<const float *>(planes);
}

void ACMMP::SetNeighbourDepthMapsDevice(const std::vector<const float *> &maps_dev, const std::vector<int> &widths,
                                        const std::vector<int> &heights)
{
    const int n = (int)maps_dev.size() + 1;
    std::vector<const float *> p(n);
    std::vector<int32_t> w(n), h(n);
    p[0] = nullptr;
    w[0] = cameras_[0].width;
    h[0] = cameras_[0].height;
    for (int i = 1; i < n; ++i) {
        p[i] = maps_dev[i - 1];
        w[i] = widths[i - 1];
        h[i] = heights[i - 1];
    }
    check(acmmp_set_depth_maps_device(ctx_, n, p.data(), w.data(), h.data()), "SetNeighbourDepthMapsDevice");
}

void ACMMP::RunPatchMatchResident(bool download)
{
    check(acmmp_run_patch_match_resident(ctx_), "RunPatchMatchResident");
    if (download) {
        check(acmmp_download_result(ctx_), "RunPatchMatchResident (download)");
        check(acmmp_result_host(ctx_, &planes_host_, &costs_host_), "RunPatchMatchResident (result)");
    } else {
        check(acmmp_synchronize(ctx_), "RunPatchMatchResident (synchronize)");
        planes_host_ = costs_host_ = nullptr;
    }
}

void ACMMP::ExportDepthDevice(float *depth_dev)
{
    check(acmmp_export_depth_device(ctx_, depth_dev), "ExportDepthDevice");
    check(acmmp_synchronize(ctx_), "ExportDepthDevice (synchronize)");
}

// ---- planar-prior stage on the device (see acmmp_host.h) ------------------------------------------------------
void ACMMP::GetSupportPointsDevice(std::vector<cv::Point> &support2DPoints)
{
    const int width = GetReferenceImageWidth(), height = GetReferenceImageHeight();
    const int capacity = ((width + 4) / 5) * ((height + 4) / 5);
    std::vector<int32_t> xy(2 * (size_t)capacity);
    int n = 0;
    check(acmmp_support_points(ctx_, xy.data(), capacity, &n), "GetSupportPointsDevice");
    support2DPoints.clear();
    support2DPoints.reserve(n);
    for (int i = 0; i < n; ++i) support2DPoints.push_back(cv::Point(xy[2 * i], xy[2 * i + 1]));
}

void ACMMP::CudaPlanarPriorFromTriangles(const std::vector<Triangle> &triangles)
{
    std::vector<int32_t> tri(6 * triangles.size());
    for (size_t i = 0; i < triangles.size(); ++i) {
        tri[6 * i + 0] = triangles[i].pt1.x; tri[6 * i + 1] = triangles[i].pt1.y;
        tri[6 * i + 2] = triangles[i].pt2.x; tri[6 * i + 3] = triangles[i].pt2.y;
        tri[6 * i + 4] = triangles[i].pt3.x; tri[6 * i + 5] = triangles[i].pt3.y;
    }
    params_.planar_prior = 1;
    check(acmmp_planar_prior_from_triangles(ctx_, tri.data(), (int)triangles.size()), "CudaPlanarPriorFromTriangles");
}

// reference ACMMP.cpp:681-845: host -> device, plus the reload of the previous stage's state from .dmb files
void ACMMP::CudaSpaceInitialization(const std::string &dense_folder, const Problem &problem)
{
    const int n = (int)images_.size();
    std::vector<const float *> ptrs(n);
    std::vector<int32_t> ws(n), hs(n);
    for (int i = 0; i < n; ++i) {
        ptrs[i] = images_[i].ptr();
        ws[i] = images_[i].cols;
        hs[i] = images_[i].rows;
    }
    check(acmmp_set_views(ctx_, n, ptrs.data(), ws.data(), hs.data(), cameras_.data()), "CudaSpaceInitialization (views)");
    const std::string folder = result_folder_of(dense_folder, problem.ref_image_id);
    const int W = cameras_[0].width, H = cameras_[0].height;

    if (params_.geom_consistency) {                          // ACMMP.cpp:726-785
        std::vector<const float *> dp(n);
        std::vector<int32_t> dw(n), dh(n);
        for (int i = 0; i < n; ++i) {
            dp[i] = depths_[i].ptr();
            dw[i] = depths_[i].cols;
            dh[i] = depths_[i].rows;
        }
        check(acmmp_set_depth_maps(ctx_, n, dp.data(), dw.data(), dh.data()), "CudaSpaceInitialization (depth maps)");
        const std::string suffix = params_.multi_geometry ? "/depths_geom.dmb" : "/depths.dmb";
        cv::Mat_<float> ref_depth, ref_cost;
        cv::Mat_<cv::Vec3f> ref_normal;
        if (readDepthDmb(folder + suffix, ref_depth) != 0 || readNormalDmb(folder + "/normals.dmb", ref_normal) != 0 ||
            readDepthDmb(folder + "/costs.dmb", ref_cost) != 0)
            throw std::runtime_error("geometric stage: cannot read the previous stage's .dmb files in " + folder);
        if (ref_depth.rows != H || ref_depth.cols != W || ref_normal.rows != H || ref_cost.rows != H)
            throw std::runtime_error("geometric stage: previous-stage .dmb files do not match the image size");
        std::vector<float> planes((size_t)4 * W * H);
        for (size_t i = 0; i < (size_t)W * H; ++i) {
            const cv::Vec3f &nr = ref_normal.ptr()[i];
            planes[4 * i + 0] = nr[0]; planes[4 * i + 1] = nr[1]; planes[4 * i + 2] = nr[2];
            planes[4 * i + 3] = ref_depth.ptr()[i];
        }
        check(acmmp_set_planes(ctx_, planes.data(), ref_cost.ptr()), "CudaSpaceInitialization (planes)");
    }
    if (params_.hierarchy) {                                 // ACMMP.cpp:788-844
        cv::Mat_<float> fine_depth, coarse_cost;
        cv::Mat_<cv::Vec3f> coarse_normal;
        if (readDepthDmb(folder + "/depths.dmb", fine_depth) != 0 ||            // JBU output, already at this level's size
            readNormalDmb(folder + "/normals.dmb", coarse_normal) != 0 ||       // previous (coarser) level
            readDepthDmb(folder + "/costs.dmb", coarse_cost) != 0)
            throw std::runtime_error("hierarchy stage: cannot read the previous level's .dmb files in " + folder);
        const int sw = coarse_normal.cols, sh = coarse_normal.rows;
        if (fine_depth.rows != H || fine_depth.cols != W) throw std::runtime_error("hierarchy stage: depths.dmb is not at this level's size");
        const bool upsample = (sw != W || sh != H);
        std::vector<float> coarse((size_t)4 * sw * sh);
        for (size_t i = 0; i < (size_t)sw * sh; ++i) {
            const cv::Vec3f &nr = coarse_normal.ptr()[i];
            coarse[4 * i + 0] = nr[0]; coarse[4 * i + 1] = nr[1]; coarse[4 * i + 2] = nr[2];
            coarse[4 * i + 3] = upsample ? coarse_cost.ptr()[i] : fine_depth.ptr()[i];     // ACMMP.cpp:823-828
        }
        check(acmmp_set_hierarchy_inputs(ctx_, coarse.data(), sw, sh, fine_depth.ptr()), "CudaSpaceInitialization (hierarchy)");
    }
}

// reference ACMMP.cu:1506-1556
void ACMMP::RunPatchMatch()
{
    check(acmmp_run_patch_match(ctx_), "RunPatchMatch");
    check(acmmp_result_host(ctx_, &planes_host_, &costs_host_), "RunPatchMatch (result)");
}

void ACMMP::GetTimings(float out[8]) { check(acmmp_last_timings(ctx_, out), "GetTimings"); }

int ACMMP::GetReferenceImageWidth() { return cameras_[0].width; }
int ACMMP::GetReferenceImageHeight() { return cameras_[0].height; }
cv::Mat_<float> ACMMP::GetReferenceImage() { return images_[0]; }
float ACMMP::GetMinDepth() { return params_.depth_min; }
float ACMMP::GetMaxDepth() { return params_.depth_max; }

// reference ACMMP.cpp:884-892; index = row * width + col; (world normal xyz, depth w)
float4 ACMMP::GetPlaneHypothesis(const int index)
{
    const float *p = planes_host_ + 4 * (size_t)index;
    return make_float4(p[0], p[1], p[2], p[3]);
}

float ACMMP::GetCost(const int index) { return costs_host_[index]; }

// reference ACMMP.cpp:904-930: per 5x5 cell the pixel of least cost, kept when that cost is below 0.1
void SupportPointsOf(const float *costs, int width, int height, std::vector<cv::Point> &support2DPoints)
{
    support2DPoints.clear();
    const int step = 5;
    for (int col = 0; col < width; col += step) {
        for (int row = 0; row < height; row += step) {
            float best = 2.0f;
            cv::Point where;
            const int c_end = std::min(width, col + step), r_end = std::min(height, row + step);
            for (int c = col; c < c_end; ++c) {
                for (int r = row; r < r_end; ++r) {
                    const float cst = costs[(size_t)r * width + c];
                    if (cst < 2.0f && best > cst) {
                        where = cv::Point(c, r);
                        best = cst;
                    }
                }
            }
            if (best < 0.1f) support2DPoints.push_back(where);
        }
    }
}

void ACMMP::GetSupportPoints(std::vector<cv::Point> &support2DPoints)
{
    SupportPointsOf(costs_host_, GetReferenceImageWidth(), GetReferenceImageHeight(), support2DPoints);
}

// reference ACMMP.cpp:932-954 (cv::Subdiv2D there)
std::vector<Triangle> ACMMP::DelaunayTriangulation(const cv::Rect boundRC, const std::vector<cv::Point> &points)
{
    std::vector<Triangle> results;
    if (points.empty()) return results;
    // cv::Subdiv2D subdiv2d(boundRC) (ACMMP.cpp:938): the same enclosing triangle when the rectangle starts at the origin
    const bool at_origin = boundRC.x == 0 && boundRC.y == 0;
    const std::vector<int> idx = DelaunayIndices(points, at_origin ? boundRC.width : 0, at_origin ? boundRC.height : 0);
    results.reserve(idx.size() / 3);
    for (size_t i = 0; i + 2 < idx.size(); i += 3) results.push_back(Triangle(points[idx[i]], points[idx[i + 1]], points[idx[i + 2]]));
    return results;
}

// reference ACMMP.cpp:287-312 (the host twin of the device lifting: radial depth on the unit ray for SPHERE,
// z-depth for PINHOLE)
float3 Get3DPointonRefCam(const int x, const int y, const float depth, const Camera &camera)
{
    float3 X;
    if (camera.model == SPHERE) {
        const float lon = (static_cast<float>(x) - camera.params[1]) / static_cast<float>(camera.width) * 2.0f * (float)M_PI;
        const float lat = -(static_cast<float>(y) - camera.params[2]) / static_cast<float>(camera.height) * (float)M_PI;
        X.x = std::cos(lat) * std::sin(lon) * depth;
        X.y = -std::sin(lat) * depth;
        X.z = std::cos(lat) * std::cos(lon) * depth;
    } else {
        X.x = depth * (x - camera.K[2]) / camera.K[0];
        X.y = depth * (y - camera.K[5]) / camera.K[4];
        X.z = depth;
    }
    return X;
}

// reference ACMMP.cpp:956-989: the plane through the three lifted vertices.  The reference takes the null vector
// of the 3x4 system [X 1] with cv::SVD::solveZ; the same 1-D null space in closed form: n = (X2-X1) x (X3-X1),
// w = -n.X1, then normalised so that |n| = 1 and w >= 0.
float4 PriorPlaneParamsOf(const Camera &cam0, const Triangle &triangle, const cv::Mat_<float> &depths)
{
    const float3 a = Get3DPointonRefCam(triangle.pt1.x, triangle.pt1.y, depths(triangle.pt1.y, triangle.pt1.x), cam0);
    const float3 b = Get3DPointonRefCam(triangle.pt2.x, triangle.pt2.y, depths(triangle.pt2.y, triangle.pt2.x), cam0);
    const float3 c = Get3DPointonRefCam(triangle.pt3.x, triangle.pt3.y, depths(triangle.pt3.y, triangle.pt3.x), cam0);
    const double ux = (double)b.x - a.x, uy = (double)b.y - a.y, uz = (double)b.z - a.z;
    const double vx = (double)c.x - a.x, vy = (double)c.y - a.y, vz = (double)c.z - a.z;
    double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    double w = -(nx * a.x + ny * a.y + nz * a.z);
    double norm = std::sqrt(nx * nx + ny * ny + nz * nz);
    if (w < 0) norm = -norm;
    if (norm == 0.0) norm = 1.0;
    return make_float4((float)(nx / norm), (float)(ny / norm), (float)(nz / norm), (float)(w / norm));
}

float4 ACMMP::GetPriorPlaneParams(const Triangle triangle, const cv::Mat_<float> &depths)
{
    return PriorPlaneParamsOf(cameras_[0], triangle, depths);
}

// reference ACMMP.cpp:991-1011
float DepthFromPlaneParamOf(const Camera &cam, const float4 plane_hypothesis, const int x, const int y)
{
    if (cam.model == SPHERE) {
        const float lon = (static_cast<float>(x) - cam.params[1]) / static_cast<float>(cam.width) * 2.0f * (float)M_PI;
        const float lat = -(static_cast<float>(y) - cam.params[2]) / static_cast<float>(cam.height) * (float)M_PI;
        const float dx = std::cos(lat) * std::sin(lon), dy = -std::sin(lat), dz = std::cos(lat) * std::cos(lon);
        const float denom = plane_hypothesis.x * dx + plane_hypothesis.y * dy + plane_hypothesis.z * dz;
        return (std::abs(denom) < 1e-6f) ? 1e6f : (-plane_hypothesis.w / denom);
    }
    return -plane_hypothesis.w * cam.K[0] /
           ((x - cam.K[2]) * plane_hypothesis.x + (cam.K[0] / cam.K[4]) * (y - cam.K[5]) * plane_hypothesis.y + cam.K[0] * plane_hypothesis.z);
}

float ACMMP::GetDepthFromPlaneParam(const float4 plane_hypothesis, const int x, const int y)
{
    return DepthFromPlaneParamOf(cameras_[0], plane_hypothesis, x, y);
}

// The CPU planar-prior stage, main.cpp:113-185: support points -> Delaunay triangles -> triangle-id mask (the
// reference's barycentric stepping rasteriser) + one plane per triangle -> pixels whose prior depth leaves the depth
// range dropped.
void PlanarPriorCpu(const Camera &cam, const cv::Mat_<float> &depths, const float *costs, float depth_min, float depth_max,
                    cv::Mat_<float> &mask_tri, std::vector<float4> &planeParams_tri)
{
    const int width = depths.cols, height = depths.rows;
    const cv::Rect imageRC(0, 0, width, height);
    std::vector<cv::Point> support2DPoints;
    SupportPointsOf(costs, width, height, support2DPoints);
    mask_tri = cv::Mat_<float>::zeros(height, width);
    planeParams_tri.clear();
    if (support2DPoints.empty()) return;
    const std::vector<int> idx = DelaunayIndices(support2DPoints, width, height);
    uint32_t tri_idx = 0;
    for (size_t t = 0; t + 2 < idx.size(); t += 3) {
        const Triangle triangle(support2DPoints[idx[t]], support2DPoints[idx[t + 1]], support2DPoints[idx[t + 2]]);
        if (!(imageRC.contains(triangle.pt1) && imageRC.contains(triangle.pt2) && imageRC.contains(triangle.pt3))) continue;
        const float L01 = std::sqrt(std::pow(triangle.pt1.x - triangle.pt2.x, 2) + std::pow(triangle.pt1.y - triangle.pt2.y, 2));
        const float L02 = std::sqrt(std::pow(triangle.pt1.x - triangle.pt3.x, 2) + std::pow(triangle.pt1.y - triangle.pt3.y, 2));
        const float L12 = std::sqrt(std::pow(triangle.pt2.x - triangle.pt3.x, 2) + std::pow(triangle.pt2.y - triangle.pt3.y, 2));
        const float max_edge_length = std::max(L01, std::max(L02, L12));
        const float step = 1.0 / max_edge_length;
        // barycentric stepping rasteriser of the reference (main.cpp:153-159)
        for (float p = 0; p < 1.0; p += step) {
            for (float q = 0; q < 1.0 - p; q += step) {
                const int x = p * triangle.pt1.x + q * triangle.pt2.x + (1.0 - p - q) * triangle.pt3.x;
                const int y = p * triangle.pt1.y + q * triangle.pt2.y + (1.0 - p - q) * triangle.pt3.y;
                mask_tri(y, x) = tri_idx + 1.0;
            }
        }
        planeParams_tri.push_back(PriorPlaneParamsOf(cam, triangle, depths));
        tri_idx++;
    }
    for (int i = 0; i < width; ++i) {
        for (int j = 0; j < height; ++j) {
            if (mask_tri(j, i) > 0) {
                const float d = DepthFromPlaneParamOf(cam, planeParams_tri[(size_t)(mask_tri(j, i) - 1)], i, j);
                if (!(d <= depth_max && d >= depth_min)) mask_tri(j, i) = 0;
            }
        }
    }
}

// reference ACMMP.cpp:847-867: masks(j, i) = 1-based triangle id as float (0 = none), PlaneParams[id - 1]
void ACMMP::CudaPlanarPriorInitialization(const std::vector<float4> &PlaneParams, const cv::Mat_<float> &masks)
{
    static_assert(sizeof(float4) == 4 * sizeof(float), "float4 layout");
    check(acmmp_set_planar_prior_inputs(ctx_, reinterpret_cast<const float *>(PlaneParams.data()), (int)PlaneParams.size(), masks.ptr()),
          "CudaPlanarPriorInitialization");
}

// reference ACMMP.cpp:1071-1122
void RunJBU(const cv::Mat_<float> &scaled_image_float, const cv::Mat_<float> &src_depthmap, const std::string &dense_folder,
            const Problem &problem, int device)
{
    const int rows = scaled_image_float.rows, cols = scaled_image_float.cols;
    const int Imagescale = std::max(rows / src_depthmap.rows, cols / src_depthmap.cols);
    if (Imagescale == 1) {
        std::cout << "Image.rows = Depthmap.rows" << std::endl;
        return;
    }
    cv::Mat_<float> depthmap(rows, cols);
    const int rc = acmmp_jbu(device, scaled_image_float.ptr(), cols, rows, src_depthmap.ptr(), src_depthmap.cols, src_depthmap.rows,
                             depthmap.ptr());
    if (rc != ACMMP_OK) throw std::runtime_error("acmmp_jbu failed");
    for (size_t i = 0; i < depthmap.total(); ++i)
        if (depthmap.ptr()[i] != depthmap.ptr()[i]) { std::cout << "wrong!" << std::endl; break; }     // the reference's NaN check
    writeDepthDmb(result_folder_of(dense_folder, problem.ref_image_id) + "/depths.dmb", depthmap);
}

// reference ACMMP.cu:1817-2105 (host part) over acmmp_fusion_* (device part)
size_t RunFusionCuda(const std::string &dense_folder, const std::vector<Problem> &problems, bool geom_consistency, int device, double *kernel_ms,
                     const std::vector<ResidentView> *resident)
{
    static_assert(sizeof(PointList) == sizeof(acmmp_point), "PointList must mirror acmmp_point");
    const size_t N = problems.size();
    std::cout << "[CUDA Fusion] Starting simple fusion with " << N << " images..." << std::endl;
    struct View {
        bool ok = false;
        Camera cam;
        cv::Mat_<float> depth, gray;
        cv::Mat_<cv::Vec3f> normal;
        cv::Mat_<cv::Vec3b> colour;          // empty: no colour source, the grey level goes to all three channels
        const ResidentView *dev = nullptr;   // maps still on the device: nothing to read back
        Problem problem;
    };
    const auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    std::map<int, const ResidentView *> resident_of;
    if (resident)
        for (const ResidentView &r : *resident) resident_of[r.ref_image_id] = &r;
    // one view: maps (from the device-resident table or the .dmb files), grey and colour image, camera at the map's size
    auto load_view = [&](const size_t i, View &v) {
        const int id = problems[i].ref_image_id;
        v.problem = problems[i];
        const auto rit = resident_of.find(id);
        int cols = 0, rows = 0;
        if (rit != resident_of.end()) {
            v.dev = rit->second;
            v.cam = v.dev->cam;
            cols = v.dev->width;
            rows = v.dev->height;
        } else {
            std::stringstream cam_path;
            cam_path << dense_folder << "/cams/" << std::setw(8) << std::setfill('0') << id << "_cam.txt";
            v.cam = ReadCamera(cam_path.str());
            const std::string folder = result_folder_of(dense_folder, id);
            if (readDepthDmb(folder + (geom_consistency ? "/depths_geom.dmb" : "/depths.dmb"), v.depth) != 0) {
                std::cerr << "Warning: Could not load depth for image " << id << std::endl;
                return;
            }
            if (readNormalDmb(folder + "/normals.dmb", v.normal) != 0) {
                std::cerr << "Warning: Could not load normals for image " << id << std::endl;
                return;
            }
            cols = v.depth.cols;
            rows = v.depth.rows;
        }
        cv::Mat_<cv::Vec3b> colour;
        const bool have_colour = LoadColourImage(dense_folder, id, colour);
        int image_cols = have_colour ? colour.cols : 0, image_rows = have_colour ? colour.rows : 0;
        if (!v.dev) {
            cv::Mat_<float> image;
            if (!LoadGreyImage(dense_folder, id, image)) {
                std::cerr << "Warning: Could not load image " << id << std::endl;
                return;
            }
            image_cols = image.cols;
            image_rows = image.rows;
            if (cols == image.cols && rows == image.rows) v.gray = image;
            else ResizeLinear(image, v.gray, cols, rows);
        }
        if (have_colour && colour.cols == image_cols && colour.rows == image_rows) {
            if (cols == colour.cols && rows == colour.rows) v.colour = colour;
            else ResizeLinearBgr(colour, v.colour, cols, rows);
        }
        // RescaleImageAndCamera (ACMMP.cpp:213-245): image and intrinsics to the depth map's resolution (a resident view's
        // camera already is the finest level's)
        v.cam.width = cols;
        v.cam.height = rows;
        if (!v.dev && !(cols == image_cols && rows == image_rows)) {
            const float scale_x = cols / static_cast<float>(image_cols), scale_y = rows / static_cast<float>(image_rows);
            if (v.cam.model == SPHERE) {
                v.cam.params[1] *= scale_x;
                v.cam.params[2] *= scale_y;
            } else {
                v.cam.K[0] *= scale_x; v.cam.K[2] *= scale_x;
                v.cam.K[4] *= scale_y; v.cam.K[5] *= scale_y;
            }
        }
        v.ok = true;
    };
    // file reads, decoding and resampling of the views on a few host threads (the JPEG decoder serialises itself)
    std::vector<View> loaded(N);
    {
        const size_t workers = std::min<size_t>(8, std::max<size_t>(1, N));
        std::vector<std::future<void>> jobs;
        for (size_t w = 0; w < workers; ++w)
            jobs.push_back(std::async(std::launch::async, [&, w]() {
                for (size_t i = w; i < N; i += workers) load_view(i, loaded[i]);
            }));
        for (auto &j : jobs) j.get();
    }
    std::vector<View> views;
    std::map<int, int> id_to_index;
    for (size_t i = 0; i < N; ++i) {
        if (!loaded[i].ok) continue;
        id_to_index[problems[i].ref_image_id] = (int)views.size();
        views.push_back(std::move(loaded[i]));
    }
    loaded.clear();
    const size_t num_valid = views.size();
    std::cout << "[CUDA Fusion] Successfully loaded " << num_valid << "/" << N << " images" << std::endl;
    if (num_valid == 0) {
        std::cerr << "Error: No valid images to process!" << std::endl;
        return 0;
    }
    const double t_loaded = now();
    acmmp_fusion *f = nullptr;
    if (acmmp_fusion_create(device, (int)num_valid, &f) != ACMMP_OK) throw std::runtime_error("RunFusionCuda: no usable sm_100 device (there is no CPU fallback)");
    unsigned char *staging[2] = {nullptr, nullptr};
    auto fail = [&](const char *what) {
        const std::string msg = std::string("RunFusionCuda (") + what + "): " + acmmp_fusion_last_error(f);
        acmmp_fusion_destroy(f);
        for (unsigned char *p : staging)
            if (p) cudaFreeHost(p);
        throw std::runtime_error(msg);
    };
    size_t max_px = 0;
    for (size_t i = 0; i < num_valid; ++i) {
        const View &v = views[i];
        const int cols = v.cam.width, rows = v.cam.height;
        max_px = std::max(max_px, (size_t)cols * rows);
        const int rc = v.dev ? acmmp_fusion_set_view_device(f, (int)i, &v.cam, cols, rows, v.dev->depth_dev, v.dev->planes4_dev, v.dev->gray_dev)
                             : acmmp_fusion_set_view(f, (int)i, &v.cam, cols, rows, v.depth.ptr(), reinterpret_cast<const float *>(v.normal.ptr()), v.gray.ptr());
        if (rc != ACMMP_OK) fail("set_view");
        if (!v.colour.empty() && acmmp_fusion_set_view_colour(f, (int)i, reinterpret_cast<const uint8_t *>(v.colour.ptr()), cols, rows) != ACMMP_OK)
            fail("set_view_colour");
    }
    const double t_uploaded = now();
    // Pass 1: how many points every view yields (the kernel alone: 2 ms per 3200x2130 view) -- the PLY header carries the
    // total.  Pass 2: the points leave the device as the file's 27-byte vertex records into a pinned buffer and go from there
    // straight into the file, a writer thread taking view i while the device fuses view i + 1.  The reference copies 36 bytes
    // + a flag per PIXEL to the host, filters there, collects a vector of PointList and issues nine fwrite calls per point
    // (ACMMP.cu:2056-2076, ACMMP.cpp:481-534).
    std::vector<std::vector<int32_t>> srcs(num_valid);
    std::vector<int> counts(num_valid, 0);
    size_t total_points = 0;
    double ms_sum = 0.0;
    for (size_t i = 0; i < num_valid; ++i) {
        const View &v = views[i];
        for (size_t j = 0; j < std::min<size_t>(v.problem.src_image_ids.size(), 32); ++j) {
            const auto it = id_to_index.find(v.problem.src_image_ids[j]);
            srcs[i].push_back(it != id_to_index.end() ? it->second : -1);
        }
        float ms = 0.f;
        const int rc = acmmp_fusion_run_ply(f, (int)i, (int)srcs[i].size(), srcs[i].data(), nullptr, 0, &counts[i], &ms);
        if (rc != ACMMP_OK && counts[i] <= 0) fail("count");          // "capacity too small" is the answer asked for
        total_points += (size_t)counts[i];
    }
    const std::string output_path = dense_folder + "/ACMMP/ACMM_model_cuda_5.ply";
    FILE *ply = OpenPlyForVertexRecords(output_path, total_points);
    for (auto &p : staging)
        if (cudaMallocHost(reinterpret_cast<void **>(&p), 27 * max_px) != cudaSuccess) { p = nullptr; fclose(ply); fail("pinned staging buffer"); }
    std::future<bool> writer;
    bool write_ok = true;
    for (size_t i = 0; i < num_valid; ++i) {
        const View &v = views[i];
        std::cout << "[CUDA Fusion] Processing image " << (i + 1) << "/" << num_valid << " (ID=" << v.problem.ref_image_id << ", " << v.cam.width
                  << "x" << v.cam.height << ")" << std::endl;
        unsigned char *buf = staging[i & 1];
        int count = 0;
        float ms = 0.f;
        if (acmmp_fusion_run_ply(f, (int)i, (int)srcs[i].size(), srcs[i].data(), buf, counts[i], &count, &ms) != ACMMP_OK || count != counts[i]) {
            if (writer.valid()) writer.get();
            fclose(ply);
            fail("run");
        }
        ms_sum += ms;
        std::cout << "  -> Generated " << count << " points" << std::endl;
        if (writer.valid()) write_ok = writer.get() && write_ok;          // the other buffer is free again after this
        writer = std::async(std::launch::async, [ply, buf, count]() { return count == 0 || fwrite(buf, 27, (size_t)count, ply) == (size_t)count; });
    }
    if (writer.valid()) write_ok = writer.get() && write_ok;
    fclose(ply);
    acmmp_fusion_destroy(f);
    for (unsigned char *p : staging) cudaFreeHost(p);
    if (!write_ok) throw std::runtime_error("short write to " + output_path);
    std::cout << "[CUDA Fusion] read maps + images " << t_loaded - t_begin << " s, upload " << t_uploaded - t_loaded << " s, fuse + write "
              << now() - t_uploaded << " s" << std::endl;
    std::cout << "[CUDA Fusion] Complete! Wrote " << total_points << " points to " << output_path << std::endl;
    if (kernel_ms) *kernel_ms = ms_sum;
    return total_points;
}
