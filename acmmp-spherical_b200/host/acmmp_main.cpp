// acmmp_main.cpp -- `acmmp_b200 dense_folder [--seed S] [--device D] [--max-views N]`: the reference's pipeline
// schedule (main.cpp:392-482) on top of the B200 library, stage by stage through the same .dmb files:
//   per pyramid level (coarsest first):
//     [level > 0] JBU of depths_geom.dmb -> depths.dmb, then photometric stage with hierarchy
//     photometric stage -> CPU planar prior -> prior stage (same object)            -> depths.dmb
//     2 x geometric-consistency stage (the second one with multi_geometry)          -> depths_geom.dmb
// Fusion (RunFusionCuda, main.cpp:478-479) is not part of this path.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <sys/stat.h>
#include <sys/types.h>

#include "acmmp_host.h"

namespace {

uint64_t g_seed = 0;
int g_device = 0;
double g_gpu_ms = 0.0, g_prior_s = 0.0;

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
        const double t0 = now_s();
        acmmp.SetPlanarPriorParams();
        const cv::Rect imageRC(0, 0, width, height);
        std::vector<cv::Point> support2DPoints;
        acmmp.GetSupportPoints(support2DPoints);
        const auto triangles = acmmp.DelaunayTriangulation(imageRC, support2DPoints);
        cv::Mat_<float> mask_tri = cv::Mat_<float>::zeros(height, width);
        std::vector<float4> planeParams_tri;
        uint32_t tri_idx = 0;
        for (const auto &triangle : triangles) {
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
            planeParams_tri.push_back(acmmp.GetPriorPlaneParams(triangle, depths));
            tri_idx++;
        }
        for (int i = 0; i < width; ++i) {
            for (int j = 0; j < height; ++j) {
                if (mask_tri(j, i) > 0) {
                    const float d = acmmp.GetDepthFromPlaneParam(planeParams_tri[(size_t)(mask_tri(j, i) - 1)], i, j);
                    if (!(d <= acmmp.GetMaxDepth() && d >= acmmp.GetMinDepth())) mask_tri(j, i) = 0;
                }
            }
        }
        g_prior_s += now_s() - t0;
        acmmp.CudaPlanarPriorInitialization(planeParams_tri, mask_tri);
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

} // namespace

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::cout << "USAGE: acmmp_b200 dense_folder [--seed S] [--device D] [--max-views N]" << std::endl;
        return -1;
    }
    const std::string dense_folder = argv[1];
    size_t max_views = 0;
    for (int i = 2; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--seed")) g_seed = std::strtoull(argv[i + 1], nullptr, 10);
        else if (!std::strcmp(argv[i], "--device")) g_device = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--max-views")) max_views = (size_t)std::atoi(argv[i + 1]);
    }
    std::vector<Problem> problems;
    GenerateSampleList(dense_folder, problems);
    mkdir((dense_folder + "/ACMMP").c_str(), 0777);
    const size_t num_images = max_views ? std::min(max_views, problems.size()) : problems.size();
    std::cout << "There are " << num_images << " problems needed to be processed!" << std::endl;
    const double t_start = now_s();
    try {
        int max_num_downscale = ComputeMultiScaleSettings(dense_folder, problems);
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
    std::cout << "{\"views\": " << num_images << ", \"wall_s\": " << now_s() - t_start << ", \"kernel_ms\": " << g_gpu_ms
              << ", \"prior_cpu_s\": " << g_prior_s << "}" << std::endl;
    return 0;
}
