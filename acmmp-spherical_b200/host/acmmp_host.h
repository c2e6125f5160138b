// acmmp_host.h -- C++ host side over the C ABI of libacmmp_b200.so.
//
// `class ACMMP` keeps the reference's host-class surface (reference ACMMP.h:57-81: same method names --
// including the `Inuput` typo --, same argument meaning, same getters) so that the reference's pipeline
// driver (main.cpp:73-238) compiles against it unchanged in spirit; acmmp_main.cpp is that driver.
// Underneath, every device-side step is one call into include/acmmp_b200.h; there is no CPU fallback.
//
// Differences a caller can observe, all deliberate:
//   * CUDA errors do not exit(): methods throw std::runtime_error with the library's message
//     (the reference prints and exit(EXIT_FAILURE)s, ACMMP.cpp:64-97).
//   * the cuRAND seed is explicit (SetSeed, default 0) instead of clock64() (ACMMP.cu:684).
//   * images are read from images/%08d.jpg through nvJPEG (luma plane = what cv::imread(IMREAD_GRAYSCALE)
//     returns up to the decoder's IDCT rounding), or from a lossless images/%08d.pgm twin when present.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include <vector_types.h>      // float2, float3, float4 (CUDA toolkit headers, host-only use)
#include <vector_functions.h>  // make_float4

#ifdef ACMMP_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#else
#include "cvlite.h"
#endif

#include "../../include/acmmp_b200.h"

// reference main.h:35-38, :40-54 -- the C ABI's acmmp_camera is byte-for-byte this struct
enum CameraModel { PINHOLE = ACMMP_MODEL_PINHOLE, SPHERE = ACMMP_MODEL_SPHERE };
typedef acmmp_camera Camera;
typedef acmmp_params PatchMatchParams;

// reference main.h:56-62
struct Problem {
    int ref_image_id = 0;
    std::vector<int> src_image_ids;
    int max_image_size = 3200;
    int num_downscale = 0;
    int cur_image_size = 3200;
};

// reference main.h:64-67
struct Triangle {
    cv::Point pt1, pt2, pt3;
    Triangle(const cv::Point a, const cv::Point b, const cv::Point c) : pt1(a), pt2(b), pt3(c) {}
};

// reference main.h:71-75 (layout == acmmp_point of the C ABI)
struct PointList {
    float3 coord;
    float3 normal;
    float3 color;
};

// ---- on-disk contract (reference ACMMP.cpp:146-209, :352-479, main.cpp:4-33) ---------------------------
int readDepthDmb(const std::string file_path, cv::Mat_<float> &depth);
int readNormalDmb(const std::string file_path, cv::Mat_<cv::Vec3f> &normal);
int writeDepthDmb(const std::string file_path, const cv::Mat_<float> &depth);
int writeNormalDmb(const std::string file_path, const cv::Mat_<cv::Vec3f> &normal);
Camera ReadCamera(const std::string &cam_path);
void GenerateSampleList(const std::string &dense_folder, std::vector<Problem> &problems);

// ---- images ------------------------------------------------------------------------------------------------
// Grey image of view `id` as float 0..255 (InuputInitialization, ACMMP.cpp:576-581).  Returns false when neither
// images/%08d.pgm nor images/%08d.jpg can be read.
bool LoadGreyImage(const std::string &dense_folder, int id, cv::Mat_<float> &image);
// Only the size (ComputeMultiScaleSettings reads whole images just for that, main.cpp:46-50).
bool ImageSize(const std::string &dense_folder, int id, int &cols, int &rows);
// Image `id` and its camera scaled to max_image_size the way InuputInitialization does (ACMMP.cpp:576-643)
bool LoadScaledView(const std::string &dense_folder, int id, int max_image_size, cv::Mat_<float> &image, Camera &camera);
// cv::resize(src, dst, Size(new_cols, new_rows), 0, 0, INTER_LINEAR) on a float image
void ResizeLinear(const cv::Mat_<float> &src, cv::Mat_<float> &dst, int new_cols, int new_rows);
// the colour image of a view (B, G, R like cv::imread(IMREAD_COLOR)) and cv::resize INTER_LINEAR on it: the fusion's inputs
bool LoadColourImage(const std::string &dense_folder, int id, cv::Mat_<cv::Vec3b> &image);
void ResizeLinearBgr(const cv::Mat_<cv::Vec3b> &src, cv::Mat_<cv::Vec3b> &dst, int new_cols, int new_rows);

// ---- host twins of the camera math the CPU prior stage uses (ACMMP.cpp:287-312) -------------------------
float3 Get3DPointonRefCam(const int x, const int y, const float depth, const Camera &camera);

// Delaunay triangulation of integer points inside `bound` (stands in for cv::Subdiv2D, ACMMP.cpp:932-954):
// exact integer predicates, incremental insertion with walking point location.  Returns vertex index triples.
std::vector<int> DelaunayIndices(const std::vector<cv::Point> &points, int rect_w = 0, int rect_h = 0);

// ---- the CPU planar-prior stage as free functions (what the ACMMP methods below and the driver use) -----------
// GetSupportPoints, ACMMP.cpp:904-930, on a W x H cost map
void SupportPointsOf(const float *costs, int width, int height, std::vector<cv::Point> &support2DPoints);
// GetPriorPlaneParams, ACMMP.cpp:956-989 / GetDepthFromPlaneParam, ACMMP.cpp:991-1011 for camera `cam`
float4 PriorPlaneParamsOf(const Camera &cam, const Triangle &triangle, const cv::Mat_<float> &depths);
float DepthFromPlaneParamOf(const Camera &cam, const float4 plane_hypothesis, const int x, const int y);
// The whole stage of main.cpp:113-185: support points -> Delaunay -> per-triangle plane + stepping rasteriser ->
// depth-range test.  depths / costs: the photometric stage's maps; depth_min / depth_max: GetMinDepth / GetMaxDepth.
void PlanarPriorCpu(const Camera &cam, const cv::Mat_<float> &depths, const float *costs, float depth_min, float depth_max,
                    cv::Mat_<float> &mask_tri, std::vector<float4> &planeParams_tri);

class ACMMP {
public:
    explicit ACMMP(int device = 0);      // the reference hard-codes device 0 (main.cpp:77)
    ~ACMMP();
    ACMMP(const ACMMP &) = delete;
    ACMMP &operator=(const ACMMP &) = delete;

    void InuputInitialization(const std::string &dense_folder, const std::vector<Problem> &problems, const int idx);
    void CudaSpaceInitialization(const std::string &dense_folder, const Problem &problem);
    void RunPatchMatch();
    void SetGeomConsistencyParams(bool multi_geometry = false);
    void SetPlanarPriorParams();
    void SetHierarchyParams();
    void SetSeed(uint64_t seed);

    int GetReferenceImageWidth();
    int GetReferenceImageHeight();
    cv::Mat_<float> GetReferenceImage();
    const Camera &GetReferenceCamera() const { return cameras_[0]; }      // not in the reference
    float4 GetPlaneHypothesis(const int index);
    float GetCost(const int index);
    // the downloaded result as the library holds it (pinned): float4 (world normal, depth) per pixel, and the costs
    const float *PlanesHost() const { return planes_host_; }
    const float *CostsHost() const { return costs_host_; }
    void GetSupportPoints(std::vector<cv::Point> &support2DPoints);
    std::vector<Triangle> DelaunayTriangulation(const cv::Rect boundRC, const std::vector<cv::Point> &points);
    float4 GetPriorPlaneParams(const Triangle triangle, const cv::Mat_<float> &depths);
    float GetDepthFromPlaneParam(const float4 plane_hypothesis, const int x, const int y);
    float GetMinDepth();
    float GetMaxDepth();
    void CudaPlanarPriorInitialization(const std::vector<float4> &PlaneParams, const cv::Mat_<float> &masks);

    // not in the reference: CUDA-event times of the last RunPatchMatch {init, passes, finalize, #passes, last pass}
    void GetTimings(float out[8]);

    // ---- GPU-resident stage chaining (not in the reference) --------------------------
    // What the reference hands from stage to stage through .dmb files and a new object per stage stays on the device
    // of ONE object per view: the driver keeps the object, feeds it the next level's views / the neighbours' depth
    // maps (device pointers) and downloads a result only where the host needs it.  Same kernels, same inputs, same
    // order => the same maps as the file-chained schedule.
    //   SetViewsHost(next_level = false): what InuputInitialization + CudaSpaceInitialization do, from host arrays
    //   SetViewsHost(next_level = true) : move to the next finer level: JBU of the current result on the device,
    //                                     hierarchy inputs, SetHierarchyParams (acmmp_next_level)
    void SetViewsHost(const std::vector<cv::Mat_<float>> &images, const std::vector<Camera> &cameras, bool next_level);
    // the same with the level images already on this object's device (dense float32 W x H each): a driver uploads every
    // view's image once per level and device instead of once per object that uses it.  ref_host (optional) is what
    // GetReferenceImage returns.
    void SetViewsDevice(const std::vector<const float *> &images_dev, const std::vector<int> &widths, const std::vector<int> &heights,
                        const std::vector<Camera> &cameras, bool next_level, const cv::Mat_<float> *ref_host = nullptr);
    void ResetModes();                                   // flags of a freshly constructed object, views kept
    // Between two stages of a resident view: everything but the stage state (planes, costs, coarse planes) goes back to
    // the device's pool for the next view's stage (acmmp_park); SetViewsDevice(..., next_level = false) with the same
    // shapes takes it up again.  keep_host_result: GetPlaneHypothesis / GetCost stay readable (a writer thread's input).
    void Park(bool keep_prior = false, bool keep_host_result = false);
    // the stage state on the device: float4 per pixel (world-frame normal, depth); stays valid while the object lives
    const float *GetPlanesDevice();
    // depth maps of the SOURCE views for the geometric term (device pointers); the reference view's own map is the
    // state on the device (acmmp_set_depth_maps_device with maps[0] == NULL)
    void SetNeighbourDepthMapsDevice(const std::vector<const float *> &maps_dev, const std::vector<int> &widths,
                                     const std::vector<int> &heights);
    void RunPatchMatchResident(bool download);           // download: fill what GetPlaneHypothesis / GetCost read
    void ExportDepthDevice(float *depth_dev);            // W*H float32, what depths*.dmb would hold

    // ---- planar-prior stage on the device: only the triangulation stays on the host ----
    // GetSupportPoints on the costs of the state on the device (no result download needed)
    void GetSupportPointsDevice(std::vector<cv::Point> &support2DPoints);
    // plane per triangle + rasteriser + depth-range test + CudaPlanarPriorInitialization, all on the device; the
    // triangles must lie inside the image and get the ids 1..n in the given order (main.cpp:140-166)
    void CudaPlanarPriorFromTriangles(const std::vector<Triangle> &triangles);

private:
    void check(int rc, const char *what);
    acmmp_ctx *ctx_ = nullptr;
    int device_ = 0;
    std::vector<cv::Mat_<float>> images_;
    std::vector<cv::Mat_<float>> depths_;
    std::vector<Camera> cameras_;
    PatchMatchParams params_;
    const float *planes_host_ = nullptr;      // pinned result buffers owned by the library
    const float *costs_host_ = nullptr;
};

// StoreColorPlyFileBinaryPointCloud (ACMMP.cpp:481-534): binary little-endian PLY, x y z nx ny nz (float) + red green blue
// (uchar); non-finite coordinates are written as 0
void StoreColorPlyFileBinaryPointCloud(const std::string &plyFilePath, const std::vector<PointList> &pc);
void StorePlyVertexRecords(const std::string &plyFilePath, const unsigned char *records27, size_t n_points);       // the same file from packed records
FILE *OpenPlyForVertexRecords(const std::string &plyFilePath, size_t n_points);     // header written, positioned at the first vertex record

// RunFusionCuda (ACMMP.cu:1817-2105): fuse the depth / normal maps of every view (depths_geom.dmb when geom_consistency,
// else depths.dmb; normals.dmb) into ACMMP/ACMM_model_cuda_5.ply.  The per-pixel consistency kernel and the compaction
// of its points run on the device (acmmp_fusion_*).  Colours: the view's colour image like the reference's
// cv::imread(IMREAD_COLOR) (LoadColourImage: images/%08d.ppm or the .jpg through nvJPEG, rescaled to the map's size; a view
// that only has a grey image gets (g, g, g)).  Returns the number of points written.  kernel_ms: optional, sum of the
// CUDA-event kernel times.
// `resident`: views whose final maps are still on `device` (the resident schedule leaves them there): their depth / normal
// maps and grey image are taken from the device instead of the .dmb files and the image folder.
struct ResidentView {
    int ref_image_id = 0;
    Camera cam;                      // at the map's size (the finest level's camera)
    int width = 0, height = 0;
    const float *depth_dev = nullptr;      // width * height
    const float *planes4_dev = nullptr;    // float4 per pixel: world-frame normal xyz (w unused)
    const float *gray_dev = nullptr;       // grey levels 0..255
};
size_t RunFusionCuda(const std::string &dense_folder, const std::vector<Problem> &problems, bool geom_consistency, int device = 0,
                     double *kernel_ms = nullptr, const std::vector<ResidentView> *resident = nullptr);

// RunJBU (ACMMP.cpp:1071-1122): joint-bilateral upsampling of depths_geom.dmb to the new level, written as depths.dmb
void RunJBU(const cv::Mat_<float> &scaled_image_float, const cv::Mat_<float> &src_depthmap, const std::string &dense_folder,
            const Problem &problem, int device = 0);
