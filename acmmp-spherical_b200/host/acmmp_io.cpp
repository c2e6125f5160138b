// acmmp_io.cpp -- the on-disk contract of the reference, kept as is: cams/%08d_cam.txt, images/%08d.jpg,
// pair.txt in; ACMMP/2333_%08d/{depths,depths_geom,normals,costs}.dmb out.  Formats: SURVEY.md section 8(b)/(f) N4.
#include <cmath>
#include <cstdio>
#include <cfloat>
#include <cstring>
#include <stdexcept>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "acmmp_host.h"

#ifdef ACMMP_WITH_NVJPEG
#include <cuda_runtime.h>
#include <mutex>
#include <nvjpeg.h>
#endif

// ---- .dmb: int32 type(=1), h, w, nb then h*w*nb float32, row major (reference ACMMP.cpp:352-479) -----------
namespace {

int read_dmb(const std::string &path, int want_nb, int &h, int &w, std::vector<float> &data)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) {
        std::cout << "Error opening file " << path << std::endl;
        return -1;
    }
    int32_t hdr[4] = {-1, 0, 0, 0};
    const size_t got = std::fread(hdr, sizeof(int32_t), 4, f);
    if (got != 4 || hdr[0] != 1 || hdr[1] <= 0 || hdr[2] <= 0 || hdr[3] != want_nb) {
        std::fclose(f);
        return -1;
    }
    h = hdr[1]; w = hdr[2];
    data.resize((size_t)h * w * want_nb);
    const size_t n = std::fread(data.data(), sizeof(float), data.size(), f);
    std::fclose(f);
    return n == data.size() ? 0 : -1;
}

int write_dmb(const std::string &path, int h, int w, int nb, const float *data)
{
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) {
        std::cout << "Error opening file " << path << std::endl;
        return -1;
    }
    const int32_t hdr[4] = {1, h, w, nb};
    std::fwrite(hdr, sizeof(int32_t), 4, f);
    std::fwrite(data, sizeof(float), (size_t)h * w * nb, f);
    std::fclose(f);
    return 0;
}

std::string view_file(const std::string &folder, int id, const char *suffix)
{
    std::stringstream s;
    s << folder << "/" << std::setw(8) << std::setfill('0') << id << suffix;
    return s.str();
}

} // namespace

int readDepthDmb(const std::string file_path, cv::Mat_<float> &depth)
{
    int h, w;
    std::vector<float> d;
    if (read_dmb(file_path, 1, h, w, d)) return -1;
    depth = cv::Mat_<float>(h, w);
    std::memcpy(depth.ptr(), d.data(), d.size() * sizeof(float));
    return 0;
}

int readNormalDmb(const std::string file_path, cv::Mat_<cv::Vec3f> &normal)
{
    int h, w;
    std::vector<float> d;
    if (read_dmb(file_path, 3, h, w, d)) return -1;
    normal = cv::Mat_<cv::Vec3f>(h, w);
    static_assert(sizeof(cv::Vec3f) == 3 * sizeof(float), "Vec3f must be three packed floats");
    std::memcpy(static_cast<void *>(normal.ptr()), d.data(), d.size() * sizeof(float));
    return 0;
}

int writeDepthDmb(const std::string file_path, const cv::Mat_<float> &depth)
{
    return write_dmb(file_path, depth.rows, depth.cols, 1, depth.ptr());
}

int writeNormalDmb(const std::string file_path, const cv::Mat_<cv::Vec3f> &normal)
{
    return write_dmb(file_path, normal.rows, normal.cols, 3, reinterpret_cast<const float *>(normal.ptr()));
}

// ---- cams/%08d_cam.txt (reference ACMMP.cpp:146-209) -------------------------------------------------------
// "extrinsic" + 4x4 row-major [R|t; 0 0 0 1], "intrinsic" + either `SPHERE\n f cx cy` or a 3x3 K, then the depth
// line: SPHERE reads `dmin dinterval ndepth dmax`, PINHOLE reads `dmin dmax _ _` (the reference's reader, kept).
Camera ReadCamera(const std::string &cam_path)
{
    Camera camera;
    std::memset(&camera, 0, sizeof(camera));
    std::ifstream file(cam_path);
    if (!file.is_open()) {
        std::cerr << "Error: Could not open camera file: " << cam_path << std::endl;
        return camera;
    }
    std::string token;
    file >> token;                                  // "extrinsic"
    for (int i = 0; i < 3; ++i) file >> camera.R[3 * i + 0] >> camera.R[3 * i + 1] >> camera.R[3 * i + 2] >> camera.t[i];
    float skip;
    for (int i = 0; i < 4; ++i) file >> skip;       // the 0 0 0 1 row
    file >> token;                                  // "intrinsic"
    file >> token;                                  // model name or K[0]
    if (token == "SPHERE") {
        camera.model = SPHERE;
        file >> camera.params[0] >> camera.params[1] >> camera.params[2];
        float dmin, dint, dmax;
        int nplanes;
        file >> dmin >> dint >> nplanes >> dmax;
        camera.depth_min = dmin;
        camera.depth_max = dmax;
    } else {
        camera.model = PINHOLE;
        camera.K[0] = std::stof(token);
        file >> camera.K[1] >> camera.K[2] >> camera.K[3] >> camera.K[4] >> camera.K[5] >> camera.K[6] >> camera.K[7] >> camera.K[8];
        float d1, d2;
        file >> camera.depth_min >> camera.depth_max >> d1 >> d2;
    }
    return camera;
}

// ---- pair.txt (reference main.cpp:4-33): N, then per view `ref_id` and `n id score ...`; score <= 0 dropped ----
void GenerateSampleList(const std::string &dense_folder, std::vector<Problem> &problems)
{
    problems.clear();
    std::ifstream file(dense_folder + "/pair.txt");
    int num_images = 0;
    file >> num_images;
    for (int i = 0; i < num_images; ++i) {
        Problem problem;
        file >> problem.ref_image_id;
        int num_src = 0;
        file >> num_src;
        for (int j = 0; j < num_src; ++j) {
            int id;
            float score;
            file >> id >> score;
            if (score <= 0.0f) continue;
            problem.src_image_ids.push_back(id);
        }
        problems.push_back(problem);
    }
}

// ---- images ------------------------------------------------------------------------------------------------
namespace {

bool read_file(const std::string &path, std::vector<unsigned char> &bytes)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    bytes.resize(n > 0 ? (size_t)n : 0);
    const size_t got = bytes.empty() ? 0 : std::fread(bytes.data(), 1, bytes.size(), f);
    std::fclose(f);
    return got == bytes.size() && !bytes.empty();
}

// binary PGM (P5), maxval <= 255
bool parse_pgm(const std::vector<unsigned char> &b, int &w, int &h, size_t &offset, bool header_only = false)
{
    size_t p = 0;
    auto token = [&](std::string &out) {
        out.clear();
        while (p < b.size()) {
            if (b[p] == '#') { while (p < b.size() && b[p] != '\n') ++p; }
            else if (std::isspace(b[p])) ++p;
            else break;
        }
        while (p < b.size() && !std::isspace(b[p])) out.push_back((char)b[p++]);
        return !out.empty();
    };
    std::string t;
    if (!token(t) || t != "P5") return false;
    if (!token(t)) return false;
    w = std::atoi(t.c_str());
    if (!token(t)) return false;
    h = std::atoi(t.c_str());
    if (!token(t) || std::atoi(t.c_str()) > 255) return false;
    offset = p + 1;                                 // exactly one whitespace byte after maxval
    return w > 0 && h > 0 && (header_only || offset + (size_t)w * h <= b.size());
}

// JPEG header scan for the frame size (SOF0..SOF15 except DHT/JPG/DAC)
bool jpeg_size(const std::vector<unsigned char> &b, int &w, int &h)
{
    size_t p = 2;
    if (b.size() < 4 || b[0] != 0xFF || b[1] != 0xD8) return false;
    while (p + 9 < b.size()) {
        if (b[p] != 0xFF) { ++p; continue; }
        const unsigned char m = b[p + 1];
        if (m == 0xFF) { ++p; continue; }
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) { p += 2; continue; }
        const size_t len = ((size_t)b[p + 2] << 8) | b[p + 3];
        if (m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            h = (b[p + 5] << 8) | b[p + 6];
            w = (b[p + 7] << 8) | b[p + 8];
            return w > 0 && h > 0;
        }
        p += 2 + len;
    }
    return false;
}

#ifdef ACMMP_WITH_NVJPEG
bool decode_jpeg_luma(const std::vector<unsigned char> &b, cv::Mat_<float> &image)
{
    // one decoder state for the process: the device threads of a multi-GPU run load their views at the same time
    static std::mutex decoder_mutex;
    std::lock_guard<std::mutex> lock(decoder_mutex);
    static nvjpegHandle_t handle = nullptr;
    static nvjpegJpegState_t state = nullptr;
    if (!handle) {
        if (nvjpegCreateSimple(&handle) != NVJPEG_STATUS_SUCCESS) { handle = nullptr; return false; }
        if (nvjpegJpegStateCreate(handle, &state) != NVJPEG_STATUS_SUCCESS) return false;
    }
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (nvjpegGetImageInfo(handle, b.data(), b.size(), &ncomp, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return false;
    const int w = ws[0], h = hs[0];
    unsigned char *dev = nullptr;
    if (cudaMalloc(&dev, (size_t)w * h) != cudaSuccess) return false;
    nvjpegImage_t out;
    std::memset(&out, 0, sizeof(out));
    out.channel[0] = dev;
    out.pitch[0] = (size_t)w;
    bool ok = nvjpegDecode(handle, state, b.data(), b.size(), NVJPEG_OUTPUT_Y, &out, nullptr) == NVJPEG_STATUS_SUCCESS;
    std::vector<unsigned char> host((size_t)w * h);
    ok = ok && cudaMemcpy(host.data(), dev, host.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(dev);
    if (!ok) return false;
    image = cv::Mat_<float>(h, w);
    float *dst = image.ptr();
    for (size_t i = 0; i < host.size(); ++i) dst[i] = (float)host[i];
    return true;
}

// cv::imread(IMREAD_COLOR): interleaved B, G, R
bool decode_jpeg_bgr(const std::vector<unsigned char> &b, cv::Mat_<cv::Vec3b> &image)
{
    static std::mutex decoder_mutex;
    std::lock_guard<std::mutex> lock(decoder_mutex);
    static nvjpegHandle_t handle = nullptr;
    static nvjpegJpegState_t state = nullptr;
    if (!handle) {
        if (nvjpegCreateSimple(&handle) != NVJPEG_STATUS_SUCCESS) { handle = nullptr; return false; }
        if (nvjpegJpegStateCreate(handle, &state) != NVJPEG_STATUS_SUCCESS) return false;
    }
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (nvjpegGetImageInfo(handle, b.data(), b.size(), &ncomp, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return false;
    const int w = ws[0], h = hs[0];
    unsigned char *dev = nullptr;
    if (cudaMalloc(&dev, (size_t)w * h * 3) != cudaSuccess) return false;
    nvjpegImage_t out;
    std::memset(&out, 0, sizeof(out));
    out.channel[0] = dev;
    out.pitch[0] = (size_t)w * 3;
    bool ok = nvjpegDecode(handle, state, b.data(), b.size(), NVJPEG_OUTPUT_BGRI, &out, nullptr) == NVJPEG_STATUS_SUCCESS;
    image = cv::Mat_<cv::Vec3b>(h, w);
    ok = ok && cudaMemcpy(image.ptr(), dev, (size_t)w * h * 3, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(dev);
    return ok;
}
#endif

} // namespace

bool LoadGreyImage(const std::string &dense_folder, int id, cv::Mat_<float> &image)
{
    std::vector<unsigned char> bytes;
    if (read_file(view_file(dense_folder + "/images", id, ".pgm"), bytes)) {
        int w, h;
        size_t off;
        if (parse_pgm(bytes, w, h, off)) {
            image = cv::Mat_<float>(h, w);
            float *dst = image.ptr();
            for (size_t i = 0; i < (size_t)w * h; ++i) dst[i] = (float)bytes[off + i];
            return true;
        }
    }
    if (read_file(view_file(dense_folder + "/images", id, ".jpg"), bytes)) {
#ifdef ACMMP_WITH_NVJPEG
        if (decode_jpeg_luma(bytes, image)) return true;
        std::cerr << "nvJPEG could not decode " << view_file(dense_folder + "/images", id, ".jpg") << std::endl;
#else
        std::cerr << "built without nvJPEG: provide images/%08d.pgm next to the .jpg files" << std::endl;
#endif
    }
    return false;
}

// The colour image of a view for the fused points (reference: cv::imread(images/%08d.jpg, IMREAD_COLOR), ACMMP.cu:1861-1862):
// images/%08d.ppm (P6, lossless twin like the .pgm for the grey path) or the .jpg through nvJPEG; a view that only has a
// grey image gets (g, g, g).
bool LoadColourImage(const std::string &dense_folder, int id, cv::Mat_<cv::Vec3b> &image)
{
    std::vector<unsigned char> bytes;
    if (read_file(view_file(dense_folder + "/images", id, ".ppm"), bytes) && bytes.size() > 2 && bytes[0] == 'P' && bytes[1] == '6') {
        bytes[1] = '5';                                     // same header grammar as P5
        int w, h;
        size_t off;
        if (parse_pgm(bytes, w, h, off) && bytes.size() >= off + (size_t)w * h * 3) {
            image = cv::Mat_<cv::Vec3b>(h, w);
            for (size_t i = 0; i < (size_t)w * h; ++i)      // file: R, G, B
                image.ptr()[i] = cv::Vec3b(bytes[off + 3 * i + 2], bytes[off + 3 * i + 1], bytes[off + 3 * i]);
            return true;
        }
    }
    if (read_file(view_file(dense_folder + "/images", id, ".pgm"), bytes)) {
        int w, h;
        size_t off;
        if (parse_pgm(bytes, w, h, off)) {
            image = cv::Mat_<cv::Vec3b>(h, w);
            for (size_t i = 0; i < (size_t)w * h; ++i) image.ptr()[i] = cv::Vec3b(bytes[off + i], bytes[off + i], bytes[off + i]);
            return true;
        }
    }
#ifdef ACMMP_WITH_NVJPEG
    if (read_file(view_file(dense_folder + "/images", id, ".jpg"), bytes) && decode_jpeg_bgr(bytes, image)) return true;
#endif
    return false;
}

// cv::resize INTER_LINEAR on an 8-bit 3-channel image (RescaleImageAndCamera, ACMMP.cpp:236): OpenCV's fixed-point scheme --
// coefficients rounded to 1/2048, horizontal pass in integers, vertical pass
// ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
void ResizeLinearBgr(const cv::Mat_<cv::Vec3b> &src, cv::Mat_<cv::Vec3b> &dst, int new_cols, int new_rows)
{
    const int sw = src.cols, sh = src.rows;
    dst = cv::Mat_<cv::Vec3b>(new_rows, new_cols);
    const double scale_x = (double)sw / new_cols, scale_y = (double)sh / new_rows;
    auto coeff = [](float f) {
        const float v = f * 2048.0f;
        return (int)std::lrintf(v);
    };
    std::vector<int> xofs(new_cols), xa0(new_cols), xa1(new_cols);
    for (int dx = 0; dx < new_cols; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
        xofs[dx] = sx;
        xa0[dx] = coeff(1.f - fx);
        xa1[dx] = coeff(fx);
    }
    std::vector<int> row0((size_t)new_cols * 3), row1((size_t)new_cols * 3);
    int cached0 = -1, cached1 = -1;
    auto hresize = [&](int sy, std::vector<int> &out) {
        const cv::Vec3b *S = src.ptr() + (size_t)sy * sw;
        for (int dx = 0; dx < new_cols; ++dx) {
            const int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sx;
            for (int k = 0; k < 3; ++k) out[3 * (size_t)dx + k] = S[sx][k] * xa0[dx] + S[sx1][k] * xa1[dx];
        }
    };
    for (int dy = 0; dy < new_rows; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        if (sy < 0) { fy = 0.f; sy = 0; }
        if (sy >= sh - 1) { fy = 0.f; sy = sh - 1; }
        const int sy1 = sy + 1 < sh ? sy + 1 : sy;
        if (cached0 != sy) {
            if (cached1 == sy) { row0.swap(row1); cached0 = sy; cached1 = -1; }
            else { hresize(sy, row0); cached0 = sy; }
        }
        if (cached1 != sy1) { hresize(sy1, row1); cached1 = sy1; }
        const int b0 = coeff(1.f - fy), b1 = coeff(fy);
        cv::Vec3b *D = dst.ptr() + (size_t)dy * new_cols;
        for (int dx = 0; dx < new_cols; ++dx)
            for (int k = 0; k < 3; ++k) {
                const int v = (((b0 * (row0[3 * (size_t)dx + k] >> 4)) >> 16) + ((b1 * (row1[3 * (size_t)dx + k] >> 4)) >> 16) + 2) >> 2;
                D[dx][k] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
    }
}

// the first `limit` bytes of a file (enough for a PGM header and, nearly always, a JPEG's frame header)
static bool read_head(const std::string &path, size_t limit, std::vector<unsigned char> &bytes)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    bytes.resize(limit);
    bytes.resize(std::fread(bytes.data(), 1, limit, f));
    std::fclose(f);
    return !bytes.empty();
}

bool ImageSize(const std::string &dense_folder, int id, int &cols, int &rows)
{
    std::vector<unsigned char> bytes;
    size_t off;
    const std::string pgm = view_file(dense_folder + "/images", id, ".pgm"), jpg = view_file(dense_folder + "/images", id, ".jpg");
    if (read_head(pgm, 128, bytes) && parse_pgm(bytes, cols, rows, off, true)) return true;
    if (read_head(jpg, 256 << 10, bytes) && jpeg_size(bytes, cols, rows)) return true;
    if (read_file(jpg, bytes) && jpeg_size(bytes, cols, rows)) return true;          // a frame header behind a large EXIF block
    return false;
}

// cv::resize INTER_LINEAR on a 1-channel float image: pixel centres at +0.5, source coordinate
// (d + 0.5) * (src / dst) - 0.5, clamped at the borders, horizontal pass then vertical pass.
void ResizeLinear(const cv::Mat_<float> &src, cv::Mat_<float> &dst, int new_cols, int new_rows)
{
    const int sw = src.cols, sh = src.rows;
    dst = cv::Mat_<float>(new_rows, new_cols);
    const double scale_x = (double)sw / new_cols, scale_y = (double)sh / new_rows;
    std::vector<int> xofs(new_cols);
    std::vector<float> xa(new_cols);
    for (int dx = 0; dx < new_cols; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
        xofs[dx] = sx;
        xa[dx] = fx;
    }
    std::vector<float> row0(new_cols), row1(new_cols);
    int cached0 = -1, cached1 = -1;
    auto hresize = [&](int sy, std::vector<float> &out) {
        const float *S = src.ptr() + (size_t)sy * sw;
        for (int dx = 0; dx < new_cols; ++dx) {
            const int sx = xofs[dx];
            const float a1 = xa[dx], a0 = 1.f - a1;
            const float s1 = S[sx + 1 < sw ? sx + 1 : sx];
            out[dx] = S[sx] * a0 + s1 * a1;
        }
    };
    for (int dy = 0; dy < new_rows; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        if (sy < 0) { fy = 0.f; sy = 0; }
        if (sy >= sh - 1) { fy = 0.f; sy = sh - 1; }
        const int sy1 = sy + 1 < sh ? sy + 1 : sy;
        if (cached0 != sy) {
            if (cached1 == sy) { row0.swap(row1); cached0 = sy; cached1 = -1; }
            else { hresize(sy, row0); cached0 = sy; }
        }
        if (cached1 != sy1) { hresize(sy1, row1); cached1 = sy1; }
        const float b1 = fy, b0 = 1.f - fy;
        float *D = dst.ptr() + (size_t)dy * new_cols;
        for (int dx = 0; dx < new_cols; ++dx) D[dx] = row0[dx] * b0 + row1[dx] * b1;
    }
}

// reference ACMMP.cpp:481-534
FILE *OpenPlyForVertexRecords(const std::string &plyFilePath, size_t n_points)
{
    std::cout << "store 3D points to ply file" << std::endl;
    FILE *outputPly = fopen(plyFilePath.c_str(), "wb");
    if (!outputPly) throw std::runtime_error("cannot write " + plyFilePath);
    fprintf(outputPly, "ply\n");
    fprintf(outputPly, "format binary_little_endian 1.0\n");
    fprintf(outputPly, "element vertex %zu\n", n_points);
    fprintf(outputPly, "property float x\n");
    fprintf(outputPly, "property float y\n");
    fprintf(outputPly, "property float z\n");
    fprintf(outputPly, "property float nx\n");
    fprintf(outputPly, "property float ny\n");
    fprintf(outputPly, "property float nz\n");
    fprintf(outputPly, "property uchar red\n");
    fprintf(outputPly, "property uchar green\n");
    fprintf(outputPly, "property uchar blue\n");
    fprintf(outputPly, "end_header\n");
    return outputPly;
}

// the same file from vertex records the device already packed (acmmp_fusion_run_ply)
void StorePlyVertexRecords(const std::string &plyFilePath, const unsigned char *records27, size_t n_points)
{
    FILE *outputPly = OpenPlyForVertexRecords(plyFilePath, n_points);
    const size_t bytes = 27 * n_points;
    const bool ok = bytes == 0 || fwrite(records27, 1, bytes, outputPly) == bytes;
    fclose(outputPly);
    if (!ok) throw std::runtime_error("short write to " + plyFilePath);
}

void StoreColorPlyFileBinaryPointCloud(const std::string &plyFilePath, const std::vector<PointList> &pc)
{
    FILE *outputPly = OpenPlyForVertexRecords(plyFilePath, pc.size());
    // one 27-byte record per point, assembled in memory and written in one go (the reference issues nine fwrite calls per
    // point inside an omp critical section)
    std::vector<unsigned char> buf(pc.size() * 27);
    for (size_t i = 0; i < pc.size(); ++i) {
        const PointList &p = pc[i];
        float3 X = p.coord;
        if (!(X.x < FLT_MAX && X.x > -FLT_MAX) || !(X.y < FLT_MAX && X.y > -FLT_MAX) || !(X.z < FLT_MAX && X.z >= -FLT_MAX)) {
            X.x = 0.0f; X.y = 0.0f; X.z = 0.0f;
        }
        const char b_color = (char)(int)p.color.x, g_color = (char)(int)p.color.y, r_color = (char)(int)p.color.z;
        unsigned char *o = buf.data() + 27 * i;
        std::memcpy(o, &X.x, 4); std::memcpy(o + 4, &X.y, 4); std::memcpy(o + 8, &X.z, 4);
        std::memcpy(o + 12, &p.normal.x, 4); std::memcpy(o + 16, &p.normal.y, 4); std::memcpy(o + 20, &p.normal.z, 4);
        o[24] = (unsigned char)r_color; o[25] = (unsigned char)g_color; o[26] = (unsigned char)b_color;
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), outputPly);
    fclose(outputPly);
}
