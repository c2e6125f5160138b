/* oracle/acmmp_oracle.h -- TEST INFRASTRUCTURE ONLY (see acmmp_oracle.c). */
#ifndef ACMMP_ORACLE_H_
#define ACMMP_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference `struct Camera`, main.h:40-54 */
typedef struct {
    int32_t model;
    float params[4];
    float R[9];
    float t[3];
    float K[9];
    int32_t width, height;
    float depth_min, depth_max;
} orc_camera;

typedef struct {
    const float *data;
    int width, height;
} orc_image;

int orc_num_threads(void);
float orc_tex2d(const orc_image *img, float x, float y);
void orc_pixel_to_dir(const orc_camera *cam, int x, int y, float dir[3]);
float orc_depth_from_plane(const orc_camera *cam, const float plane[4], int x, int y);
void orc_point_on_world(float x, float y, float depth, const orc_camera *cam, float X[3]);
void orc_project(const float X[3], const orc_camera *cam, float pt[2], float *depth);
float orc_bilateral_ncc(const orc_image *ref_img, const orc_camera *ref_cam, const orc_image *src_img,
                        const orc_camera *src_cam, int x, int y, const float plane[4]);
float orc_geom_cost(const orc_image *depth_img, const orc_camera *ref_cam, const orc_camera *src_cam,
                    const float plane[4], int x, int y);
float orc_init_cost(int n_images, const orc_image *imgs, const orc_camera *cams, int x, int y, const float plane[4],
                    uint32_t *selected_views);

/* whole-map drivers (OpenMP over rows) */
void orc_ncc_map(const orc_image *ref_img, const orc_camera *ref_cam, const orc_image *src_img, const orc_camera *src_cam,
                 const float *planes4, float *out);
void orc_geom_map(const orc_image *depth_img, const orc_camera *ref_cam, const orc_camera *src_cam, const float *planes4,
                  float *out);
void orc_warp_map(const orc_camera *ref_cam, const orc_camera *src_cam, const float *planes4, float *out4);
void orc_initcost_map(int n_images, const orc_image *imgs, const orc_camera *cams, const float *planes4, float *out,
                      uint32_t *views);
void orc_jbu(const float *image, int cols, int rows, const float *depth, int s_width, int s_height, float *out);
void orc_depth_normal(const orc_camera *cam, float *planes4);
void orc_median_filter(int width, int height, float *planes4, const float *costs, int colour);

/* cuRAND XORWOW: state = {d, v0..v4} */
void orc_curand_init(uint64_t seed, uint64_t subsequence, uint64_t offset, uint32_t state[6]);
uint32_t orc_curand(uint32_t state[6]);
float orc_curand_uniform(uint32_t state[6]);
/* RandomInitialization branch (i) (ACMMP.cu:686-689): planes, costs, views, rand states for the whole map */
void orc_random_init(int n_images, const orc_image *imgs, const orc_camera *cams, float depth_min, float depth_max,
                     uint64_t seed, float *planes4, float *costs, uint32_t *views, uint32_t *rand6, int with_costs);

/* one checkerboard pass (CheckerboardPropagation + PlaneHypothesisRefinement, ACMMP.cu:797-1325) with
 * read-old / write-new neighbour semantics.  Any of depth_imgs / prior_planes / plane_masks / pre_costs
 * may be NULL when the corresponding flag is 0. */
typedef struct {
    int geom, prior, hierarchy, as_compiled;
    float depth_min, depth_max;
} orc_pass_flags;
void orc_checkerboard_pass(int n_images, const orc_image *imgs, const orc_image *depth_imgs, const orc_camera *cams,
                           const orc_pass_flags *flags, int colour, int iter, const float *planes_in,
                           const float *costs_in, float *planes_out, float *costs_out, const float *pre_costs,
                           uint32_t *selected_views, uint32_t *rand6, const float *prior_planes4,
                           const uint32_t *plane_masks);

/* planar-prior pass only: emulate the other extreme of the reference's data race (see acmmp_oracle.c) */
/* SimpleFusionKernel, ACMMP.cu:1664-1814, one reference view; points = w*h*9 floats (coord, normal, color in the kernel's
 * B, G, R order), flags = w*h ints; bgr (or bgr[i]) may be NULL: the grey level serves all three channels */
void orc_fuse_view(int n_views, const orc_camera *cams, const float *const *depths, const float *const *normals3,
                   const float *const *gray, const unsigned char *const *bgr, int ref, int n_src, const int *src_idx,
                   float *points, int *flags);
void orc_set_race_emulation(const float *late_planes, float *center_planes_out);

#ifdef __cplusplus
}
#endif
#endif
