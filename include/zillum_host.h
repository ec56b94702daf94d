/* zillum_host.h — C surface of the C++ host library (libzillum_host.so).
 *
 * The host side of this framework is C++ and mirrors the reference's own classes
 * (Scene, Camera, BVH, EnvironmentMap, Integrator, NaivePathIntegrator, ... — see
 * zillumgl_b200/host/).  This header is only the thin handle-based shim that lets the
 * Python test / benchmark harness (ctypes) drive those classes; C++ users include the
 * class headers directly, exactly as Application.cpp does with the reference's
 * (src/Application.cpp:336-356, 644-663).
 */
#ifndef ZILLUM_HOST_H
#define ZILLUM_HOST_H
#include "zillum_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ZhScene ZhScene;
typedef struct ZhIntegrator ZhIntegrator;

/* ---- Scene (src/core/Scene.h) ---- */
ZhScene* zh_scene_create(void);
void     zh_scene_destroy(ZhScene*);
int      zh_scene_load(ZhScene*, const char* xmlPath);                        /* Scene::load            */
int      zh_scene_load_builtin(ZhScene*, const char* name, int w, int h);     /* synthetic configs      */
int      zh_scene_load_xml_text(ZhScene*, const char* xml);
int      zh_scene_flatten(ZhScene*);                /* host half of Scene::createGLContext (BVH, tables) */
int      zh_scene_upload(ZhScene*);                 /* device half: zl_scene_create                      */
const ZlSceneDesc* zh_scene_desc(ZhScene*);         /* valid until the next flatten/load                 */
ZlScene* zh_scene_device(ZhScene*);
/* info[16]: numVertices, numTriangles, bvhSize, objPrimCount, nLightTriangles, numMaterials,
 *           filmW, filmH, sampler, numTextures, envW, envH, numLightMeshes                 */
void     zh_scene_info(ZhScene*, int* info);
/* times[3]: BVH build s, MTBVH flatten s, whole flatten() s */
void     zh_scene_times(ZhScene*, double* times);
void     zh_scene_light_meshes(ZhScene*, int* firstTri, int* numTris, float* power3);
void     zh_scene_set_camera(ZhScene*, const float* pos3, const float* angleDeg3, float fovDeg, float lensRadius, float focalDist);
void     zh_scene_camera(ZhScene*, ZlCamera* out);
void     zh_scene_set_sampler(ZhScene*, int sampler);
/* 1: skip the host MTBVH flatten; zl_scene_create threads the six orderings on the device (call before flatten) */
void     zh_scene_set_device_mtbvh(ZhScene*, int on);
/* 1: no host BVH at all; zl_scene_create builds the reference's tree on the device and threads it (call before flatten) */
void     zh_scene_set_device_bvh(ZhScene*, int on);
void     zh_scene_set_env_rotation(ZhScene*, float radians);
/* the scene before flattening (objects first, then lights) — test accessors: tests/test_ref_parity.py hands the same
 * model instances to the reference's own Scene::load / createGLContext (oracle/_ref) and compares the flattened arrays.
 * info[3]: isLight, numMeshes, numMaterials; trs9: translate, scale, rotate (Model.h); counts[4]: vertices, indices, texIndex, matIndex */
int      zh_scene_num_models(ZhScene*);
void     zh_scene_model_info(ZhScene*, int model, int* info, float* trs9, float* power3, char* pathOut, int pathCap);
void     zh_scene_model_mesh_counts(ZhScene*, int model, int mesh, int* counts);
void     zh_scene_model_mesh_data(ZhScene*, int model, int mesh, float* pos, float* nrm, float* tex, uint32_t* idx);
void     zh_scene_model_materials(ZhScene*, int model, float* mats16);
int      zh_num_images(void);                                        /* Resource::getAllImages() */
void     zh_image(int index, int* w, int* h, unsigned char* rgb8);   /* rgb8 may be NULL (size query) */
const char* zh_builtin_scene_xml(const char* name, int w, int h);             /* static buffer */

/* ---- Integrators (src/core/Integrator.h) ---- */
/* type: "path" (NaivePathIntegrator) | "light" (LightPathIntegrator) | "triple" (TriplePathIntegrator).
 * externalFilm: NULL, or device memory of w*h*4 floats owned by the caller.              */
ZhIntegrator* zh_integrator_create(const char* type, ZhScene*, int w, int h, void* externalFilm, void* stream);
void     zh_integrator_destroy(ZhIntegrator*);
/* parameters by the names of the reference's param structs: maxDepth, russianRoulette,
 * sampleLight, lightEnvUniformSample, lightPortion, finiteSample, maxSample,
 * threadBlocksOnePass, LPTBlocksOnePass, LPTLoopsPerPass, kernelVariant                   */
int      zh_integrator_set(ZhIntegrator*, const char* name, double value);
double   zh_integrator_get(ZhIntegrator*, const char* name);
void     zh_integrator_set_sample_shard(ZhIntegrator*, int first, int stride);
int      zh_integrator_render_one_pass(ZhIntegrator*);                        /* Integrator::renderOnePass; 0 or the launch error (Integrator::lastError) */
void     zh_integrator_reset(ZhIntegrator*);                                  /* Integrator::reset         */
void     zh_integrator_params(ZhIntegrator*, int kernel, ZlRenderParams* out);/* uniforms of the next pass */
ZlFilm*  zh_integrator_film(ZhIntegrator*);
float    zh_integrator_result_scale(ZhIntegrator*);                           /* reference semantics       */
float    zh_integrator_true_scale(ZhIntegrator*);                             /* 1 / true sample count     */
int      zh_integrator_cur_sample(ZhIntegrator*);
/* rgba: w*h*4 floats = film * scale (scale <= 0: use true_scale) */
int      zh_integrator_get_frame(ZhIntegrator*, float scale, float* rgba);
/* pipelined read-back (Integrator::getFrameAsync / waitFrame): rgbaPinned is page-locked host memory */
int      zh_integrator_get_frame_async(ZhIntegrator*, float scale, float* rgbaPinned);
int      zh_integrator_get_frame_rgb_async(ZhIntegrator*, float scale, float* rgbPinned);   /* packed RGB: w*h*3 floats */
int      zh_integrator_wait_frame(ZhIntegrator*);
/* kernelVariant 2 (two passes in flight): make the integrator's stream wait for them (Integrator::flush) */
int      zh_integrator_flush(ZhIntegrator*);
/* Integrator::snapshotAsync: consistent device-side copy of the film (w*h*4 floats) for a reduce-before-copy frame path over several GPUs */
int      zh_integrator_snapshot_async(ZhIntegrator*, void* dstDevice);

/* ---- host preparation exposed for tests (oracle cross-checks) ---- */
int      zh_build_bvh(const float* vertices, int numVertices, const uint32_t* indices, int numTriangles,
                      float* boundsOut, int32_t* hitTableOut, double* seconds2);
void     zh_alias_table(const float* pdf, int n, int32_t* alias, float* prob);
float    zh_env_tables(const float* rgb, int w, int h, int32_t* alias, float* prob);
uint32_t zh_sobol_sample(uint32_t index, int dim);
void     zh_noise_texture(int w, int h, float* out);

/* ---- image output (headless EXR / PFM) ---- */
int      zh_write_pfm(const char* path, const float* rgba, int w, int h);
int      zh_write_exr(const char* path, const float* rgba, int w, int h);
/* 8-bit RGB, rows in film order (row 0 = bottom), flipped on write like the reference's screenshot (Application.cpp:371-380) */
int      zh_write_png(const char* path, const unsigned char* rgb8, int w, int h);
/* ---- image input: 8-bit RGB albedo textures, the stbi_load(path, &w, &h, &n, 3) of src/core/Image.cpp:10-34 (PNG, JPEG, TGA, BMP, PPM).
 * First call with rgb8 = NULL for the size; row 0 = top of the image.  Non-zero: unreadable or unsupported file. */
int      zh_load_byte_image(const char* path, int* w, int* h, unsigned char* rgb8);
/* display stage of the reference's frame loop (post_proc.glsl via Application.cpp:644-663): tone-mapped, gamma-encoded frame.
 * scale <= 0: 1 / true sample count.  toneMapper 0 none, 1 filmic (reference default), 2 ACES.  rgba / rgb8 may be NULL. */
int      zh_integrator_post_process(ZhIntegrator*, float scale, int toneMapper, float* rgba, unsigned char* rgb8);

#ifdef __cplusplus
}
#endif
#endif
