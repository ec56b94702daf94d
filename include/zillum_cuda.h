/* zillum_cuda.h — C ABI of the B200 (sm_100a) rendering hot path.
 *
 * This is the drop-in boundary for ZillumGL's GLSL compute-shader path: every entry
 * point below replaces one piece of the reference's OpenGL plumbing.  Citations are
 * relative to the reference tree (HummaWhite/ZillumGL).
 *
 *   scene upload   <- Scene::createGLContext, 13x TextureBuffered::createFromVector
 *                     (src/core/Scene.cpp:245-258) + Sampler::gen{SobolSeq,Noise}Texture
 *                     (src/core/Scene.cpp:260-264) + EnvironmentMap ctor
 *                     (src/core/EnvironmentMap.cpp:8-59)
 *   render params  <- the ~25 by-name uniforms of updateUniforms()
 *                     (src/integrator/NaivePath.cpp:39-66, LightPath.cpp:40-66,
 *                      TriplePath.cpp:44-77) and uSpp/uFreeCounter (NaivePath.cpp:97-98)
 *   pass launches  <- Pipeline::dispatchCompute (src/core/Pipeline.cpp:72-81) as called
 *                     from NaivePath.cpp:100, LightPath.cpp:104, TriplePath.cpp:120,126
 *   film ops       <- util/img_clear_*.glsl, util/img_copy_1x32f_4x32f.glsl,
 *                     Texture2D film objects (LightPath.cpp:7-15)
 *
 * Conventions: plain pointers and sizes only; host arrays are borrowed for the duration
 * of the call; every function returns 0 on success or a non-zero error code (a
 * cudaError_t value, or ZL_ERR_*), with zl_last_error_string() giving the text.  Nothing
 * in the library aborts.  Handles are not thread-safe; use one host thread per GPU.
 * `stream` arguments are cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef ZILLUM_CUDA_H
#define ZILLUM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZL_ABI_VERSION 3   /* v3: ZlSceneDesc::sizeIndices (device-side MTBVH threading), zl_film_postprocess, zl_scene_read_nodes */

enum {
    ZL_OK = 0,
    ZL_ERR_INVALID_ARGUMENT = 10001,
    ZL_ERR_NO_DEVICE = 10002,
    ZL_ERR_OUT_OF_MEMORY = 10003
};

/* Material types, same numbering as Material.h:35 / material.glsl:12-17. */
enum { ZL_MAT_LAMBERTIAN = 0, ZL_MAT_PRINCIPLED = 1, ZL_MAT_METAL_WORKFLOW = 2,
       ZL_MAT_DIELECTRIC = 3, ZL_MAT_THIN_DIELECTRIC = 4 };

/* Host-side description of a flattened scene: exactly the arrays the reference uploads
 * as buffer textures.  All pointers are HOST pointers.              */
typedef struct ZlSceneDesc {
    /* geometry — Scene.cpp:148-195,245-248 */
    const float*    vertices;      /* 3*numVertices, world space, objects then lights   */
    const float*    normals;       /* 3*numVertices                                     */
    const float*    texcoords;     /* 2*numTexcoords (object vertices only; may be 0)   */
    const uint32_t* indices;       /* 3*numTriangles, global vertex ids                 */
    /* MTBVH — PackedBVH, BVH.h:13-17, BVH.cpp:298-346 */
    const float*    bounds;        /* 6*bvhSize: pMin.xyz,pMax.xyz per node (pre-order); NULL: the BVH is built on the device */
    const int32_t*  hitTable;      /* 6 faces * bvhSize * (node, prim|-1, miss); NULL: see sizeIndices */
    /* materials — Scene.cpp:251-252, Material.h:32-53 */
    const int32_t*  matTexIndices; /* objPrimCount: (texId<<16 | matId), texId -1 = none*/
    const float*    materials;     /* 16 floats (4 texels) per material                 */
    /* area lights — Scene.cpp:200-243,253-255 */
    const float*    lightPower;    /* 3*numLightTriangles                               */
    const int32_t*  lightAlias;    /* numLightTriangles                                 */
    const float*    lightProb;     /* numLightTriangles                                 */
    /* albedo texture array — Texture.cpp:134-171 (sRGB8, layers padded to max size)    */
    const uint8_t*  texels;        /* numTextures * texMaxH * texMaxW * 3, may be NULL  */
    const float*    texUVScale;    /* 2*numTextures                                     */
    /* environment map — EnvironmentMap.cpp:8-59; rounded to RGB16F on upload           */
    const float*    envMap;        /* 3*envW*envH float RGB, row 0 = +Z pole; may be NULL*/
    const int32_t*  envAlias;      /* (envW+1)*envH, column envW = row marginal         */
    const float*    envAliasProb;  /* (envW+1)*envH                                     */
    /* sampler — Sampler.cpp:48-80, SobolMatrices256x32.h */
    const float*    noise;         /* 2*noiseW*noiseH, RG32F seed image                 */
    const uint32_t* sobolMatrices; /* 256*32 generator matrix columns                   */

    int32_t numVertices, numTexcoords, numTriangles, bvhSize;
    int32_t objPrimCount, numMaterials, numLightTriangles;
    int32_t numTextures, texMaxW, texMaxH;
    int32_t envW, envH;
    int32_t noiseW, noiseH;
    float   lightSum;              /* Scene::lightSumPdf                                */
    float   envSum;                /* float(int(EnvironmentMap::mSumPdf)), EnvironmentMap.h:22 */
    /* The builder's pre-order tree (BVH::sizeIndices, BVH.h:39 / BVH.cpp:217-296): subtree node count, or
     * primIndex | 0x80000000 for a leaf.  Optional.  When hitTable is NULL and sizeIndices is given, the six
     * threaded orderings of BVH::buildHitTable (BVH.cpp:298-346) are computed ON THE DEVICE from bounds +
     * sizeIndices: no 18*bvhSize-int table is built or uploaded.                              */
    const int32_t*  sizeIndices;   /* bvhSize, may be NULL when hitTable is given                           */
} ZlSceneDesc;

/* camera.glsl:5-14 uniforms, produced by Camera::update (Camera.cpp:149-162). */
typedef struct ZlCamera {
    float F[3], R[3], U[3];
    float matInv[9];               /* inverse(mat3(R,U,F)), column-major (NaivePath.cpp:53-54) */
    float pos[3];
    float tanFOV, asp, lensRadius, focalDist;
} ZlCamera;

/* The per-integrator uniforms, one POD passed by pointer to the launch shims. */
typedef struct ZlRenderParams {
    ZlCamera camera;
    int32_t filmW, filmH;          /* uFilmSize                                         */
    int32_t maxDepth;              /* uMaxDepth                                         */
    int32_t russianRoulette;       /* uRussianRoulette                                  */
    int32_t sampleLight;           /* uSampleLight (path only)                          */
    int32_t lightEnvUniformSample; /* uLightEnvUniformSample                            */
    float   lightPortion;          /* uLightSamplePortion                               */
    int32_t sampler;               /* uSampler: 0 = hash RNG, 1 = Sobol                 */
    float   envRotation;           /* uEnvRotation                                      */
    int32_t spp;                   /* uSpp = pass index                                 */
    int32_t freeCounter;           /* uFreeCounter                                      */
    int32_t blocksOnePass;         /* uBlocksOnePass (light, triple-LPT)                */
    int32_t loopsPerPass;          /* uLoopsPerPass (triple-LPT)                        */
    float   scale;                 /* uScale (triple-LPT), TriplePath.cpp:76-77         */
} ZlRenderParams;

#define ZL_LIGHT_GROUP_SIZE 1536   /* gl_WorkGroupSize.x of the light kernels (LightPath.cpp:3) */

typedef struct ZlScene ZlScene;    /* opaque device scene                               */
typedef struct ZlFilm  ZlFilm;     /* opaque device film: W*H float4 (rgb sum, w unused) */

/* ---- library ---- */
int         zl_abi_version(void);
const char* zl_last_error_string(void);
int         zl_device_count(int* count);
int         zl_set_device(int device);
int         zl_device_synchronize(void);

/* ---- scene (replaces Scene::createGLContext uploads) ---- */
int zl_scene_create(const ZlSceneDesc* desc, ZlScene** out);
int zl_scene_destroy(ZlScene* scene);
/* mirrors glContext.material->write(...) in src/gui/Editor.cpp:73 */
int zl_scene_update_materials(ZlScene* scene, int first, int count, const float* materials);
/* BVH::build (src/accelerator/BVH.cpp:116-144, 217-296) on the device: 16-bucket binned SAH, one triangle per leaf, pre-order
 * nodes; the reference's tree (level-synchronous build, csrc/zl_bvh_build.cuh).  Host arrays in and out:
 * boundsOut 6*(2T-1) floats, sizeIndicesOut 2T-1 ints (BVH.h:39), levelsOut (optional) the number of levels run.
 * zl_scene_create does the same internally when ZlSceneDesc::bounds is NULL.                                    */
int zl_build_bvh(const float* vertices, int numVertices, const uint32_t* indices, int numTriangles,
                 float* boundsOut, int32_t* sizeIndicesOut, int* levelsOut);
/* Read back threaded node records [first, first+count) of MTBVH face 0..5 as the reference's texels: per entry
 * 6 floats of bounds (pMin, pMax) into boundsOut and (primIndex | -1, missIndex) into linksOut.  For tests.   */
int zl_scene_read_nodes(const ZlScene* scene, int face, size_t first, size_t count, float* boundsOut, int32_t* linksOut);
/* wall time of the device-side preparation done by zl_scene_create (0 where the host supplied the data): BVH build,
 * MTBVH threading, and the number of levels the build ran */
int zl_scene_prep_times(const ZlScene* scene, double* bvhBuildMs, double* mtbvhThreadMs, int* bvhLevels);
/* milliseconds this scene creation spent on the one-time lazy loading of the device-build kernels (0 unless it was the first device
 * build of the process); kept apart from bvhBuildMs */
double zl_scene_cuda_init_ms(const ZlScene* scene);
/* bytes of device memory held by the scene, and by the MTBVH node records alone */
int zl_scene_memory(const ZlScene* scene, size_t* totalBytes, size_t* nodeBytes);

/* ---- film (replaces the rgba32f frame / r32f 3WxH film textures) ---- */
int zl_film_create(int width, int height, ZlFilm** out);
/* film living in caller-owned device memory (width*height*4 floats), e.g. a torch tensor */
int zl_film_create_external(int width, int height, void* devicePtr, ZlFilm** out);
int zl_film_destroy(ZlFilm* film);
int zl_film_clear(ZlFilm* film, void* stream);                  /* util/img_clear_*.glsl */
/* After zl_launch_path_pass(..., variant 2, stream) passes run on internal streams, two in flight.  Every zl_film_* call orders
 * itself behind them; zl_film_flush makes `stream` wait for them, for callers that use zl_film_device_ptr() memory directly
 * (e.g. an NCCL all-reduce of a torch-owned film) or that time the passes with events on `stream`.                       */
int zl_film_flush(ZlFilm* film, void* stream);
void* zl_film_device_ptr(ZlFilm* film);
/* rgba32f W*H frame on the host; rgb = sum * scale, a = 1 (img_copy_1x32f_4x32f.glsl) */
int zl_film_download(ZlFilm* film, float scale, float* rgbaHost, void* stream);
/* Pipelined form: the resolve runs on `stream` (on the film stream while variant-2 passes are in flight), the
 * device->host copy on an internal copy stream, so the passes launched next overlap the copy.  rgbaHostPinned must
 * be page-locked host memory and stay valid until the zl_film_download_wait() that completes it returns.  Four
 * read-backs may be in flight per film (own staging buffers, allocated on first use); zl_film_download_wait() blocks
 * until the OLDEST one is complete (no-op with none in flight); a fifth call takes over the oldest slot (it first waits,
 * on the device, for that slot's copy).  The frame is resolved into its staging buffer at once; on a film that passes
 * are launched on, the cudaMemcpyAsync call itself is issued behind the NEXT zl_launch_*_pass (or by the wait): some
 * hosts do not return from it before the copy is complete, and the next pass must already be queued by then.        */
int zl_film_download_async(ZlFilm* film, float scale, float* rgbaHostPinned, void* stream);
/* the same read-back without the frame's constant alpha: packed RGB, W*H*3 floats (12 bytes per pixel over PCIe instead of 16) */
int zl_film_download_rgb_async(ZlFilm* film, float scale, float* rgbHostPinned, void* stream);
int zl_film_download_wait(ZlFilm* film);
/* Multi-GPU frames, reduce before copy: a consistent device-side copy of the film (w*h*4 floats) into caller memory, ordered like the
 * read-backs above (behind every pass launched so far, ahead of later ones) and after the work already queued on `stream`; `stream`
 * waits for the copy.  The caller then reduce-scatters the copy over NVLink (ncclReduceScatter / torch.distributed) and each rank
 * reads back only its rows through an external film over the reduced slice (zl_film_create_external + zl_film_download_rgb_async):
 * N ranks move one frame over PCIe per step instead of N (bench.py, e2e at N > 1). */
int zl_film_snapshot_async(ZlFilm* film, void* dstDevice, void* stream);
/* Display stage (src/shader/post_proc.glsl:12-59, dispatched by Application.cpp:644-663): rgb = film * resultScale,
 * clamped to [0, 1e30], tone mapped (0 = none, 1 = filmic [reference default, Application.cpp:98], 2 = ACES), gamma 1/2.2.
 * rgbaHost (W*H*4 floats, a = 1) is the reference's rgba32f result texture; rgb8Host (W*H*3 bytes) its
 * GL_UNSIGNED_BYTE read-back for screenshots (Texture2D::readFromDevice, Texture.cpp:96-102).  Either may be NULL.
 * Rows are in film order (row 0 = bottom); the reference flips on write (Application.cpp:376).                  */
int zl_film_postprocess(ZlFilm* film, float resultScale, int toneMapper, float* rgbaHost, unsigned char* rgb8Host, void* stream);
/* in-place sum over all ranks of an NCCL communicator (ncclComm_t passed as void*) */
int zl_film_allreduce(ZlFilm* film, void* ncclComm, void* stream);

/* ---- pass launches (replace Pipeline::dispatchCompute of the four integrator kernels) ---- */
/* variant: 0 = megakernel (one thread per path; camera kernels: warp = 8x4 pixel tile),
 *          1 = wavefront: generate / shade-per-material-type / sort / trace / resolve stages over
 *              device-side queues (csrc/zl_wavefront*.cuh).  Same arithmetic per path; the path
 *              and the triple camera pass are bit-identical to variant 0, splat passes differ by
 *              atomic summation order.
 *          2 = variant 1 with several passes in flight on internal streams (path tracer: three, four for films below 2^20
 *              pixels, one workspace each; light / triple: three).  Path: all film writes and reads ordered
 *              on one film stream, film bit-identical to variants 0 / 1.  Light / triple: splats are atomics, kept
 *              apart from the resolve kernels' plain adds; a frame read is ordered between whole passes (for the
 *              triple tracer a pass = the camera pass followed by its light pass).  See zl_film_flush.
 *          3 = variant 1 replayed as a CUDA graph: the first pass with a configuration runs with plain launches, the second is
 *              stream-captured, later ones are one cudaGraphLaunch each (uSpp / uFreeCounter come from a device pair the graph's
 *              first node writes).  For films whose pass is shorter than the host can issue its 30-60 launches.  Same kernels,
 *              order and arguments as variant 1: bit-identical films.                                                          */
int zl_launch_path_pass      (ZlScene*, ZlFilm*, const ZlRenderParams*, int variant, void* stream);
int zl_launch_light_pass     (ZlScene*, ZlFilm*, const ZlRenderParams*, int variant, void* stream);
int zl_launch_triple_pt_pass (ZlScene*, ZlFilm*, const ZlRenderParams*, int variant, void* stream);
int zl_launch_triple_lpt_pass(ZlScene*, ZlFilm*, const ZlRenderParams*, int variant, void* stream);

/* Instrumented pass (same arithmetic, separately compiled with visit counters): renders one
 * pass of kind 0 = path, 1 = light, 2 = triple-PT, 3 = triple-LPT into `film` and returns
 * counters6 = {rays, hit-table entries visited, leaf triangle tests, shading points,
 * splats, paths}.  For the roofline byte model; never used in a timed region.            */
int zl_counted_pass(ZlScene*, ZlFilm*, const ZlRenderParams*, int kind, unsigned long long* counters6);
/* Of the last zl_counted_pass (kind 0): the shadow rays the reference casts (they are part of counters6) but the production
 * wavefront pass does not trace — NEE samples rejected after the visibility test or whose contribution is exactly zero
 * (csrc/zl_wavefront.cuh, wfShadeKernel).  counters3 = {rays, hit-table entries, triangle tests}. */
int zl_counted_pass_untraced(unsigned long long* counters3);

/* ---- traversal on an explicit ray set (ID parity test and the Mrays/s metric) ----
 * rays: n * 6 floats (ori.xyz, dir.xyz) on the HOST; anyhit=0 -> bvhHit semantics
 * (intersection.glsl:395-427: outIds = closest prim or -1, outT = dist or 1e8);
 * anyhit=1 -> bvhTest (intersection.glsl:367-393) with per-ray max distance tMax[i]
 * (NULL = 1e8): outIds = 1 if occluded else 0.  outSteps (may be NULL) receives the
 * bvhDebug-style visit counters (intersection.glsl:331-365): 2 ints per ray =
 * (hit-table entries visited, leaf triangle tests).                                   */
int zl_trace_rays(ZlScene*, const float* rays, size_t n, int anyhit, const float* tMax,
                  int32_t* outIds, float* outT, int32_t* outSteps);

/* Device-resident ray-set interface for benchmarking (no host copies in the timed call). */
typedef struct ZlRaySet ZlRaySet;
int zl_rayset_create(const float* raysHost, size_t n, ZlRaySet** out);
int zl_rayset_destroy(ZlRaySet*);
int zl_rayset_trace(ZlScene*, ZlRaySet*, int anyhit, int variant, void* stream);
int zl_rayset_download(ZlRaySet*, int32_t* outIds, float* outT);
int zl_rayset_set_tmax(ZlRaySet*, const float* tMaxHost);      /* per-ray max distance for anyhit */
size_t zl_rayset_size(ZlRaySet*);
int zl_rayset_download_rays(ZlRaySet*, float* raysHost);       /* n*6 floats */
/* pixel-centre primary rays (thinLensCameraSampleRay with u = 0, camera.glsl:62-77) */
int zl_rayset_create_primary(const ZlRenderParams*, ZlRaySet** out);

/* ---- per-function evaluation for known-answer tests (device functions on arrays) ----
 * op selects a device function of the shading library; in/out are HOST arrays of
 * n*inStride / n*outStride floats.  The op table follows the declaration.                      */
int zl_debug_eval(ZlScene*, const ZlRenderParams*, int op, const float* in, int inStride,
                  float* out, int outStride, size_t n);
/* op table: integers travel as raw bit patterns inside the float arrays ("b:" below).
 *  op                        inputs                                     outputs
 *  ZL_KAT_HASH               b:seed                                     b:hash            random.glsl:5-13
 *  ZL_KAT_SOBOL              b:index b:dim                              b:value           Sampler.cpp:19-28
 *  ZL_KAT_CUBEMAP_FACE       dir3                                       b:face            math.glsl:125-131
 *  ZL_KAT_BOXHIT             b:k ray6 (k = entry of the ray's face)     hit tMin          intersection.glsl:226-329
 *  ZL_KAT_TRIANGLE           b:tri ray6                                 hit t             intersection.glsl:63-121
 *  ZL_KAT_SURFACE            b:tri p3                                   ns3 ng3 uv2       intersection.glsl:188-224
 *  ZL_KAT_CAMERA_RAY         uv2 u4                                     ori3 dir3         camera.glsl:62-77
 *  ZL_KAT_CAMERA_II          ref3 u2                                    wi3 Ii3 dist uv2 pdf   camera.glsl:109-127
 *  ZL_KAT_CAMERA_PDF         ray6                                       pdfPos pdfDir     camera.glsl:129-142
 *  ZL_KAT_BSDF_EVAL          b:mat b:tex uv2 wo3 wi3 n3 b:mode          bsdf3 pdf         material_loader.glsl:99-151
 *  ZL_KAT_BSDF_SAMPLE        b:mat b:tex uv2 wo3 n3 b:mode u3 b:seed    wi3 pdf bsdf3 eta b:flag   material_loader.glsl:153-169
 *  ZL_KAT_ENV_LE             wi3                                        rgb3 pdfLi        light.glsl:163-179
 *  ZL_KAT_ENV_SAMPLE         u4                                         wi3 pdf           light.glsl:181-205
 *  ZL_KAT_LIGHT_LE           b:light x3 wo3 y3                          rgb3 pdfLi(x<-y)  light.glsl:79-98
 *  ZL_KAT_LIGHT_SAMPLE_LE    b:light u4                                 ray6 Le3 pdfPos pdfDir  light.glsl:111-120
 *  ZL_KAT_SAMPLE_LIGHT_ENV   x3 ud us4                                  wi3 coef3 pdf     light.glsl:221-235
 *  ZL_KAT_LIBM               b:fn x y  (fn 0 sin 1 cos 2 atan(y,x) 3 asin 4 acos 5 log 6 pow(x,y) 7 exp)   value   include/zl_libm.h (the GLSL built-ins)
 */
enum { ZL_KAT_HASH = 0, ZL_KAT_SOBOL, ZL_KAT_CUBEMAP_FACE, ZL_KAT_BOXHIT, ZL_KAT_TRIANGLE,
       ZL_KAT_SURFACE, ZL_KAT_CAMERA_RAY, ZL_KAT_CAMERA_II, ZL_KAT_CAMERA_PDF, ZL_KAT_BSDF_EVAL,
       ZL_KAT_BSDF_SAMPLE, ZL_KAT_ENV_LE, ZL_KAT_ENV_SAMPLE, ZL_KAT_LIGHT_LE,
       ZL_KAT_LIGHT_SAMPLE_LE, ZL_KAT_SAMPLE_LIGHT_ENV, ZL_KAT_LIBM, ZL_KAT_COUNT };

/* Measurement aid (bench.py traversal micro-benchmark): over the closest-hit walks of a ray set, out4 = {node visits summed over lanes,
 * triangle tests summed over lanes, DISTINCT node records per warp-step summed over steps, distinct triangles per warp-step}.  The first
 * pair is the algorithmic work; the second is what a warp in lock step really requests from the memory system. */
int zl_rayset_unique_sectors(ZlScene*, ZlRaySet*, unsigned long long* out4);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
unsigned long long zl_launch_count(void);

/* Per-stage device timing of the pass launchers (measurement aid; replaces nothing in the reference, whose
 * only timing is the frame counter of Application.cpp:644-663).  While enabled, every group of launches of a
 * pass is bracketed by CUDA events on the launch stream.  zl_stage_timing_read synchronises the device,
 * returns the summed milliseconds and launch counts per stage since the last read, and clears them. */
enum { ZL_STAGE_GENERATE = 0,   /* camera / emission sampling (wf*GenerateKernel) */
       ZL_STAGE_SHADE,          /* per-material shade kernels */
       ZL_STAGE_SORT,           /* ray-queue counting sort */
       ZL_STAGE_TRACE,          /* MTBVH traversal of the shadow + extension queues (the dominant kernel) */
       ZL_STAGE_RESOLVE,        /* ended paths -> film */
       ZL_STAGE_MEGAKERNEL,     /* variant 0: one kernel per pass */
       ZL_STAGE_COUNT };
int zl_stage_timing_enable(int enable);
int zl_stage_timing_read(double* msPerStage /* [ZL_STAGE_COUNT] */, unsigned long long* launchesPerStage /* [ZL_STAGE_COUNT] */);

/* measured L2 / DRAM read bandwidth of a streaming read kernel over `bytes` (GB/s) */
int zl_measure_read_bandwidth(size_t bytes, int iters, double* gbPerSec);

#ifdef __cplusplus
}
#endif
#endif /* ZILLUM_CUDA_H */
