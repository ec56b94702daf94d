#include "Integrator.h"
#include <cstdio>

namespace zillum {

Integrator::~Integrator() { if (mFilm) zl_film_destroy(mFilm); }

void Integrator::recreateFrameTex(int width, int height) {
    if (mFilm) { zl_film_destroy(mFilm); mFilm = nullptr; }
    int rc = mExternalFilm ? zl_film_create_external(width, height, mExternalFilm, &mFilm) : zl_film_create(width, height, &mFilm);
    if (rc != 0) std::fprintf(stderr, "[Integrator] film allocation failed: %s\n", zl_last_error_string());
    if (mFilm) zl_film_clear(mFilm, mStream);
    mFrameW = width; mFrameH = height;
    mFrame.clear();        // the host copy is sized by the first getFrame(): 132 MB of page faults at 3840x2160 that callers of the async read-backs never need
    mFrame.shrink_to_fit();
}

const std::vector<float>& Integrator::getFrame() {
    mFrame.resize((size_t)mFrameW * mFrameH * 4, 0.0f);
    if (mFilm) zl_film_download(mFilm, resultScale(), mFrame.data(), mStream);
    return mFrame;
}

int Integrator::downloadFrame(float scale, float* rgba) { return mFilm ? zl_film_download(mFilm, scale, rgba, mStream) : ZL_ERR_INVALID_ARGUMENT; }

int Integrator::snapshotAsync(void* dstDevice) { return mFilm ? zl_film_snapshot_async(mFilm, dstDevice, mStream) : ZL_ERR_INVALID_ARGUMENT; }

int Integrator::getFrameAsync(float* dstPinned, float scale, int channels) {
    if (!mFilm) return ZL_ERR_INVALID_ARGUMENT;
    const float sc = scale > 0.0f ? scale : trueScale();
    return channels == 3 ? zl_film_download_rgb_async(mFilm, sc, dstPinned, mStream) : zl_film_download_async(mFilm, sc, dstPinned, mStream);
}
int Integrator::postProcess(float scale, int toneMapper, float* rgba, unsigned char* rgb8) {
    if (!mFilm) return ZL_ERR_INVALID_ARGUMENT;
    return zl_film_postprocess(mFilm, scale > 0.0f ? scale : trueScale(), toneMapper, rgba, rgb8, mStream);
}
int Integrator::flush() { return mFilm ? zl_film_flush(mFilm, mStream) : ZL_ERR_INVALID_ARGUMENT; }
int Integrator::waitFrame() { return mFilm ? zl_film_download_wait(mFilm) : ZL_ERR_INVALID_ARGUMENT; }

void Integrator::reportLaunchError(const char* what) {
    std::fprintf(stderr, "[Integrator] %s pass failed (%d): %s\n", what, mLastError, zl_last_error_string());
}

// the scene / camera uniforms every kernel receives (NaivePath.cpp:39-60)
ZlRenderParams Integrator::baseParams() const {
    ZlRenderParams p{};
    const Scene* scene = mStatus.scene;
    p.camera = scene->camera.uniforms();
    p.filmW = mStatus.renderSize[0];
    p.filmH = mStatus.renderSize[1];
    p.envRotation = scene->envRotation;
    p.spp = mCurSample;
    // renderOnePass() increments before it dispatches; with sample sharding the free counter follows
    // the pass index, as it would on one GPU
    p.freeCounter = (mShardStride > 1) ? mCurSample + 1 : mFreeCounter + 1;
    p.blocksOnePass = 0;
    p.loopsPerPass = 1;
    p.scale = 1.0f;
    return p;
}

// ---------------------------------------------------------------------------------------------
// NaivePathIntegrator (src/integrator/NaivePath.cpp)
// ---------------------------------------------------------------------------------------------
void NaivePathIntegrator::init(Scene* scene, int width, int height, PipelinePtr ctx) {
    mStream = ctx;
    mStatus.scene = scene;
    mStatus.renderSize[0] = width; mStatus.renderSize[1] = height;
    recreateFrameTex(width, height);
    mCurSample = mShardFirst;
}

ZlRenderParams NaivePathIntegrator::params(int) const {
    ZlRenderParams p = baseParams();
    p.maxDepth = mParam.maxDepth;
    p.russianRoulette = mParam.russianRoulette;
    p.sampleLight = mParam.sampleLight;
    p.lightEnvUniformSample = mParam.lightEnvUniformSample;
    p.lightPortion = mParam.lightPortion;
    p.sampler = mStatus.scene->sampler;   // the XML sampler, not mParam.sampler (App. B #23, NaivePath.cpp:47)
    return p;
}

void NaivePathIntegrator::renderOnePass() {
    if (mShouldReset) { reset(mStatus); mShouldReset = false; }
    if (mParam.finiteSample && mCurSample > mParam.maxSample) { mFreeCounter++; mRenderFinished = true; return; }   // the call still counts (NaivePath.cpp:79): uFreeCounter after a restart
    ZlRenderParams p = params();
    // a failed launch (workspace allocation, film mismatch, kernel error) must not count as a rendered pass
    if (!mDryRun && (mLastError = zl_launch_path_pass(mStatus.scene->glContext, mFilm, &p, mParam.kernelVariant, mStream)) != 0) return reportLaunchError("path");
    mFreeCounter++;
    mCurSample += mShardStride;
    mPasses++;
}

void NaivePathIntegrator::reset(const RenderStatus& status) {
    bool resized = status.renderSize[0] != mStatus.renderSize[0] || status.renderSize[1] != mStatus.renderSize[1];
    mStatus = status;
    if (resized || status.resetLevel == ResetLevel::FullReset || !mFilm) recreateFrameTex(status.renderSize[0], status.renderSize[1]);
    else zl_film_clear(mFilm, mStream);   // the GL kernel overwrites the frame when uSpp == 0 (path_integ_naive.glsl:170)
    mCurSample = mShardFirst;
    mPasses = 0;
}

// ---------------------------------------------------------------------------------------------
// LightPathIntegrator (src/integrator/LightPath.cpp)
// ---------------------------------------------------------------------------------------------
void LightPathIntegrator::init(Scene* scene, int width, int height, PipelinePtr ctx) {
    mStream = ctx;
    mStatus.scene = scene;
    mStatus.renderSize[0] = width; mStatus.renderSize[1] = height;
    recreateFrameTex(width, height);
    mCurSample = mShardFirst;
}

ZlRenderParams LightPathIntegrator::params(int) const {
    ZlRenderParams p = baseParams();
    p.maxDepth = mParam.maxDepth;
    p.russianRoulette = mParam.russianRoulette;
    p.sampler = 0;                        // LightPath.cpp:48-49
    p.blocksOnePass = mParam.threadBlocksOnePass;
    return p;
}

float LightPathIntegrator::trueScale() const {
    double paths = (double)mPasses * mParam.threadBlocksOnePass * ZL_LIGHT_GROUP_SIZE;
    double pixels = (double)mStatus.renderSize[0] * mStatus.renderSize[1];
    return paths > 0 ? (float)(pixels / paths) : 0.0f;
}

void LightPathIntegrator::renderOnePass() {
    int width = mStatus.renderSize[0], height = mStatus.renderSize[1];
    if (mShouldReset) { reset(mStatus); mShouldReset = false; }
    if (mParam.finiteSample && mParam.samplePerPixel > (float)mParam.maxSample) { mFreeCounter++; mRenderFinished = true; return; }   // LightPath.cpp:85
    ZlRenderParams p = params();
    if (!mDryRun && (mLastError = zl_launch_light_pass(mStatus.scene->glContext, mFilm, &p, mParam.kernelVariant, mStream)) != 0) return reportLaunchError("light");
    mFreeCounter++;
    // no img_copy pass: the film already is the rgba frame (float4 film + vector red)
    mParam.samplePerPixel += static_cast<float>(mParam.threadBlocksOnePass) * ZL_LIGHT_GROUP_SIZE / (width * height);
    mCurSample += mShardStride;
    mPasses++;
}

void LightPathIntegrator::reset(const RenderStatus& status) {
    int width = status.renderSize[0], height = status.renderSize[1];
    bool resized = width != mStatus.renderSize[0] || height != mStatus.renderSize[1];
    mStatus = status;
    if (resized || !mFilm) recreateFrameTex(width, height);
    else zl_film_clear(mFilm, mStream);
    mCurSample = mShardFirst;
    mPasses = 0;
    mParam.samplePerPixel = static_cast<float>(mParam.threadBlocksOnePass) * ZL_LIGHT_GROUP_SIZE / (width * height);   // sic: LightPath.cpp:133
}

// ---------------------------------------------------------------------------------------------
// TriplePathIntegrator (src/integrator/TriplePath.cpp)
// ---------------------------------------------------------------------------------------------
void TriplePathIntegrator::init(Scene* scene, int width, int height, PipelinePtr ctx) {
    mStream = ctx;
    mStatus.scene = scene;
    mStatus.renderSize[0] = width; mStatus.renderSize[1] = height;
    recreateFrameTex(width, height);
    mCurSample = mShardFirst;
}

ZlRenderParams TriplePathIntegrator::params(int kernel) const {
    ZlRenderParams p = baseParams();
    p.maxDepth = mParam.maxDepth;
    p.russianRoulette = mParam.russianRoulette;
    if (kernel == 0) {
        p.sampler = mStatus.scene->sampler;   // TriplePath.cpp:71
    } else {
        p.sampler = 0;                        // TriplePath.cpp:72
        p.blocksOnePass = mParam.LPTBlocksOnePass;
        p.loopsPerPass = mParam.LPTLoopsPerPass;
        p.scale = static_cast<float>(p.filmW * p.filmH) / (mParam.LPTBlocksOnePass * mParam.LPTLoopsPerPass * ZL_LIGHT_GROUP_SIZE);
    }
    return p;
}

void TriplePathIntegrator::renderOnePass() {
    if (mShouldReset) { reset(mStatus); mShouldReset = false; }
    if (mParam.finiteSample && mCurSample > mParam.maxSample) { mFreeCounter++; mRenderFinished = true; return; }   // TriplePath.cpp:97
    ZlRenderParams pt = params(0), lpt = params(1);
    // same stream => the LPT pass starts after the PT pass, like the GL memory barrier between them
    if (!mDryRun && (mLastError = zl_launch_triple_pt_pass(mStatus.scene->glContext, mFilm, &pt, mParam.kernelVariant, mStream)) != 0) return reportLaunchError("triple PT");
    if (!mDryRun && (mLastError = zl_launch_triple_lpt_pass(mStatus.scene->glContext, mFilm, &lpt, mParam.kernelVariant, mStream)) != 0) return reportLaunchError("triple LPT");
    mFreeCounter++;
    mParam.samplePerPixel += 1.0f;
    mCurSample += mShardStride;
    mPasses++;
}

void TriplePathIntegrator::reset(const RenderStatus& status) {
    bool resized = status.renderSize[0] != mStatus.renderSize[0] || status.renderSize[1] != mStatus.renderSize[1];
    mStatus = status;
    if (resized || !mFilm) recreateFrameTex(status.renderSize[0], status.renderSize[1]);
    else zl_film_clear(mFilm, mStream);
    mParam.samplePerPixel = 1.0f;   // sic: TriplePath.cpp:153
    mCurSample = mShardFirst;
    mPasses = 0;
}

}  // namespace zillum
