// Integrator plugin interface + the three integrators on the hot path.  Mirrors the
// reference's src/core/Integrator.h: same abstract interface (init / renderOnePass / reset /
// getFrame / resultScale / recreateFrameTex / setStatus / setShouldReset; the ImGui hooks
// are no-ops), same parameter structs with the same defaults, same per-pass bookkeeping
// (mCurSample, mFreeCounter, samplePerPixel).  GL dispatch is replaced by the C-ABI launch
// shims of include/zillum_cuda.h; the frame is a host RGBA float image instead of a GL
// texture.  Additions for the 8xB200 box: setSampleShard() partitions passes by sample
// index, and the film may live in caller-provided device memory (for the NCCL all-reduce).
#pragma once
#include <memory>
#include <vector>
#include "Scene.h"

namespace zillum {

enum class ResetLevel { ResetFrame, ResetUniform, FullReset };

struct RenderStatus {
    Scene* scene = nullptr;
    int renderSize[2] = {0, 0};
    ResetLevel resetLevel = ResetLevel::ResetFrame;
};

using PipelinePtr = void*;   // cudaStream_t (NULL = default stream); the GL Pipeline object is gone

class Integrator {
public:
    virtual ~Integrator();
    virtual void init(Scene* scene, int width, int height, PipelinePtr ctx) = 0;
    virtual void renderOnePass() = 0;
    virtual void reset(const RenderStatus& status) = 0;
    virtual void renderSettingsGUI() {}
    virtual void renderProgressGUI() {}

    // accumulated radiance * resultScale(), RGBA float, row 0 = bottom; synchronises the stream
    virtual const std::vector<float>& getFrame();
    int downloadFrame(float scale, float* rgba);
    int snapshotAsync(void* dstDevice);                 // zl_film_snapshot_async on this integrator's stream (multi-GPU reduce-before-copy frames)        // film * scale into caller memory, on this integrator's stream
    // Pipelined frame read-back: enqueue "film * scale -> dstPinned" behind the passes rendered so far and
    // return at once; the copy overlaps the passes launched next.  waitFrame() blocks until dstPinned is
    // complete.  dstPinned: page-locked host memory, width*height*4 floats.  scale <= 0: trueScale().
    virtual int getFrameAsync(float* dstPinned, float scale = -1.0f, int channels = 4);     // channels 3: packed RGB (no constant alpha), width*height*3 floats
    virtual int waitFrame();
    // kernelVariant 2 keeps several passes in flight on internal streams; flush() makes this integrator's stream wait for them
    // (getFrame / getFrameAsync / postProcess / reset order themselves; only direct users of the film memory need it)
    virtual int flush();
    // Display stage of the reference's frame loop (post_proc.glsl dispatched after renderOnePass, Application.cpp:644-663):
    // film * scale (scale <= 0: trueScale()), tone mapped (0 none, 1 filmic = Config::toneMapping default, 2 ACES), gamma 1/2.2.
    // rgba: width*height*4 floats, rgb8: width*height*3 bytes (the screenshot read-back); either may be null.  Rows in film order.
    virtual int postProcess(float scale, int toneMapper, float* rgba, unsigned char* rgb8);
    virtual float resultScale() const = 0;
    // sum / true sample count (App. B #20 documents the reference's off-by-one resultScale)
    virtual float trueScale() const = 0;
    virtual void recreateFrameTex(int width, int height);

    void setStatus(const RenderStatus& status) { mStatus = status; }
    void setShouldReset() { mShouldReset = true; }

    // Sample-index partitioning: this instance renders passes
    // first, first+stride, ... with the uSpp/uFreeCounter values a single GPU would use.
    void setSampleShard(int first, int stride) { mShardFirst = first; mShardStride = stride; mCurSample = first; }
    // Place the film in caller-owned device memory (W*H*4 floats); call before init().
    void setExternalFilm(void* devicePtr) { mExternalFilm = devicePtr; }
    ZlFilm* film() const { return mFilm; }
    int curSample() const { return mCurSample; }
    bool renderFinished() const { return mRenderFinished; }
    // uniforms of the NEXT pass (what renderOnePass() will launch with)
    virtual ZlRenderParams params(int kernel = 0) const = 0;
    unsigned long long passesRendered() const { return mPasses; }
    // 0, or the error code of the last failed zl_launch_*_pass: renderOnePass() keeps the reference's void signature, a failed
    // launch leaves mCurSample / mPasses / samplePerPixel where they were and is reported here (and on stderr)
    int lastError() const { return mLastError; }
    // dry run: renderOnePass() does the bookkeeping (pass index, free counter, samplePerPixel) and launches nothing — for CPU
    // tests of the host logic (sample shards over gloo); nothing is rendered, so nothing here is a CPU fallback
    void setDryRun(bool on) { mDryRun = on; }

protected:
    ZlRenderParams baseParams() const;
    bool mRenderFinished = false;
    bool mPassFinished = false;
    int mCurSample = 0;
    int mFreeCounter = 0;
    RenderStatus mStatus;
    bool mShouldReset = false;
    double mTime = 0.0;

    int mShardFirst = 0, mShardStride = 1;
    unsigned long long mPasses = 0;
    int mLastError = 0;
    bool mDryRun = false;
    void reportLaunchError(const char* what);
    ZlFilm* mFilm = nullptr;
    void* mExternalFilm = nullptr;
    PipelinePtr mStream = nullptr;
    std::vector<float> mFrame;              // host copy of the frame, sized by the first getFrame()
    int mFrameW = 0, mFrameH = 0;
};

using IntegratorPtr = std::shared_ptr<Integrator>;

struct PathIntegParam {
    int maxDepth = 4;
    bool russianRoulette = false;
    bool sampleLight = true;
    bool lightEnvUniformSample = false;
    float lightPortion = 0.5f;
    bool finiteSample = false;
    int maxSample = 64;
    int sampler = 1;
    int kernelVariant = 1;   // 0 = megakernel, 1 = wavefront / ray regeneration (B200 addition; bit-identical film, 3.2x faster), 2 = wavefront with three (small films: four) passes in flight (bit-identical film, +17 % at 4K, +100 % at 720p)
};

class NaivePathIntegrator : public Integrator {
public:
    void init(Scene* scene, int width, int height, PipelinePtr ctx) override;
    void renderOnePass() override;
    void reset(const RenderStatus& status) override;
    float resultScale() const override { return 1.0f / (mCurSample + 1); }
    float trueScale() const override { return mPasses ? 1.0f / (float)mPasses : 0.0f; }
    ZlRenderParams params(int kernel = 0) const override;
    PathIntegParam mParam;
};

struct LightPathIntegParam {
    int maxDepth = 4;
    bool russianRoulette = false;
    bool finiteSample = false;
    int maxSample = 64;
    int threadBlocksOnePass = 32;
    float samplePerPixel = 0.0f;
    int kernelVariant = 1;   // 0 = megakernel, 1 = wavefront (B200 addition)
};

class LightPathIntegrator : public Integrator {
public:
    void init(Scene* scene, int width, int height, PipelinePtr ctx) override;
    void renderOnePass() override;
    void reset(const RenderStatus& status) override;
    float resultScale() const override { return 1.0f / mParam.samplePerPixel; }
    float trueScale() const override;
    ZlRenderParams params(int kernel = 0) const override;
    LightPathIntegParam mParam;
};

struct TriplePathIntegParam {
    int maxDepth = 4;
    bool russianRoulette = false;
    int LPTBlocksOnePass = 64;
    int LPTLoopsPerPass = 1;
    bool finiteSample = false;
    int maxSample = 64;
    float samplePerPixel = 0.0f;
    int PTSampler = 1;
    bool limitTime = true;
    double maxTime = 30.0;
    int kernelVariant = 1;   // 0 = megakernel, 1 = wavefront (B200 addition)
};

class TriplePathIntegrator : public Integrator {
public:
    void init(Scene* scene, int width, int height, PipelinePtr ctx) override;
    void renderOnePass() override;
    void reset(const RenderStatus& status) override;
    float resultScale() const override { return 1.0f / mParam.samplePerPixel; }
    float trueScale() const override { return mPasses ? 1.0f / (float)mPasses : 0.0f; }
    // kernel 0 = PT pass uniforms, 1 = LPT pass uniforms
    ZlRenderParams params(int kernel = 0) const override;
    TriplePathIntegParam mParam;
};

}  // namespace zillum
