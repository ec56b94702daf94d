// zillum_render — headless driver: scene.xml (or a built-in scene) -> linear-radiance EXR / PFM (+ tone-mapped PNG).
// Replaces the GLFW/ImGui main loop of the reference (src/main.cpp:5-9, Application::run
// src/Application.cpp:644-687): load the scene, create the integrator the XML names, call
// renderOnePass() spp times, download the frame, write it.  No window; --png adds the reference's
// screenshot (post_proc.glsl tone mapping + gamma, captureImage() Application.cpp:371-380).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include "ImageIO.h"
#include "Integrator.h"

using namespace zillum;

static void usage() {
    std::fprintf(stderr,
                 "usage: zillum_render <scene.xml | builtin:NAME> [--integrator path|light|triple] [--spp N]\n"
                 "                     [--size WxH] [--depth N] [--rr] [--variant 0|1|2] [--device N] [--out image.exr|image.pfm]\n"
                 "                     [--png image.png] [--tonemap none|filmic|aces]\n"
                 "built-in scenes: default cornell sponza sponza_light rungholt rungholt_small\n");
}

int main(int argc, char** argv) {
    if (argc < 2) { usage(); return 2; }
    std::string scenePath = argv[1], integ, out = "render.pfm", png, tonemap = "filmic";
    int spp = 64, width = 0, height = 0, depth = -1, variant = 2, device = 0;
    bool rr = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
        if (a == "--integrator") integ = next();
        else if (a == "--spp") spp = std::atoi(next());
        else if (a == "--size") { if (std::sscanf(next(), "%dx%d", &width, &height) != 2) { usage(); return 2; } }
        else if (a == "--depth") depth = std::atoi(next());
        else if (a == "--rr") rr = true;
        else if (a == "--variant") variant = std::atoi(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--png") png = next();
        else if (a == "--tonemap") tonemap = next();
        else { usage(); return 2; }
    }
    if (zl_set_device(device) != 0) { std::fprintf(stderr, "zillum_render: %s\n", zl_last_error_string()); return 1; }

    Scene scene;
    bool ok;
    if (scenePath.rfind("builtin:", 0) == 0) ok = scene.loadBuiltin(scenePath.substr(8), width > 0 ? width : 1280, height > 0 ? height : 720);
    else ok = scene.load(scenePath);
    if (!ok) { std::fprintf(stderr, "zillum_render: cannot load scene '%s'\n", scenePath.c_str()); return 1; }
    if (width <= 0 || height <= 0) { width = scene.filmWidth; height = scene.filmHeight; }
    scene.filmWidth = width; scene.filmHeight = height;          // the noise / seed image follows the film size (Scene.cpp:263)
    scene.buildBvhOnDevice = true;        // BVH::build + the six MTBVH orderings on the device (zl_scene_create)
    scene.createGLContext(true);
    if (!scene.glContext) { std::fprintf(stderr, "zillum_render: scene upload failed: %s\n", zl_last_error_string()); return 1; }

    if (integ.empty()) integ = scene.integratorType;
    std::unique_ptr<Integrator> integrator;
    if (integ == "light" || integ == "lightPath") {
        auto* p = new LightPathIntegrator();
        if (depth >= 0) p->mParam.maxDepth = depth;
        p->mParam.russianRoulette = rr;
        p->mParam.threadBlocksOnePass = (width * height + ZL_LIGHT_GROUP_SIZE - 1) / ZL_LIGHT_GROUP_SIZE;   // 1 spp-equivalent per pass
        p->mParam.kernelVariant = variant;
        integrator.reset(p);
    } else if (integ == "triple" || integ == "triplePath") {
        auto* p = new TriplePathIntegrator();
        if (depth >= 0) p->mParam.maxDepth = depth;
        p->mParam.russianRoulette = rr;
        p->mParam.kernelVariant = variant;
        integrator.reset(p);
    } else {
        auto* p = new NaivePathIntegrator();
        if (depth >= 0) p->mParam.maxDepth = depth;
        p->mParam.russianRoulette = rr;
        p->mParam.kernelVariant = variant;
        integrator.reset(p);
    }
    integrator->init(&scene, width, height, nullptr);
    RenderStatus status;
    status.scene = &scene; status.renderSize[0] = width; status.renderSize[1] = height;
    integrator->setStatus(status);

    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < spp; i++) integrator->renderOnePass();
    integrator->flush();
    zl_device_synchronize();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::vector<float> frame((size_t)width * height * 4);
    if (zl_film_download(integrator->film(), integrator->trueScale(), frame.data(), nullptr) != 0) {
        std::fprintf(stderr, "zillum_render: %s\n", zl_last_error_string());
        return 1;
    }
    bool exr = out.size() > 4 && out.compare(out.size() - 4, 4, ".exr") == 0;
    ok = exr ? writeEXR(out, frame.data(), width, height) : writePFM(out, frame.data(), width, height);
    if (!ok) { std::fprintf(stderr, "zillum_render: cannot write '%s'\n", out.c_str()); return 1; }
    if (!png.empty()) {
        const int tm = tonemap == "none" ? 0 : tonemap == "aces" ? 2 : 1;                  // Config::toneMapping = 1 (Application.cpp:98)
        std::vector<unsigned char> rgb8((size_t)width * height * 3);
        if (integrator->postProcess(-1.0f, tm, nullptr, rgb8.data()) != 0) { std::fprintf(stderr, "zillum_render: %s\n", zl_last_error_string()); return 1; }
        if (!writePNG(png, rgb8.data(), width, height)) { std::fprintf(stderr, "zillum_render: cannot write '%s'\n", png.c_str()); return 1; }
    }
    double mean = 0.0;
    size_t nan = 0, inf = 0;          // the reference filters NaN results, not infinite ones (light_path_integ.glsl:118): a splat can be inf
    for (size_t i = 0; i < (size_t)width * height; i++)
        for (int c = 0; c < 3; c++) { const float v = frame[4 * i + c]; if (v != v) nan++; else if (v - v != 0.0f) inf++; else mean += v; }
    mean /= 3.0 * width * height;
    double bvhMs = 0.0, threadMs = 0.0; int levels = 0;
    zl_scene_prep_times(scene.glContext, &bvhMs, &threadMs, &levels);
    std::printf("{\"scene\": \"%s\", \"integrator\": \"%s\", \"variant\": %d, \"width\": %d, \"height\": %d, \"passes\": %d, \"seconds\": %.4f, "
                "\"ms_per_pass\": %.4f, \"mean_radiance\": %.6f, \"nan_components\": %zu, \"inf_components\": %zu, \"triangles\": %d, \"device_bvh_build_ms\": %.2f, "
                "\"device_mtbvh_thread_ms\": %.2f, \"out\": \"%s\"}\n",
                scenePath.c_str(), integ.c_str(), variant, width, height, spp, sec, sec * 1e3 / (spp > 0 ? spp : 1), mean, nan, inf,
                (int)(scene.host.indices.size() / 3), bvhMs, threadMs, out.c_str());
    return 0;
}
