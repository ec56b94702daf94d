#include "EnvironmentMap.h"
#include <cmath>
#include "AliasTable.h"
#include "ImageIO.h"

namespace zillum {

EnvironmentMap::EnvironmentMap(std::vector<float> rgb, int width, int height)
    : mPixels(std::move(rgb)), mWidth(width), mHeight(height) {
    const size_t rowStride = (size_t)width + 1;
    mAlias.assign(rowStride * height, 0);
    mProb.assign(rowStride * height, 0.0f);
    // weight = Rec.709 luminance * sin(theta) (EnvironmentMap.cpp:24-42)
    for (int i = 0; i < height; i++) {
        float sinTheta = std::sin((float)(i + 0.5f) / height * 3.141592653589793f);
        for (int j = 0; j < width; j++) {
            const float* p = &mPixels[3 * ((size_t)i * width + j)];
            mProb[i * rowStride + j] = (0.2126f * p[0] + 0.7152f * p[1] + 0.0722f * p[2]) * sinTheta;
        }
    }
    // per-row conditional tables; their sums become the marginal weights in column `width`
    for (int i = 0; i < height; i++)
        mProb[i * rowStride + width] =
            AliasTable::buildStrided<int32_t>(&mAlias[i * rowStride], &mProb[i * rowStride], width, 1, true);
    mSumPdf = AliasTable::buildStrided<int32_t>(&mAlias[width], &mProb[width], height, (int)rowStride, true);
}

EnvironmentMapPtr EnvironmentMap::createBlack() {
    return std::make_shared<EnvironmentMap>(std::vector<float>(3, 0.0f), 1, 1);
}

// Analytic sky + sun used as the synthetic HDR environment for the Sponza-class configs
// (the reference ships no assets: res/ holds only scene.xml).
EnvironmentMapPtr EnvironmentMap::createProceduralSky(int width, int height, float sunElevationDeg, float sunAzimuthDeg) {
    std::vector<float> rgb((size_t)width * height * 3);
    const float pi = 3.14159265358979f;
    float se = sunElevationDeg * pi / 180.0f, sa = sunAzimuthDeg * pi / 180.0f;
    float sun[3] = {std::cos(sa) * std::cos(se), std::sin(sa) * std::cos(se), std::sin(se)};
    for (int i = 0; i < height; i++) {
        float phi = (i + 0.5f) / height * pi;                 // polar angle from +Z
        for (int j = 0; j < width; j++) {
            float theta = (j + 0.5f) / width * 2.0f * pi;
            float d[3] = {std::cos(theta) * std::sin(phi), std::sin(theta) * std::sin(phi), std::cos(phi)};
            float up = d[2];
            float cs = d[0] * sun[0] + d[1] * sun[1] + d[2] * sun[2];
            float* p = &rgb[3 * ((size_t)i * width + j)];
            if (up >= 0.0f) {
                float horizon = std::pow(1.0f - up, 3.0f);
                p[0] = 0.35f + 0.55f * horizon; p[1] = 0.55f + 0.40f * horizon; p[2] = 1.00f + 0.10f * horizon;
            } else {
                float g = 0.18f * std::exp(4.0f * up);
                p[0] = 0.30f * g + 0.05f; p[1] = 0.27f * g + 0.05f; p[2] = 0.22f * g + 0.05f;
            }
            float glow = std::pow(std::fmax(cs, 0.0f), 64.0f) * 4.0f;
            float disk = cs > 0.9994f ? 4000.0f : 0.0f;        // ~2 degree sun disk
            p[0] += glow * 1.0f + disk * 1.0f; p[1] += glow * 0.85f + disk * 0.93f; p[2] += glow * 0.6f + disk * 0.8f;
        }
    }
    return std::make_shared<EnvironmentMap>(std::move(rgb), width, height);
}

EnvironmentMapPtr EnvironmentMap::create(const std::string& path) {
    if (path.empty()) return createBlack();
    if (path.rfind("builtin:sky", 0) == 0) return createProceduralSky(2048, 1024, 40.0f, 60.0f);
    int w = 0, h = 0;
    std::vector<float> rgb;
    if (!loadFloatImage(path, rgb, w, h)) return createBlack();
    return std::make_shared<EnvironmentMap>(std::move(rgb), w, h);
}

}  // namespace zillum
