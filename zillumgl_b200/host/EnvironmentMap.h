// Environment map + two-level alias tables (reference: src/core/EnvironmentMap.{h,cpp}).
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace zillum {

class EnvironmentMap;
using EnvironmentMapPtr = std::shared_ptr<EnvironmentMap>;

class EnvironmentMap {
public:
    // rgb: width*height*3 floats, row 0 = the +Z pole (v = 0), as stb_image delivers a lat-long HDR.
    EnvironmentMap(std::vector<float> rgb, int width, int height);
    int width() const { return mWidth; }
    int height() const { return mHeight; }
    int sumPdf() const { return (int)mSumPdf; }   // int on purpose: EnvironmentMap.h:22
    const std::vector<float>& pixels() const { return mPixels; }
    const std::vector<int32_t>& aliasTable() const { return mAlias; }   // (W+1) x H, column W = row marginal
    const std::vector<float>& aliasProb() const { return mProb; }

    static EnvironmentMapPtr create(const std::string& path);           // PFM / Radiance .hdr / "builtin:sky"
    static EnvironmentMapPtr createBlack();                              // stands in for a missing map
    static EnvironmentMapPtr createProceduralSky(int width, int height, float sunElevationDeg, float sunAzimuthDeg);

private:
    std::vector<float> mPixels;
    std::vector<int32_t> mAlias;
    std::vector<float> mProb;
    int mWidth, mHeight;
    float mSumPdf = 0.0f;
};

}  // namespace zillum
