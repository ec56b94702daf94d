// 64-byte material record = 4 RGBA32F texels, the layout contract with the device loader
// (reference: src/core/Material.h:32-53, material_loader.glsl:3-97).
#pragma once
#include <optional>
#include "Math.h"
#include "Xml.h"

namespace zillum {

struct Material {
    enum { Lambertian = 0, Principled, MetalWorkflow, Dielectric, ThinDielectric };
    Vec3f baseColor = Vec3f(1.0f);
    float roughness = 1.0f;
    float subsurface = 0.0f, metallic = 0.0f, specular = 1.0f, specularTint = 1.0f;
    float sheen = 0.0f, sheenTint = 1.0f, clearcoat = 0.0f, clearcoatGloss = 0.0f;
    float ior = 1.5f;
    int type = Lambertian;
    float padding2 = 0.0f, padding3 = 0.0f;
};
static_assert(sizeof(Material) == 64, "Material must be 4 RGBA32F texels");

// <material type="..."> -> Material; "default" -> nullopt (MaterialLoader.cpp:5-69)
std::optional<Material> loadMaterial(const XmlNode& node);

}  // namespace zillum
