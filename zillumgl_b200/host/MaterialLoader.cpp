#include <sstream>
#include "Material.h"

namespace zillum {

std::optional<Material> loadMaterial(const XmlNode& node) {
    Material material;
    const std::string type = node.attribute("type");
    // a missing child leaves the default in place (MaterialLoader.cpp:10-26)
    auto loadFloat = [&](const char* name, float& value) {
        XmlNode c = node.child(name);
        if (!c) return;
        std::stringstream ss(c.attribute("value"));
        ss >> value;
    };
    auto loadVec3f = [&](const char* name, Vec3f& value) {
        XmlNode c = node.child(name);
        if (!c) return;
        std::stringstream ss(c.attribute("value"));
        ss >> value.x >> value.y >> value.z;
    };
    if (type == "default") return std::nullopt;
    if (type == "principled") {
        loadVec3f("baseColor", material.baseColor);
        loadFloat("subsurface", material.subsurface);
        loadFloat("metallic", material.metallic);
        loadFloat("roughness", material.roughness);
        loadFloat("specular", material.specular);
        loadFloat("specularTint", material.specularTint);
        loadFloat("sheen", material.sheen);
        loadFloat("sheenTint", material.sheenTint);
        loadFloat("clearcoat", material.clearcoat);
        loadFloat("clearcoatGloss", material.clearcoatGloss);
        material.type = Material::Principled;
    } else if (type == "metalWorkflow") {
        loadVec3f("baseColor", material.baseColor);
        loadFloat("metallic", material.metallic);
        loadFloat("roughness", material.roughness);
        material.type = Material::MetalWorkflow;
    } else if (type == "dielectric") {
        loadVec3f("baseColor", material.baseColor);
        loadFloat("ior", material.ior);
        loadFloat("roughness", material.roughness);
        material.type = Material::Dielectric;
    } else if (type == "thinDielectric") {
        loadVec3f("baseColor", material.baseColor);
        loadFloat("ior", material.ior);
        material.type = Material::ThinDielectric;
    } else if (type == "lambertian") {
        loadVec3f("baseColor", material.baseColor);
        material.type = Material::Lambertian;
    }
    return material;
}

}  // namespace zillum
