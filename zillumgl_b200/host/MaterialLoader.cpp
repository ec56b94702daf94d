// <material type="..."> of scene.xml -> the 64-byte Material record.  The schema is the reference's
// (src/core/MaterialLoader.cpp:28-66): which child tags each type reads, "default" meaning "keep the
// model's own materials", absent children leaving the struct defaults, unknown types yielding the
// default Lambertian record.  Here the schema is data: one row per (type, tag) with the record offset.
#include <cstddef>
#include <cstring>
#include <sstream>
#include "Material.h"

namespace zillum {
namespace {

struct Field { const char* tag; size_t offset; int arity; };     // arity floats read from the child's "value" attribute
struct Schema { const char* type; int id; Field fields[11]; };

#define ZL_F(name) {#name, offsetof(Material, name), 1}
constexpr Field kBaseColor = {"baseColor", offsetof(Material, baseColor), 3};
constexpr Field kEnd = {nullptr, 0, 0};
constexpr Schema kSchemas[] = {
    {"principled", Material::Principled, {kBaseColor, ZL_F(subsurface), ZL_F(metallic), ZL_F(roughness), ZL_F(specular), ZL_F(specularTint),
                                           ZL_F(sheen), ZL_F(sheenTint), ZL_F(clearcoat), ZL_F(clearcoatGloss), kEnd}},
    {"metalWorkflow", Material::MetalWorkflow, {kBaseColor, ZL_F(metallic), ZL_F(roughness), kEnd}},
    {"dielectric", Material::Dielectric, {kBaseColor, ZL_F(ior), ZL_F(roughness), kEnd}},
    {"thinDielectric", Material::ThinDielectric, {kBaseColor, ZL_F(ior), kEnd}},
    {"lambertian", Material::Lambertian, {kBaseColor, kEnd}},
};
#undef ZL_F

}  // namespace

std::optional<Material> loadMaterial(const XmlNode& node) {
    const std::string type = node.attribute("type");
    if (type == "default") return std::nullopt;
    Material material;
    for (const Schema& schema : kSchemas) {
        if (type != schema.type) continue;
        for (const Field* f = schema.fields; f->tag; f++) {
            const XmlNode child = node.child(f->tag);
            if (!child) continue;
            std::stringstream values(child.attribute("value"));
            float* dst = reinterpret_cast<float*>(reinterpret_cast<char*>(&material) + f->offset);
            for (int i = 0; i < f->arity; i++) values >> dst[i];      // same stream semantics as the reference's chained >> (MaterialLoader.cpp:16-25)
        }
        material.type = schema.id;
        break;
    }
    return material;
}

}  // namespace zillum
