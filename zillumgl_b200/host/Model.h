// Mesh / model-instance containers and the model pool (reference: src/core/Mesh.h,
// Model.{h,cpp}, Resource.{h,cpp}).  Assimp is not available, so models come from a small
// Wavefront OBJ/MTL reader or from "builtin:" procedural meshes; the containers, the TRS
// model matrix with its constant +90 degree X rotation (Model.cpp:6,62-72) and the
// matIndex/texIndex bookkeeping (Resource.cpp:124-168) are the reference's.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "Material.h"
#include "Math.h"

namespace zillum {

struct MeshData {
    std::vector<Vec3f> positions;
    std::vector<Vec2f> texcoords;
    std::vector<Vec3f> normals;
    std::vector<uint32_t> indices;
    // model space is Y-up; helper for generators that think in the Z-up world frame
    static Vec3f fromZUp(float x, float y, float z) { return Vec3f(x, z, -y); }
    void addTriangle(uint32_t a, uint32_t b, uint32_t c) { indices.push_back(a); indices.push_back(b); indices.push_back(c); }
    uint32_t addVertex(Vec3f p, Vec3f n, Vec2f t) { positions.push_back(p); normals.push_back(n); texcoords.push_back(t); return (uint32_t)positions.size() - 1; }
};
using MeshDataPtr = std::shared_ptr<MeshData>;

struct MeshInstance {
    int texIndex = -1;
    int matIndex = 0;
    int globalMatIndex = -1;
    MeshDataPtr meshData;
};
using MeshInstancePtr = std::shared_ptr<MeshInstance>;

class ModelInstance;
using ModelInstancePtr = std::shared_ptr<ModelInstance>;

class ModelInstance {
public:
    void setPos(Vec3f p) { mPos = p; }
    void setScale(Vec3f s) { mScale = s; }
    void setScale(float x, float y, float z) { mScale = Vec3f(x, y, z); }
    void setRotation(Vec3f angle) { mRotation = angle; }
    void setName(const std::string& n) { mName = n; }
    void setPath(const std::string& p) { mPath = p; }
    Vec3f pos() const { return mPos; }
    Vec3f scale() const { return mScale; }
    Vec3f rotation() const { return mRotation; }
    const std::string& name() const { return mName; }
    const std::string& path() const { return mPath; }
    Affine modelMatrix() const;                        // Model.cpp:62-72
    std::vector<MeshInstancePtr>& meshInstances() { return mMeshInstances; }
    std::vector<Material>& materials() { return mMaterials; }
    ModelInstancePtr copy() const;                     // shares MeshData, copies instances + materials

private:
    std::vector<MeshInstancePtr> mMeshInstances;
    std::vector<Material> mMaterials;
    Vec3f mPos{0, 0, 0}, mScale{1, 1, 1}, mRotation{0, 0, 0};
    std::string mName, mPath;
};

struct ByteImage { std::vector<unsigned char> rgb; int width = 0, height = 0; };
using ByteImagePtr = std::shared_ptr<ByteImage>;

namespace Resource {
// "path" is a Wavefront .obj file or "builtin:<mesh>[?k=v&...]" (square, cube, sphere,
// teapotBody, teapotCap, sponza, rungholt, ...).  Models are pooled by path.
ModelInstancePtr openModelInstance(const std::string& path, Vec3f pos = Vec3f(0.0f), Vec3f scale = Vec3f(1.0f), Vec3f rotation = Vec3f(0.0f));
int addImage(const std::string& path);                 // -1 on failure; pooled by path
int addImage(ByteImagePtr img, const std::string& key);
const std::vector<ByteImagePtr>& getAllImages();
void clear();
}  // namespace Resource

// procedural meshes (ProceduralMeshes.cpp)
ModelInstancePtr makeBuiltinModel(const std::string& spec);

}  // namespace zillum
