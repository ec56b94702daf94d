// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref/libzillum_ref.so.
//
// ref_host.cpp — runs the reference's OWN host code.  The build compiles, unmodified and where
// they lie, src/accelerator/{AABB,BVH}.cpp, src/core/{Buffer,Texture,Image,Camera,Sampler,
// EnvironmentMap,Model,Scene,MaterialLoader}.cpp and src/integrator/{NaivePath,LightPath,
// TriplePath}.cpp against the stand-in headers of ref_shim/include (glm subset, GL entry points
// over host memory = ref_gl.cpp, pugixml / stb_image / ImGui stand-ins).  This file supplies
//   1. the members of the reference's classes whose own definitions cannot be built here:
//      Shader (src/core/Shader.cpp needs a GLSL compiler: the programs are the pre-translated
//      ones of ref_glsl2cpp.py, looked up by file name), the three Pipeline statics the
//      integrators call (Pipeline.cpp:72-115), and Resource (Resource.cpp needs Assimp: "model
//      files" are in-memory meshes registered by the test, pooled and copied like Resource.cpp:94-117);
//   2. an extern "C" surface over the reference's classes for tests/ref_lib.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "accelerator/BVH.h"
#include "math/AliasTable.h"
#include "core/Camera.h"
#include "core/Sampler.h"
#include "core/EnvironmentMap.h"
#include "core/Scene.h"
#include "core/Integrator.h"

#include "glsl_shim.h"
#include "ref_gl.h"
#include "../include/zillum_cuda.h"

// ================================================================================================
// 1a. Shader: uniforms by name on a pre-translated program (replaces src/core/Shader.cpp)
// ================================================================================================
namespace {

struct ShaderState {
    glsl::Program* program = nullptr;
    std::map<std::string, GLuint> samplers;      // sampler uniform -> texture object name (glBindTextureUnit + glUniform1i)
};
std::map<const Shader*, ShaderState>& shaderStates() { static std::map<const Shader*, ShaderState> m; return m; }
std::map<uint32_t, GLuint>& imageUnits() { static std::map<uint32_t, GLuint> m; return m; }   // glBindImageTexture

glsl::Program* findProgram(const std::string& name) {
    for (glsl::Program* p = glsl::programList(); p; p = p->next)
        if (name == p->name) return p;
    return nullptr;
}
const glsl::UniformEntry* findUniform(glsl::Program* p, const std::string& name) {
    for (int i = 0; i < p->numUniforms; i++)
        if (name == p->uniforms[i].name) return &p->uniforms[i];
    return nullptr;
}
const float* srgbLut() {
    static float lut[256]; static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; i++) {          // GL_SRGB texel decode (OpenGL 4.5 §8.24), evaluated in binary64
            double c = i / 255.0;
            lut[i] = (float)((c <= 0.04045) ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        }
        init = true;
    }
    return lut;
}

}  // namespace

Shader::Shader(const File::path& path, const glm::ivec3& computeSize, const std::string& extensionStr) :
    GLStateObject(GLStateObjectType::Shader), mName(path.generic_string()), mExtensionStr(extensionStr), mComputeGroupSize(computeSize) {
    glsl::Program* p = findProgram(mName);
    if (!p) { std::fprintf(stderr, "[zillum_ref] shader %s was not pre-translated\n", mName.c_str()); std::abort(); }
    if (p->localSize[0] != computeSize.x || p->localSize[1] != computeSize.y || p->localSize[2] != computeSize.z) {
        std::fprintf(stderr, "[zillum_ref] %s: work-group size %dx%dx%d differs from the translated %dx%dx%d\n", mName.c_str(),
                     computeSize.x, computeSize.y, computeSize.z, p->localSize[0], p->localSize[1], p->localSize[2]);
        std::abort();
    }
    shaderStates()[this].program = p;
    static uint32_t nextId = 1;
    mId = nextId++;
}
Shader::Shader(const ShaderSource&, const std::string&) : GLStateObject(GLStateObjectType::Shader) { std::abort(); }
Shader::Shader(const File::path&) : GLStateObject(GLStateObjectType::Shader) { std::abort(); }
Shader::~Shader() { shaderStates().erase(this); }
void Shader::enable() {}
void Shader::disable() {}
ShaderPtr Shader::createFromText(const File::path& path, const glm::ivec3& computeSize, const std::string& extensionStr) {
    return std::make_shared<Shader>(path, computeSize, extensionStr);
}
int Shader::getUniformLocation(const std::string& name) {
    glsl::Program* p = shaderStates()[this].program;
    for (int i = 0; i < p->numUniforms; i++) if (name == p->uniforms[i].name) return i;
    return -1;                                      // like GL: setting an absent uniform is ignored
}
void Shader::set1i(const std::string& name, int v) {
    const glsl::UniformEntry* u = findUniform(shaderStates()[this].program, name); if (!u) return;
    switch (u->type) {
    case glsl::U_INT: *(int*)u->ptr = v; break;
    case glsl::U_UINT: *(glsl::uint*)u->ptr = (glsl::uint)v; break;
    case glsl::U_BOOL: *(bool*)u->ptr = v != 0; break;
    default: std::fprintf(stderr, "[zillum_ref] set1i on %s\n", name.c_str()); std::abort();
    }
}
void Shader::set1f(const std::string& name, float v) {
    const glsl::UniformEntry* u = findUniform(shaderStates()[this].program, name); if (!u) return;
    if (u->type != glsl::U_FLOAT) { std::fprintf(stderr, "[zillum_ref] set1f on %s\n", name.c_str()); std::abort(); }
    *(float*)u->ptr = v;
}
void Shader::set2i(const std::string& name, int a, int b) { setVec2i(name, glm::ivec2(a, b)); }
void Shader::set2f(const std::string& name, float a, float b) { setVec2(name, glm::vec2(a, b)); }
void Shader::set3f(const std::string& name, float a, float b, float c) { setVec3(name, glm::vec3(a, b, c)); }
void Shader::set4f(const std::string& name, float a, float b, float c, float d) { setVec4(name, glm::vec4(a, b, c, d)); }
void Shader::setVec2(const std::string& name, const glm::vec2& v) {
    const glsl::UniformEntry* u = findUniform(shaderStates()[this].program, name); if (!u) return;
    if (u->type != glsl::U_VEC2) std::abort();
    *(glsl::vec2*)u->ptr = glsl::vec2(v.x, v.y);
}
void Shader::setVec2i(const std::string& name, const glm::ivec2& v) {
    const glsl::UniformEntry* u = findUniform(shaderStates()[this].program, name); if (!u) return;
    if (u->type != glsl::U_IVEC2) std::abort();
    *(glsl::ivec2*)u->ptr = glsl::ivec2(v.x, v.y);
}
void Shader::setVec3(const std::string& name, const glm::vec3& v) {
    const glsl::UniformEntry* u = findUniform(shaderStates()[this].program, name); if (!u) return;
    if (u->type != glsl::U_VEC3) std::abort();
    *(glsl::vec3*)u->ptr = glsl::vec3(v.x, v.y, v.z);
}
void Shader::setVec4(const std::string& name, const glm::vec4& v) {
    const glsl::UniformEntry* u = findUniform(shaderStates()[this].program, name); if (!u) return;
    if (u->type != glsl::U_VEC4) std::abort();
    *(glsl::vec4*)u->ptr = glsl::vec4(v.x, v.y, v.z, v.w);
}
void Shader::setMat3(const std::string& name, const glm::mat3& m) {
    const glsl::UniformEntry* u = findUniform(shaderStates()[this].program, name); if (!u) return;
    if (u->type != glsl::U_MAT3) std::abort();
    *(glsl::mat3*)u->ptr = glsl::mat3(glsl::vec3(m[0].x, m[0].y, m[0].z), glsl::vec3(m[1].x, m[1].y, m[1].z), glsl::vec3(m[2].x, m[2].y, m[2].z));
}
void Shader::setMat4(const std::string&, const glm::mat4&) { std::abort(); }
void Shader::setTexture(const std::string& name, TexturePtr tex, uint32_t /*slot*/) {
    if (!tex) return;
    shaderStates()[this].samplers[name] = tex->id();
}

// ================================================================================================
// 1b. Pipeline: the compute entry points (replaces src/core/Pipeline.cpp:72-115)
// ================================================================================================
std::map<TexturePtr, Pipeline::TextureBindParam> Pipeline::mImageBindRec;

void Pipeline::bindTextureToImage(TexturePtr texture, uint32_t unit, int, ImageAccess, TextureFormat) { imageUnits()[unit] = texture->id(); }
void Pipeline::clearBindingRecord() { mImageBindRec.clear(); }
void Pipeline::memoryBarrier(MemoryBarrierBit) {}

static void resolveBindings(const Shader* shader, glsl::Program* p) {
    ShaderState& st = shaderStates()[shader];
    for (int i = 0; i < p->numUniforms; i++) {
        const glsl::UniformEntry& u = p->uniforms[i];
        if (u.type >= glsl::U_SAMPLER_BUFFER && u.type <= glsl::U_SAMPLER_2D_ARRAY) {
            glsl::TexBinding b;
            auto it = st.samplers.find(u.name);
            refgl::Object* o = it == st.samplers.end() ? nullptr : refgl::object(it->second);
            if (o) {
                size_t bytes = 0;
                b.data = refgl::texels(*o, &bytes);
                b.comps = o->comps; b.w = o->width; b.h = o->height; b.layers = o->layers; b.kind = o->kind;
                if (o->kind == refgl::K_SRGB8) b.lut = srgbLut();
                // a view of a buffer through another sampler class keeps the element count (uMatTypes over uMaterials)
            }
            *(glsl::TexBinding*)u.ptr = b;
        } else if (u.type == glsl::U_IMAGE_2D) {
            glsl::image2D im;
            auto it = imageUnits().find((uint32_t)u.binding);
            refgl::Object* o = it == imageUnits().end() ? nullptr : refgl::object(it->second);
            if (o) { im.data = (float*)o->bytes.data(); im.w = o->width; im.h = o->height; im.comps = o->comps; }
            *(glsl::image2D*)u.ptr = im;
        }
    }
}

void Pipeline::dispatchCompute(int xNum, int yNum, int zNum, ShaderPtr shader) {
    glsl::Program* p = shaderStates()[shader.get()].program;
    resolveBindings(shader.get(), p);
    const long nx = (long)xNum * p->localSize[0], ny = (long)yNum * p->localSize[1], nz = (long)zNum * p->localSize[2];
    if (ny == 1 && nz == 1) {
#pragma omp parallel for schedule(dynamic, 256)
        for (long x = 0; x < nx; x++) p->invoke((glsl::uint)x, 0u, 0u);
    } else {
#pragma omp parallel for schedule(dynamic, 1)
        for (long row = 0; row < ny * nz; row++)
            for (long x = 0; x < nx; x++) p->invoke((glsl::uint)x, (glsl::uint)(row % ny), (glsl::uint)(row / ny));
    }
}

// ================================================================================================
// 1c. Resource: pooled "model files" and images held in memory (replaces src/core/Resource.cpp)
// ================================================================================================
std::vector<ImagePtr> Resource::imagePool;
std::map<File::path, int> Resource::mapPathToImageIndex;
std::vector<MeshDataPtr> Resource::meshDataPool;
std::map<File::path, ModelInstancePtr> Resource::mapPathToModelInstance;

ImagePtr Resource::getImageByIndex(int index) { return (index >= 0 && index < (int)imagePool.size()) ? imagePool[index] : nullptr; }
ImagePtr Resource::getImageByPath(const File::path& path) {
    auto it = mapPathToImageIndex.find(path);
    return it == mapPathToImageIndex.end() ? nullptr : getImageByIndex(it->second);
}
int Resource::addImage(const File::path& path, ImageDataType type) {      // Resource.cpp:24-35
    auto it = mapPathToImageIndex.find(path);
    if (it != mapPathToImageIndex.end()) return it->second;
    ImagePtr img = Image::createFromFile(path, type);
    if (!img) return -1;
    mapPathToImageIndex[path] = (int)imagePool.size();
    imagePool.push_back(img);
    return (int)imagePool.size() - 1;
}
namespace {
// "model files": what Assimp would have produced for a path (Resource.cpp:37-92), registered by the test
struct ModelFileMesh { MeshDataPtr data; std::string texturePath; int matIndex; };
struct ModelFile { std::vector<ModelFileMesh> meshes; std::vector<Material> materials; };
std::map<std::string, ModelFile>& modelFiles() { static std::map<std::string, ModelFile> m; return m; }
}
ModelInstancePtr Resource::createNewModelInstance(const File::path& path) {
    auto it = modelFiles().find(path.generic_string());
    if (it == modelFiles().end()) {
        std::fprintf(stderr, "[zillum_ref] model %s was not registered (no Assimp here)\n", path.generic_string().c_str());
        return nullptr;
    }
    auto model = std::make_shared<ModelInstance>();
    model->setPath(path);
    for (const ModelFileMesh& m : it->second.meshes) {
        auto inst = std::make_shared<MeshInstance>();                       // Resource.cpp:119-168 without the Assimp reads
        inst->matIndex = m.matIndex;
        if (!m.texturePath.empty()) inst->texIndex = Resource::addImage(m.texturePath, ImageDataType::Int8);
        inst->meshData = m.data;
        model->meshInstances().push_back(inst);
    }
    model->materials() = it->second.materials;
    return model;
}
ModelInstancePtr Resource::getModelInstanceByPath(const File::path& path) {
    auto it = mapPathToModelInstance.find(path);
    return it == mapPathToModelInstance.end() ? nullptr : it->second;
}
ModelInstancePtr Resource::openModelInstance(const File::path& path, const glm::vec3& pos, const glm::vec3& scale, const glm::vec3& rotation) {
    ModelInstancePtr raw = getModelInstanceByPath(path);                   // Resource.cpp:102-113: pooled original, per-use copy
    if (!raw) raw = createNewModelInstance(path);
    if (!raw) std::abort();
    ModelInstancePtr inst = raw->copy();
    inst->setPos(pos); inst->setScale(scale); inst->setRotation(rotation);
    return inst;
}
void Resource::clear() { imagePool.clear(); mapPathToImageIndex.clear(); meshDataPool.clear(); mapPathToModelInstance.clear(); }

// ================================================================================================
// 2. extern "C" surface
// ================================================================================================
namespace {
template <typename T> size_t copyOut(const std::vector<T>& v, void* out, size_t maxBytes) {
    size_t n = v.size() * sizeof(T);
    if (out && n <= maxBytes && n) std::memcpy(out, v.data(), n);
    return n;
}
size_t readBufferTexture(const TextureBufferedPtr& t, void* out, size_t maxBytes) {
    if (!t) return 0;
    size_t n = (size_t)t->size();
    if (out && n <= maxBytes && n) t->read(0, (int64_t)n, out);
    return n;
}
size_t readTexture2D(const Texture2DPtr& t, void* out, size_t maxBytes) {
    if (!t) return 0;
    refgl::Object* o = refgl::object(t->id());
    size_t n = o->bytes.size();
    if (out && n <= maxBytes && n) std::memcpy(out, o->bytes.data(), n);
    return n;
}
}  // namespace

extern "C" {

// ---- host preparation, one reference function per call ----

// BVH::build (BVH.cpp:116-144: quickBuild + buildHitTable).  boundsOut 6*(2T-1) floats, hitTableOut 18*(2T-1) ints.
int zr_build_bvh(const float* vertices, int numVertices, const uint32_t* indices, int numTriangles, float* boundsOut, int32_t* hitTableOut) {
    std::vector<glm::vec3> v((size_t)numVertices);
    for (int i = 0; i < numVertices; i++) v[i] = glm::vec3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
    std::vector<uint32_t> idx(indices, indices + 3 * (size_t)numTriangles);
    FILE* keep = stdout; (void)keep;
    BVH bvh(v, idx);
    PackedBVH packed = bvh.build();
    static_assert(sizeof(AABB) == 24, "AABB is two packed vec3");
    std::memcpy(boundsOut, packed.bounds.data(), packed.bounds.size() * sizeof(AABB));
    std::memcpy(hitTableOut, packed.hitTable.data(), packed.hitTable.size() * sizeof(int));
    return (int)packed.bounds.size();
}
// AliasTable::build<int32_t> (AliasTable.h:12-56)
void zr_alias_table(const float* pdf, int n, int32_t* alias, float* prob) {
    auto [a, p] = AliasTable::build<int32_t>(std::vector<float>(pdf, pdf + n));
    std::memcpy(alias, a.data(), n * sizeof(int32_t));
    std::memcpy(prob, p.data(), n * sizeof(float));
}
// Sampler::sobolSample over the reference's own matrices (Sampler.cpp:19-28, SobolMatrices256x32.h)
uint32_t zr_sobol_sample(uint32_t index, int dim) { return Sampler::sobolSample(index, dim); }
void zr_sobol_matrices(uint32_t* out /*256*32*/) { std::memcpy(out, SobolMatrices, 256 * 32 * sizeof(uint32_t)); }
// EnvironmentMap ctor (EnvironmentMap.cpp:8-59): alias / prob tables (w+1) x h, the RGB16F texels as binary32, returns sumPdf()
float zr_env_tables(const float* rgb, int w, int h, int32_t* alias, float* prob, float* texelsOut) {
    zr_register_image("mem:env", w, h, 3, rgb, nullptr);
    EnvironmentMap env("mem:env");
    readTexture2D(env.aliasTable(), alias, (size_t)(w + 1) * h * 4);
    readTexture2D(env.aliasProb(), prob, (size_t)(w + 1) * h * 4);
    if (texelsOut) readTexture2D(env.envMap(), texelsOut, (size_t)w * h * 12);
    return (float)env.sumPdf();
}
// Camera (Camera.cpp:3-7,149-162) + the camera uniforms of NaivePath.cpp:49-58
void zr_camera_update(const float* pos, const float* angleDeg, float fovDeg, float aspect, float lensRadius, float focalDist, ZlCamera* out) {
    Camera camera(glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(angleDeg[0], angleDeg[1], angleDeg[2]));
    camera.setFOV(fovDeg); camera.setAspect(aspect); camera.setLensRadius(lensRadius); camera.setFocalDist(focalDist);
    auto put = [](float* o, const glm::vec3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
    put(out->F, camera.front()); put(out->R, camera.right()); put(out->U, camera.up()); put(out->pos, camera.pos());
    glm::mat3 camMatrix(camera.right(), camera.up(), camera.front());
    glm::mat3 inv = glm::inverse(camMatrix);
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) out->matInv[3 * c + r] = inv[c][r];
    out->tanFOV = glm::tan(glm::radians(camera.FOV() * 0.5f));
    out->asp = camera.aspect(); out->lensRadius = camera.lensRadius(); out->focalDist = camera.focalDist();
}
// Sampler::genNoiseTexture (Sampler.cpp:66-80) as this libstdc++ defines std::default_random_engine
void zr_noise_texture(int w, int h, float* out) { readTexture2D(Sampler::genNoiseTexture(w, h), out, (size_t)w * h * 8); }

// ---- whole scenes through Scene::load + Scene::createGLContext (Scene.cpp:58-270) ----

void zr_full_reset(void) { Resource::clear(); modelFiles().clear(); zr_clear_images(); Pipeline::clearBindingRecord(); imageUnits().clear(); }

// register an in-memory "model file"; meshes are added with zr_full_model_add_mesh, materials with zr_full_model_set_materials
void zr_full_register_model(const char* path) { modelFiles()[path] = ModelFile(); }
// texturePath: "" or the path of an 8-bit RGB image registered with zr_register_image (the mesh's diffuse texture, Resource.cpp:148-163)
void zr_full_model_add_mesh(const char* path, int nVerts, const float* pos, const float* nrm, const float* tex, int nIdx, const uint32_t* idx,
                            const char* texturePath, int matIndex) {
    auto data = std::make_shared<MeshData>();
    for (int i = 0; i < nVerts; i++) {
        data->positions.push_back(glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
        data->normals.push_back(glm::vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
        data->texcoords.push_back(tex ? glm::vec2(tex[2 * i], tex[2 * i + 1]) : glm::vec2(0, 0));      // Resource.cpp:131-133
    }
    data->indices.assign(idx, idx + nIdx);
    modelFiles()[path].meshes.push_back(ModelFileMesh{data, texturePath ? texturePath : "", matIndex});
}
void zr_full_model_set_materials(const char* path, int n, const float* mats16) {
    static_assert(sizeof(Material) == 64, "Material is 4 texels");
    std::vector<Material>& m = modelFiles()[path].materials;
    m.resize(n);
    if (n) std::memcpy((void*)m.data(), mats16, (size_t)n * sizeof(Material));
}

// Scene::load (Scene.cpp:58-127) on the XML text, then Scene::createGLContext(true) (Scene.cpp:133-270)
void* zr_full_scene_load(const char* xmlText, const float* noise /* 2*filmW*filmH of the scene file, or NULL */) {
    char tmpl[] = "/tmp/zillum_ref_scene_XXXXXX";
    int fd = mkstemp(tmpl);
    if (fd < 0) return nullptr;
    size_t len = std::strlen(xmlText);
    if (write(fd, xmlText, len) != (ssize_t)len) { close(fd); return nullptr; }
    close(fd);
    Scene* scene = new Scene;
    bool ok = scene->load(tmpl);
    unlink(tmpl);
    if (!ok) { delete scene; return nullptr; }
    scene->createGLContext(true);
    if (noise)     // std::default_random_engine is implementation-defined (Sampler.cpp:71-73): tests substitute the product's seed image
        scene->noiseTex = Texture2D::createFromMemory(TextureFormat::Col2x32f, scene->filmWidth, scene->filmHeight,
                                                      TextureSourceFormat::Col2f, DataType::F32, noise);
    return scene;
}
void zr_full_scene_destroy(void* s) { delete (Scene*)s; }

// ints: vertexCount, triangles (indices / 3), boxCount, objPrimCount, nLightTriangles, numMaterials, filmWidth, filmHeight, sampler,
//       numTextures, envW, envH; floats: lightSumPdf, envMap->sumPdf(), envRotation
void zr_full_scene_info(void* sp, int* ints, float* floats) {
    Scene* s = (Scene*)sp;
    ints[0] = s->vertexCount; ints[1] = (int)(s->glContext.index->size() / 12); ints[2] = s->boxCount; ints[3] = s->objPrimCount;
    ints[4] = s->nLightTriangles; ints[5] = (int)s->materials.size(); ints[6] = s->filmWidth; ints[7] = s->filmHeight; ints[8] = s->sampler;
    ints[9] = s->glContext.textures ? s->glContext.textures->numTextures() : 0;
    ints[10] = s->envMap->width(); ints[11] = s->envMap->height();
    floats[0] = s->lightSumPdf; floats[1] = (float)s->envMap->sumPdf(); floats[2] = s->envRotation;
}
// the uploaded arrays of SceneGLContext (Scene.cpp:245-258) and the env / noise textures, read back byte for byte
size_t zr_full_scene_array(void* sp, const char* name, void* out, size_t maxBytes) {
    Scene* s = (Scene*)sp;
    SceneGLContext& g = s->glContext;
    const std::string n(name);
    if (n == "vertices") return readBufferTexture(g.vertex, out, maxBytes);
    if (n == "normals") return readBufferTexture(g.normal, out, maxBytes);
    if (n == "texcoords") return readBufferTexture(g.texCoord, out, maxBytes);
    if (n == "indices") return readBufferTexture(g.index, out, maxBytes);
    if (n == "bounds") return readBufferTexture(g.bound, out, maxBytes);
    if (n == "hitTable") return readBufferTexture(g.hitTable, out, maxBytes);
    if (n == "matTexIndices") return readBufferTexture(g.matTexIndex, out, maxBytes);
    if (n == "materials") return readBufferTexture(g.material, out, maxBytes);
    if (n == "lightPower") return readBufferTexture(g.lightPower, out, maxBytes);
    if (n == "lightAlias") return readBufferTexture(g.lightAlias, out, maxBytes);
    if (n == "lightProb") return readBufferTexture(g.lightProb, out, maxBytes);
    if (n == "texUVScale") return readBufferTexture(g.texUVScale, out, maxBytes);
    if (n == "envMap") return readTexture2D(s->envMap->envMap(), out, maxBytes);
    if (n == "envAlias") return readTexture2D(s->envMap->aliasTable(), out, maxBytes);
    if (n == "envAliasProb") return readTexture2D(s->envMap->aliasProb(), out, maxBytes);
    if (n == "noise") return readTexture2D(s->noiseTex, out, maxBytes);
    if (n == "texels") {
        if (!g.textures) return 0;
        refgl::Object* o = refgl::object(g.textures->id());
        if (out && o->bytes.size() <= maxBytes && !o->bytes.empty()) std::memcpy(out, o->bytes.data(), o->bytes.size());
        return o->bytes.size();
    }
    return 0;
}
void zr_full_scene_set(void* sp, const char* name, float v) {
    Scene* s = (Scene*)sp;
    const std::string n(name);
    if (n == "envRotation") s->envRotation = v;
    else if (n == "sampler") s->sampler = (int)v;
}
void zr_full_scene_camera(void* sp, ZlCamera* out) {
    Scene* s = (Scene*)sp;
    const Camera& c = s->camera;
    const glm::vec3 p = c.pos(), a = c.angle();
    const float pos[3] = {p.x, p.y, p.z}, ang[3] = {a.x, a.y, a.z};
    zr_camera_update(pos, ang, c.FOV(), c.aspect(), c.lensRadius(), c.focalDist(), out);
}

// ---- the reference's integrator host glue (NaivePath.cpp, LightPath.cpp, TriplePath.cpp), driven like Application.cpp:336-356,644-663 ----
struct ZrIntegrator { IntegratorPtr integ; std::string kind; Scene* scene; int w, h; };

void* zr_full_integrator_create(void* sp, const char* kind, int w, int h) {
    Scene* scene = (Scene*)sp;
    ZrIntegrator* z = new ZrIntegrator{nullptr, kind, scene, w, h};
    if (z->kind == "path") z->integ = std::make_shared<NaivePathIntegrator>();
    else if (z->kind == "light") z->integ = std::make_shared<LightPathIntegrator>();
    else if (z->kind == "triple") z->integ = std::make_shared<TriplePathIntegrator>();
    else { delete z; return nullptr; }
    scene->camera.setAspect((float)w / h);
    z->integ->init(scene, w, h, nullptr);
    z->integ->setStatus({scene, {w, h}, ResetLevel::FullReset});
    z->integ->setShouldReset();
    return z;
}
void zr_full_integrator_destroy(void* zp) { delete (ZrIntegrator*)zp; }
int zr_full_integrator_set(void* zp, const char* name, double v) {
    ZrIntegrator* z = (ZrIntegrator*)zp;
    const std::string n(name);
    if (z->kind == "path") {
        PathIntegParam& p = std::static_pointer_cast<NaivePathIntegrator>(z->integ)->mParam;
        if (n == "maxDepth") p.maxDepth = (int)v; else if (n == "russianRoulette") p.russianRoulette = v != 0;
        else if (n == "sampleLight") p.sampleLight = v != 0; else if (n == "lightEnvUniformSample") p.lightEnvUniformSample = v != 0;
        else if (n == "lightPortion") p.lightPortion = (float)v; else if (n == "finiteSample") p.finiteSample = v != 0;
        else if (n == "maxSample") p.maxSample = (int)v; else return 1;
    } else if (z->kind == "light") {
        LightPathIntegParam& p = std::static_pointer_cast<LightPathIntegrator>(z->integ)->mParam;
        if (n == "maxDepth") p.maxDepth = (int)v; else if (n == "russianRoulette") p.russianRoulette = v != 0;
        else if (n == "finiteSample") p.finiteSample = v != 0; else if (n == "maxSample") p.maxSample = (int)v;
        else if (n == "threadBlocksOnePass") p.threadBlocksOnePass = (int)v; else return 1;
    } else {
        TriplePathIntegParam& p = std::static_pointer_cast<TriplePathIntegrator>(z->integ)->mParam;
        if (n == "maxDepth") p.maxDepth = (int)v; else if (n == "russianRoulette") p.russianRoulette = v != 0;
        else if (n == "finiteSample") p.finiteSample = v != 0; else if (n == "maxSample") p.maxSample = (int)v;
        else if (n == "LPTBlocksOnePass") p.LPTBlocksOnePass = (int)v; else if (n == "LPTLoopsPerPass") p.LPTLoopsPerPass = (int)v; else return 1;
    }
    z->integ->setShouldReset();
    return 0;
}
void zr_full_integrator_render_one_pass(void* zp) { ((ZrIntegrator*)zp)->integ->renderOnePass(); }
float zr_full_integrator_result_scale(void* zp) { return ((ZrIntegrator*)zp)->integ->resultScale(); }
// the rgba32f frame texture the display stage reads (Application.cpp:652-660), unscaled
size_t zr_full_integrator_get_frame(void* zp, float* rgba, size_t maxBytes) { return readTexture2D(((ZrIntegrator*)zp)->integ->getFrame(), rgba, maxBytes); }

}  // extern "C"
