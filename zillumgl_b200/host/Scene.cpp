#include "Scene.h"
#include <chrono>
#include <cstdio>
#include <sstream>
#include "AliasTable.h"
#include "Xml.h"

namespace zillum {

Scene::~Scene() { if (glContext) zl_scene_destroy(glContext); }

static Vec3f parseVec3(const std::string& s, Vec3f def = Vec3f(0.0f)) {
    std::stringstream ss(s);
    Vec3f v = def;
    ss >> v.x >> v.y >> v.z;
    return v;
}

// <modelInstance path= name= type=object|light> (Scene.cpp:23-56)
static bool loadModelInstance(Scene& scene, const XmlNode& node, const std::string& baseDir) {
    std::string path = node.attribute("path");
    if (path.rfind("builtin:", 0) != 0 && !path.empty() && path[0] != '/') path = baseDir + path;
    auto model = Resource::openModelInstance(path, Vec3f(0.0f));
    if (!model) return false;
    model->setName(node.attribute("name"));
    XmlNode t = node.child("transform");
    model->setPos(parseVec3(t.attribute("translate")));
    Vec3f s = parseVec3(t.attribute("scale"));
    model->setScale(s.x, s.y, s.z);
    model->setRotation(parseVec3(t.attribute("rotate")));
    if (node.attribute("type") == "light") {
        scene.addLight(model, parseVec3(node.child("radiance").attribute("value")));
        return true;
    }
    auto material = loadMaterial(node.child("material"));
    if (material)
        for (auto& m : model->materials()) m = material.value();
    scene.addObject(model);
    return true;
}

bool Scene::load(const std::string& path) {
    std::string err;
    XmlNode doc = XmlNode::parseFile(path, &err);
    if (!doc) { std::fprintf(stderr, "[Scene] %s: %s\n", path.c_str(), err.c_str()); return false; }
    size_t slash = path.find_last_of("/\\");
    std::string baseDir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
    return loadXml(doc, baseDir);
}

bool Scene::loadXmlText(const std::string& text, const std::string& baseDir) {
    std::string err;
    XmlNode doc = XmlNode::parseString(text, &err);
    if (!doc) { std::fprintf(stderr, "[Scene] xml: %s\n", err.c_str()); return false; }
    return loadXml(doc, baseDir);
}

bool Scene::loadXml(const XmlNode& doc, const std::string& baseDir) {
    Resource::clear();
    clear();
    XmlNode scene = doc.child("scene");
    if (!scene) { std::fprintf(stderr, "[Scene] no <scene> element\n"); return false; }
    {
        XmlNode integrator = scene.child("integrator");
        integratorType = integrator.attribute("type");
        XmlNode size = integrator.child("size");
        filmWidth = size.attributeInt("width");
        filmHeight = size.attributeInt("height");
    }
    sampler = (scene.child("sampler").attribute("type") == "sobol") ? 1 : 0;
    {
        XmlNode cam = scene.child("camera");
        camera.setPos(parseVec3(cam.child("position").attribute("value")));
        camera.setAngle(parseVec3(cam.child("angle").attribute("value")));
        camera.setFOV(cam.child("fov").attributeFloat("value"));
        camera.setAspect(static_cast<float>(filmWidth) / filmHeight);
        camera.setLensRadius(cam.child("lensRadius").attributeFloat("value"));
        camera.setFocalDist(cam.child("focalDistance").attributeFloat("value"));
        originalCamera = previewCamera = camera;
    }
    for (auto& inst : scene.child("modelInstances").children())
        if (!loadModelInstance(*this, inst, baseDir)) return false;
    {
        std::string env = scene.child("envMap").attribute("path");
        if (!env.empty() && env.rfind("builtin:", 0) != 0 && env[0] != '/') env = baseDir + env;
        envMap = EnvironmentMap::create(env);
        envRotation = scene.child("envMap").hasAttribute("rotation") ? scene.child("envMap").attributeFloat("rotation") : 0.0f;
    }
    return true;
}

void Scene::clear() {
    objects.clear();
    lights.clear();
    materials.clear();
}

void Scene::flatten(bool resetTextures) {
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    SceneHostArrays& h = host;
    h.vertices.clear(); h.normals.clear(); h.texCoords.clear(); h.indices.clear(); h.matTexIndices.clear();
    h.lightMeshFirstTri.clear(); h.lightMeshNumTris.clear(); h.lightMeshPower.clear();
    std::vector<Material> sceneMaterials = materials;   // the reference appends to `materials` on every call; we rebuild
    lightSumPdf = 0.0f;
    nLightTriangles = 0;
    objPrimCount = 0;
    uint32_t offIndVertex = 0, offIndMaterial = (uint32_t)sceneMaterials.size();

    // Scene.cpp:197-243 appends element by element; here the sizes are known first and every mesh fills its own slice of the arrays
    // (same arithmetic per element, so the same arrays), in parallel over the elements of a mesh: 12.6 M vertices in the Rungholt-class scene
    {
        size_t nv = 0, ni = 0, nt = 0, np = 0;
        for (auto& object : objects)
            for (auto& mi : object->meshInstances()) {
                nv += mi->meshData->positions.size(); ni += mi->meshData->indices.size();
                nt += mi->meshData->texcoords.size(); np += mi->meshData->indices.size() / 3;
            }
        for (auto& light : lights)
            for (auto& mi : light.first->meshInstances()) { nv += mi->meshData->positions.size(); ni += mi->meshData->indices.size(); }
        h.vertices.resize(nv); h.normals.resize(nv); h.texCoords.resize(nt); h.indices.resize(ni); h.matTexIndices.resize(np);
    }
    size_t atVertex = 0, atIndex = 0, atTex = 0, atPrim = 0;
    auto appendGeometry = [&](ModelInstance& inst, bool isObject) {
        Affine model = inst.modelMatrix();
        Mat3f normalMat = normalMatrix(model);      // mat3(transpose(inverse(mat4))), Scene.cpp:153-154
        for (auto& mi : inst.meshInstances()) {
            const MeshData& md = *mi->meshData;
            const long nPos = (long)md.positions.size(), nNrm = (long)std::min(md.normals.size(), md.positions.size()), nInd = (long)md.indices.size();
            Vec3f* outV = h.vertices.data() + atVertex;
            Vec3f* outN = h.normals.data() + atVertex;
            uint32_t* outI = h.indices.data() + atIndex;
#pragma omp parallel for schedule(static) if (nPos > (1 << 16))
            for (long i = 0; i < nPos; i++) outV[i] = model.point(md.positions[i]);
#pragma omp parallel for schedule(static) if (nNrm > (1 << 16))
            for (long i = 0; i < nNrm; i++) outN[i] = normalize(normalMat * md.normals[i]);
#pragma omp parallel for schedule(static) if (nInd > (1 << 16))
            for (long i = 0; i < nInd; i++) outI[i] = md.indices[i] + offIndVertex;
            if (isObject) {
                std::copy(md.texcoords.begin(), md.texcoords.end(), h.texCoords.begin() + atTex);
                atTex += md.texcoords.size();
                const uint32_t mt = offIndMaterial + (uint32_t)(mi->texIndex << 16 | mi->matIndex);
                std::fill(h.matTexIndices.begin() + atPrim, h.matTexIndices.begin() + atPrim + nInd / 3, mt);
                atPrim += nInd / 3;
                mi->globalMatIndex = (mi->matIndex != -1) ? mi->matIndex + (int)offIndMaterial : -1;
                objPrimCount += (int)(nInd / 3);
            }
            atVertex += nPos; atIndex += nInd;
            offIndVertex += (uint32_t)nPos;
        }
    };
    for (auto& object : objects) {
        sceneMaterials.insert(sceneMaterials.end(), object->materials().begin(), object->materials().end());
        appendGeometry(*object, true);
        offIndMaterial += (uint32_t)object->materials().size();
    }
    for (auto& light : lights) appendGeometry(*light.first, false);
    h.materials = sceneMaterials;

    if (buildBvhOnDevice || h.indices.empty()) {   // (no triangles: the reference's BVH::build would index treeSize = -1, BVH.cpp:120; the callers report it)
        bvhBuildSeconds = bvhFlattenSeconds = 0.0;
        h.bounds.clear(); h.hitTable.clear(); h.sizeIndices.clear();
    } else {
        BVH bvh(h.vertices, h.indices);
        PackedBVH packed = bvh.build(!threadMtbvhOnDevice);
        bvhBuildSeconds = bvh.buildSeconds;
        bvhFlattenSeconds = bvh.flattenSeconds;
        h.bounds = std::move(packed.bounds);
        h.hitTable = std::move(packed.hitTable);
        h.sizeIndices = std::move(packed.sizeIndices);
    }

    // light sampling table: mesh power split by triangle area (Scene.cpp:200-243)
    h.lightPower.clear();
    std::vector<float> pdf;
    int tri = objPrimCount;
    for (auto& light : lights) {
        const Vec3f sumPower = light.second;
        for (auto& mi : light.first->meshInstances()) {
            int nTris = (int)(mi->meshData->indices.size() / 3);
            h.lightMeshFirstTri.push_back(tri); h.lightMeshNumTris.push_back(nTris); h.lightMeshPower.push_back(sumPower);
            auto area2 = [&](int t) {
                Vec3f va = h.vertices[h.indices[3 * t]], vb = h.vertices[h.indices[3 * t + 1]], vc = h.vertices[h.indices[3 * t + 2]];
                return length(cross(vc - va, vb - va));
            };
            float sumArea = 0.0f;
            for (int i = 0; i < nTris; i++) sumArea += area2(tri + i);
            for (int i = 0; i < nTris; i++) {
                Vec3f power = sumPower * area2(tri + i) / sumArea;
                float lum = dot(power, Vec3f(0.299f, 0.587f, 0.114f));
                h.lightPower.push_back(power);
                pdf.push_back(lum);
                lightSumPdf += lum;
            }
            nLightTriangles += nTris;
            tri += nTris;
        }
    }
    auto table = AliasTable::build<int32_t>(pdf);
    h.lightAlias = std::move(table.first);
    h.lightProb = std::move(table.second);

    if (resetTextures) {
        // Texture2DArray: layers padded to the largest image (Texture.cpp:134-171)
        const auto& images = Resource::getAllImages();
        h.numTextures = (int)images.size();
        h.texMaxW = h.texMaxH = 0;
        for (auto& img : images) if (img) { h.texMaxW = std::max(h.texMaxW, img->width); h.texMaxH = std::max(h.texMaxH, img->height); }
        h.texels.assign((size_t)h.numTextures * h.texMaxW * h.texMaxH * 3, 0);
        h.texScales.assign(h.numTextures, Vec2f{0, 0});
        for (int i = 0; i < h.numTextures; i++) {
            auto& img = images[i];
            if (!img) continue;
            for (int y = 0; y < img->height; y++)
                std::copy(&img->rgb[(size_t)y * img->width * 3], &img->rgb[(size_t)(y + 1) * img->width * 3],
                          &h.texels[(((size_t)i * h.texMaxH + y) * h.texMaxW) * 3]);
            h.texScales[i] = Vec2f{(float)img->width / h.texMaxW, (float)img->height / h.texMaxH};
        }
        h.noise = Sampler::genNoiseTexture(filmWidth, filmHeight);
    }
    if (!envMap) envMap = EnvironmentMap::createBlack();
    vertexCount = (int)h.vertices.size();
    triangleCount = (int)h.vertices.size() / 3;   // sic (Scene.cpp:266)
    boxCount = 2 * (int)(h.indices.size() / 3) - 1;
    flattenSeconds = std::chrono::duration<double>(clk::now() - t0).count();
}

ZlSceneDesc Scene::desc() const {
    const SceneHostArrays& h = host;
    ZlSceneDesc d{};
    d.vertices = &h.vertices[0].x; d.normals = &h.normals[0].x;
    d.texcoords = h.texCoords.empty() ? nullptr : &h.texCoords[0].x;
    d.indices = h.indices.data();
    d.bounds = h.bounds.empty() ? nullptr : &h.bounds[0].pMin.x;
    d.hitTable = h.hitTable.empty() ? nullptr : h.hitTable.data();
    d.sizeIndices = h.sizeIndices.empty() ? nullptr : h.sizeIndices.data();
    d.matTexIndices = (const int32_t*)h.matTexIndices.data();
    d.materials = &h.materials[0].baseColor.x;
    d.lightPower = h.lightPower.empty() ? nullptr : &h.lightPower[0].x;
    d.lightAlias = h.lightAlias.data(); d.lightProb = h.lightProb.data();
    d.texels = h.texels.empty() ? nullptr : h.texels.data();
    d.texUVScale = h.texScales.empty() ? nullptr : &h.texScales[0].x;
    d.envMap = envMap->pixels().data(); d.envAlias = envMap->aliasTable().data(); d.envAliasProb = envMap->aliasProb().data();
    d.noise = h.noise.data();
    d.sobolMatrices = Sampler::SobolMatrices;
    d.numVertices = (int)h.vertices.size(); d.numTexcoords = (int)h.texCoords.size();
    d.numTriangles = (int)(h.indices.size() / 3); d.bvhSize = 2 * d.numTriangles - 1;
    d.objPrimCount = objPrimCount; d.numMaterials = (int)h.materials.size(); d.numLightTriangles = nLightTriangles;
    d.numTextures = h.numTextures; d.texMaxW = h.texMaxW; d.texMaxH = h.texMaxH;
    d.envW = envMap->width(); d.envH = envMap->height();
    d.noiseW = filmWidth; d.noiseH = filmHeight;
    d.lightSum = lightSumPdf;
    d.envSum = (float)envMap->sumPdf();
    return d;
}

int Scene::upload() {
    if (glContext) { zl_scene_destroy(glContext); glContext = nullptr; }
    ZlSceneDesc d = desc();
    int rc = zl_scene_create(&d, &glContext);
    if (rc != 0) std::fprintf(stderr, "[Scene] zl_scene_create failed: %s\n", zl_last_error_string());
    return rc;
}

}  // namespace zillum
