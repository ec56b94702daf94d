// C shim over the C++ host classes (see include/zillum_host.h).
#include <cstring>
#include <string>
#include "../../include/zillum_host.h"
#include "AliasTable.h"
#include "ImageIO.h"
#include "Integrator.h"

using namespace zillum;

struct ZhScene { Scene scene; ZlSceneDesc desc; };
struct ZhIntegrator { IntegratorPtr integ; std::string type; ZhScene* scene; int w, h; };

extern "C" {

ZhScene* zh_scene_create(void) { return new ZhScene(); }
void zh_scene_destroy(ZhScene* s) { delete s; }
int zh_scene_load(ZhScene* s, const char* path) { return s->scene.load(path) ? 0 : 1; }
int zh_scene_load_builtin(ZhScene* s, const char* name, int w, int h) { return s->scene.loadBuiltin(name, w, h) ? 0 : 1; }
int zh_scene_load_xml_text(ZhScene* s, const char* xml) { return s->scene.loadXmlText(xml) ? 0 : 1; }
int zh_scene_flatten(ZhScene* s) { s->scene.flatten(true); s->desc = s->scene.desc(); return s->scene.host.indices.empty() ? 1 : 0; }   // 1: a scene without triangles cannot be rendered
int zh_scene_upload(ZhScene* s) { return s->scene.upload(); }
const ZlSceneDesc* zh_scene_desc(ZhScene* s) { s->desc = s->scene.desc(); return &s->desc; }
ZlScene* zh_scene_device(ZhScene* s) { return s->scene.glContext; }
void zh_scene_info(ZhScene* s, int* info) {
    const Scene& sc = s->scene;
    int v[13] = {(int)sc.host.vertices.size(), (int)(sc.host.indices.size() / 3), sc.boxCount, sc.objPrimCount, sc.nLightTriangles,
                 (int)sc.host.materials.size(), sc.filmWidth, sc.filmHeight, sc.sampler, sc.host.numTextures,
                 sc.envMap ? sc.envMap->width() : 0, sc.envMap ? sc.envMap->height() : 0, (int)sc.host.lightMeshFirstTri.size()};
    std::memcpy(info, v, sizeof v);
}
void zh_scene_times(ZhScene* s, double* t) { t[0] = s->scene.bvhBuildSeconds; t[1] = s->scene.bvhFlattenSeconds; t[2] = s->scene.flattenSeconds; }
void zh_scene_light_meshes(ZhScene* s, int* first, int* num, float* power) {
    const auto& h = s->scene.host;
    for (size_t i = 0; i < h.lightMeshFirstTri.size(); i++) {
        first[i] = h.lightMeshFirstTri[i]; num[i] = h.lightMeshNumTris[i];
        power[3 * i] = h.lightMeshPower[i].x; power[3 * i + 1] = h.lightMeshPower[i].y; power[3 * i + 2] = h.lightMeshPower[i].z;
    }
}
void zh_scene_set_camera(ZhScene* s, const float* pos, const float* ang, float fov, float lens, float focal) {
    Camera& c = s->scene.camera;
    c.setPos(Vec3f(pos[0], pos[1], pos[2]));
    c.setAngle(Vec3f(ang[0], ang[1], ang[2]));
    c.setFOV(fov);
    c.setLensRadius(lens);
    c.setFocalDist(focal);
}
void zh_scene_camera(ZhScene* s, ZlCamera* out) { *out = s->scene.camera.uniforms(); }
void zh_scene_set_sampler(ZhScene* s, int sampler) { s->scene.sampler = sampler; }
void zh_scene_set_device_mtbvh(ZhScene* s, int on) { s->scene.threadMtbvhOnDevice = on != 0; }
void zh_scene_set_device_bvh(ZhScene* s, int on) { s->scene.buildBvhOnDevice = on != 0; }
void zh_scene_set_env_rotation(ZhScene* s, float r) { s->scene.envRotation = r; }
// ---- the scene BEFORE flattening, for tests that feed the same models to the reference's own Scene (oracle/_ref) ----
static ModelInstancePtr modelAt(ZhScene* s, int m, Vec3f* power) {
    const Scene& sc = s->scene;
    if (m < (int)sc.objects.size()) return sc.objects[m];
    m -= (int)sc.objects.size();
    if (power) *power = sc.lights[m].second;
    return sc.lights[m].first;
}
int zh_scene_num_models(ZhScene* s) { return (int)(s->scene.objects.size() + s->scene.lights.size()); }
void zh_scene_model_info(ZhScene* s, int m, int* info, float* trs9, float* power3, char* pathOut, int pathCap) {
    Vec3f power(0.0f);
    ModelInstancePtr model = modelAt(s, m, &power);
    info[0] = m >= (int)s->scene.objects.size(); info[1] = (int)model->meshInstances().size(); info[2] = (int)model->materials().size();
    const Vec3f t = model->pos(), sc = model->scale(), r = model->rotation();
    const float v[9] = {t.x, t.y, t.z, sc.x, sc.y, sc.z, r.x, r.y, r.z};
    std::memcpy(trs9, v, sizeof v);
    power3[0] = power.x; power3[1] = power.y; power3[2] = power.z;
    std::strncpy(pathOut, model->path().c_str(), pathCap - 1); pathOut[pathCap - 1] = 0;
}
void zh_scene_model_mesh_counts(ZhScene* s, int m, int k, int* counts) {
    const MeshInstancePtr& mi = modelAt(s, m, nullptr)->meshInstances()[k];
    counts[0] = (int)mi->meshData->positions.size(); counts[1] = (int)mi->meshData->indices.size(); counts[2] = mi->texIndex; counts[3] = mi->matIndex;
}
void zh_scene_model_mesh_data(ZhScene* s, int m, int k, float* pos, float* nrm, float* tex, uint32_t* idx) {
    const MeshData& d = *modelAt(s, m, nullptr)->meshInstances()[k]->meshData;
    std::memcpy(pos, d.positions.data(), d.positions.size() * sizeof(Vec3f));
    std::memcpy(nrm, d.normals.data(), d.normals.size() * sizeof(Vec3f));
    std::memcpy(tex, d.texcoords.data(), d.texcoords.size() * sizeof(Vec2f));
    std::memcpy(idx, d.indices.data(), d.indices.size() * sizeof(uint32_t));
}
void zh_scene_model_materials(ZhScene* s, int m, float* mats16) {
    auto& mats = modelAt(s, m, nullptr)->materials();
    if (!mats.empty()) std::memcpy(mats16, mats.data(), mats.size() * sizeof(Material));
}
int zh_num_images(void) { return (int)Resource::getAllImages().size(); }
void zh_image(int i, int* w, int* h, unsigned char* rgb8) {
    const ByteImagePtr& img = Resource::getAllImages()[i];
    *w = img->width; *h = img->height;
    if (rgb8) std::memcpy(rgb8, img->rgb.data(), img->rgb.size());
}
const char* zh_builtin_scene_xml(const char* name, int w, int h) {
    static std::string buf;
    buf = Scene::builtinXml(name, w, h);
    return buf.c_str();
}

ZhIntegrator* zh_integrator_create(const char* type, ZhScene* s, int w, int h, void* externalFilm, void* stream) {
    auto* z = new ZhIntegrator();
    z->type = type; z->scene = s; z->w = w; z->h = h;
    if (z->type == "path") z->integ = std::make_shared<NaivePathIntegrator>();
    else if (z->type == "light") z->integ = std::make_shared<LightPathIntegrator>();
    else if (z->type == "triple") z->integ = std::make_shared<TriplePathIntegrator>();
    else { delete z; return nullptr; }
    if (externalFilm) z->integ->setExternalFilm(externalFilm);
    s->scene.camera.setAspect((float)w / h);   // Application keeps aspect = render size (Application.cpp reset path)
    z->integ->init(&s->scene, w, h, stream);
    RenderStatus st; st.scene = &s->scene; st.renderSize[0] = w; st.renderSize[1] = h; st.resetLevel = ResetLevel::ResetFrame;
    z->integ->setStatus(st);
    z->integ->reset(st);
    return z;
}
void zh_integrator_destroy(ZhIntegrator* z) { delete z; }

int zh_integrator_set(ZhIntegrator* z, const char* nameC, double v) {
    std::string name = nameC;
    if (name == "dryRun") { z->integ->setDryRun(v != 0); return 0; }
    if (auto* p = dynamic_cast<NaivePathIntegrator*>(z->integ.get())) {
        auto& m = p->mParam;
        if (name == "maxDepth") m.maxDepth = (int)v; else if (name == "russianRoulette") m.russianRoulette = v != 0;
        else if (name == "sampleLight") m.sampleLight = v != 0; else if (name == "lightEnvUniformSample") m.lightEnvUniformSample = v != 0;
        else if (name == "lightPortion") m.lightPortion = (float)v; else if (name == "finiteSample") m.finiteSample = v != 0;
        else if (name == "maxSample") m.maxSample = (int)v; else if (name == "kernelVariant") m.kernelVariant = (int)v;
        else return 1;
        return 0;
    }
    if (auto* p = dynamic_cast<LightPathIntegrator*>(z->integ.get())) {
        auto& m = p->mParam;
        if (name == "maxDepth") m.maxDepth = (int)v; else if (name == "russianRoulette") m.russianRoulette = v != 0;
        else if (name == "finiteSample") m.finiteSample = v != 0; else if (name == "maxSample") m.maxSample = (int)v;
        else if (name == "threadBlocksOnePass") m.threadBlocksOnePass = (int)v; else if (name == "kernelVariant") m.kernelVariant = (int)v;
        else return 1;
        return 0;
    }
    if (auto* p = dynamic_cast<TriplePathIntegrator*>(z->integ.get())) {
        auto& m = p->mParam;
        if (name == "maxDepth") m.maxDepth = (int)v; else if (name == "russianRoulette") m.russianRoulette = v != 0;
        else if (name == "finiteSample") m.finiteSample = v != 0; else if (name == "maxSample") m.maxSample = (int)v;
        else if (name == "LPTBlocksOnePass") m.LPTBlocksOnePass = (int)v; else if (name == "LPTLoopsPerPass") m.LPTLoopsPerPass = (int)v;
        else if (name == "kernelVariant") m.kernelVariant = (int)v;
        else return 1;
        return 0;
    }
    return 1;
}
double zh_integrator_get(ZhIntegrator* z, const char* nameC) {
    std::string name = nameC;
    if (auto* p = dynamic_cast<NaivePathIntegrator*>(z->integ.get())) {
        auto& m = p->mParam;
        if (name == "maxDepth") return m.maxDepth; if (name == "russianRoulette") return m.russianRoulette;
        if (name == "sampleLight") return m.sampleLight; if (name == "lightEnvUniformSample") return m.lightEnvUniformSample;
        if (name == "lightPortion") return m.lightPortion; if (name == "maxSample") return m.maxSample;
        if (name == "finiteSample") return m.finiteSample; if (name == "sampler") return m.sampler; if (name == "kernelVariant") return m.kernelVariant;
    }
    if (auto* p = dynamic_cast<LightPathIntegrator*>(z->integ.get())) {
        auto& m = p->mParam;
        if (name == "maxDepth") return m.maxDepth; if (name == "russianRoulette") return m.russianRoulette;
        if (name == "threadBlocksOnePass") return m.threadBlocksOnePass; if (name == "samplePerPixel") return m.samplePerPixel;
        if (name == "maxSample") return m.maxSample; if (name == "finiteSample") return m.finiteSample; if (name == "kernelVariant") return m.kernelVariant;
    }
    if (auto* p = dynamic_cast<TriplePathIntegrator*>(z->integ.get())) {
        auto& m = p->mParam;
        if (name == "maxDepth") return m.maxDepth; if (name == "russianRoulette") return m.russianRoulette;
        if (name == "LPTBlocksOnePass") return m.LPTBlocksOnePass; if (name == "LPTLoopsPerPass") return m.LPTLoopsPerPass;
        if (name == "samplePerPixel") return m.samplePerPixel; if (name == "maxSample") return m.maxSample;
        if (name == "finiteSample") return m.finiteSample; if (name == "PTSampler") return m.PTSampler; if (name == "kernelVariant") return m.kernelVariant;
    }
    return -1e300;
}
void zh_integrator_set_sample_shard(ZhIntegrator* z, int first, int stride) { z->integ->setSampleShard(first, stride); }
int zh_integrator_render_one_pass(ZhIntegrator* z) { z->integ->renderOnePass(); return z->integ->lastError(); }
void zh_integrator_reset(ZhIntegrator* z) {
    RenderStatus st; st.scene = &z->scene->scene; st.renderSize[0] = z->w; st.renderSize[1] = z->h; st.resetLevel = ResetLevel::ResetFrame;
    z->integ->setStatus(st);
    z->integ->reset(st);
}
void zh_integrator_params(ZhIntegrator* z, int kernel, ZlRenderParams* out) { *out = z->integ->params(kernel); }
ZlFilm* zh_integrator_film(ZhIntegrator* z) { return z->integ->film(); }
float zh_integrator_result_scale(ZhIntegrator* z) { return z->integ->resultScale(); }
float zh_integrator_true_scale(ZhIntegrator* z) { return z->integ->trueScale(); }
int zh_integrator_cur_sample(ZhIntegrator* z) { return z->integ->curSample(); }
int zh_integrator_get_frame(ZhIntegrator* z, float scale, float* rgba) {
    if (scale <= 0.0f) scale = z->integ->trueScale();
    return z->integ->downloadFrame(scale, rgba);      // on the integrator's own stream: ordered after its passes
}

int zh_integrator_flush(ZhIntegrator* z) { return z->integ->flush(); }
int zh_integrator_snapshot_async(ZhIntegrator* z, void* dstDevice) { return z->integ->snapshotAsync(dstDevice); }
int zh_integrator_get_frame_async(ZhIntegrator* z, float scale, float* rgbaPinned) { return z->integ->getFrameAsync(rgbaPinned, scale); }
int zh_integrator_get_frame_rgb_async(ZhIntegrator* z, float scale, float* rgbPinned) { return z->integ->getFrameAsync(rgbPinned, scale, 3); }
int zh_integrator_wait_frame(ZhIntegrator* z) { return z->integ->waitFrame(); }

int zh_build_bvh(const float* vertices, int numVertices, const uint32_t* indices, int numTriangles,
                 float* boundsOut, int32_t* hitTableOut, double* seconds2) {
    std::vector<Vec3f> v(numVertices);
    for (int i = 0; i < numVertices; i++) v[i] = Vec3f(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
    std::vector<uint32_t> idx(indices, indices + 3 * (size_t)numTriangles);
    BVH bvh(v, idx);
    PackedBVH p = bvh.build();
    std::memcpy(boundsOut, p.bounds.data(), p.bounds.size() * sizeof(AABB));
    std::memcpy(hitTableOut, p.hitTable.data(), p.hitTable.size() * sizeof(int));
    if (seconds2) { seconds2[0] = bvh.buildSeconds; seconds2[1] = bvh.flattenSeconds; }
    return (int)p.bounds.size();
}
void zh_alias_table(const float* pdf, int n, int32_t* alias, float* prob) {
    auto t = AliasTable::build<int32_t>(std::vector<float>(pdf, pdf + n));
    std::memcpy(alias, t.first.data(), sizeof(int32_t) * n);
    std::memcpy(prob, t.second.data(), sizeof(float) * n);
}
float zh_env_tables(const float* rgb, int w, int h, int32_t* alias, float* prob) {
    EnvironmentMap env(std::vector<float>(rgb, rgb + 3 * (size_t)w * h), w, h);
    std::memcpy(alias, env.aliasTable().data(), sizeof(int32_t) * env.aliasTable().size());
    std::memcpy(prob, env.aliasProb().data(), sizeof(float) * env.aliasProb().size());
    return (float)env.sumPdf();
}
uint32_t zh_sobol_sample(uint32_t index, int dim) { return Sampler::sobolSample(index, dim); }
void zh_noise_texture(int w, int h, float* out) {
    auto n = Sampler::genNoiseTexture(w, h);
    std::memcpy(out, n.data(), n.size() * sizeof(float));
}
int zh_write_pfm(const char* path, const float* rgba, int w, int h) { return writePFM(path, rgba, w, h) ? 0 : 1; }
int zh_write_exr(const char* path, const float* rgba, int w, int h) { return writeEXR(path, rgba, w, h) ? 0 : 1; }
int zh_write_png(const char* path, const unsigned char* rgb8, int w, int h) { return writePNG(path, rgb8, w, h) ? 0 : 1; }
int zh_load_byte_image(const char* path, int* w, int* h, unsigned char* rgb8) {
    std::vector<unsigned char> rgb;
    int ww = 0, hh = 0;
    if (!path || !loadByteImage(path, rgb, ww, hh)) return 1;
    if (w) *w = ww;
    if (h) *h = hh;
    if (rgb8) std::memcpy(rgb8, rgb.data(), rgb.size());
    return 0;
}
int zh_integrator_post_process(ZhIntegrator* z, float scale, int toneMapper, float* rgba, unsigned char* rgb8) {
    return z->integ->postProcess(scale, toneMapper, rgba, rgb8);
}

}  // extern "C"
