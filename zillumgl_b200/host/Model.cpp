#include "Model.h"
#include "ImageIO.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace zillum {

Affine ModelInstance::modelMatrix() const {
    // translate * scale * ConstRot(+90deg about X: Y-up -> Z-up) * Rz(rot.x) * Rx(rot.y) * Ry(rot.z)
    Mat3f S{{mScale.x, 0, 0}, {0, mScale.y, 0}, {0, 0, mScale.z}};
    Mat3f m = S * zillum::rotation(radians(90.0f), Vec3f(1.0f, 0.0f, 0.0f));
    m = m * zillum::rotation(radians(mRotation.x), Vec3f(0.0f, 0.0f, 1.0f));
    m = m * zillum::rotation(radians(mRotation.y), Vec3f(1.0f, 0.0f, 0.0f));
    m = m * zillum::rotation(radians(mRotation.z), Vec3f(0.0f, 1.0f, 0.0f));
    Affine a;
    a.m = m;
    a.t = mPos;
    return a;
}

ModelInstancePtr ModelInstance::copy() const {
    auto model = std::make_shared<ModelInstance>();
    *model = *this;
    for (auto& mi : model->mMeshInstances) mi = std::make_shared<MeshInstance>(*mi);
    model->mName = mName + "'";
    return model;
}

// ------------------------------------------------------------------------------------------
// Wavefront OBJ/MTL: triangulates polygons as fans, one mesh per material group, vertices
// joined on identical (v, vt, vn) triples, smooth normals generated when the file has none,
// v texture coordinate flipped (aiProcess_FlipUVs).  Diffuse colour Kd and map_Kd (PNG, JPEG,
// TGA, BMP or PPM: ImageDecode.cpp) are read from the MTL file.
// ------------------------------------------------------------------------------------------
static std::string dirName(const std::string& p) {
    size_t s = p.find_last_of("/\\");
    return s == std::string::npos ? std::string() : p.substr(0, s + 1);
}

static ModelInstancePtr loadObj(const std::string& path) {
    std::ifstream f(path);
    if (!f) { std::fprintf(stderr, "[Model] cannot open %s\n", path.c_str()); return nullptr; }
    std::vector<Vec3f> P, N;
    std::vector<Vec2f> T;
    struct Group { std::string mtl; MeshDataPtr mesh; std::map<std::string, uint32_t> join; bool hasNormals = true; };
    std::vector<Group> groups;
    std::map<std::string, int> groupOfMtl;
    std::map<std::string, Material> mtlColors;
    std::map<std::string, std::string> mtlTextures;
    std::vector<std::string> mtlOrder;
    auto loadMtl = [&](const std::string& file) {
        std::ifstream m(dirName(path) + file);
        std::string line, cur;
        while (std::getline(m, line)) {
            std::stringstream ss(line);
            std::string k;
            ss >> k;
            if (k == "newmtl") { ss >> cur; mtlColors[cur] = Material(); mtlOrder.push_back(cur); }
            else if (k == "Kd" && !cur.empty()) { Vec3f c; ss >> c.x >> c.y >> c.z; mtlColors[cur].baseColor = c; }
            else if (k == "map_Kd" && !cur.empty()) {      // the file name is the last token (options such as "-s 1 1 1" precede it); Windows-authored files use backslashes
                std::string t, last;
                while (ss >> t) last = t;
                for (char& c : last) if (c == '\\') c = '/';
                if (!last.empty()) mtlTextures[cur] = last;
            }
        }
    };
    int cur = -1;
    auto useGroup = [&](const std::string& mtl) {
        auto it = groupOfMtl.find(mtl);
        if (it == groupOfMtl.end()) {
            groupOfMtl[mtl] = (int)groups.size();
            groups.push_back({mtl, std::make_shared<MeshData>(), {}, true});
            cur = (int)groups.size() - 1;
        } else cur = it->second;
    };
    std::string line;
    while (std::getline(f, line)) {
        std::stringstream ss(line);
        std::string k;
        ss >> k;
        if (k == "v") { Vec3f v; ss >> v.x >> v.y >> v.z; P.push_back(v); }
        else if (k == "vn") { Vec3f v; ss >> v.x >> v.y >> v.z; N.push_back(v); }
        else if (k == "vt") { Vec2f v; ss >> v.x >> v.y; v.y = 1.0f - v.y; T.push_back(v); }
        else if (k == "mtllib") { std::string m; ss >> m; for (char& c : m) if (c == '\\') c = '/'; loadMtl(m); }
        else if (k == "usemtl") { std::string m; ss >> m; useGroup(m); }
        else if (k == "f") {
            if (cur < 0) useGroup("");
            Group& g = groups[cur];
            std::vector<uint32_t> poly;
            std::string tok;
            bool bad = false;
            while (ss >> tok) {
                int vi = 0, ti = 0, ni = 0;
                if (std::sscanf(tok.c_str(), "%d/%d/%d", &vi, &ti, &ni) == 3) {}
                else if (std::sscanf(tok.c_str(), "%d//%d", &vi, &ni) == 2) { ti = 0; }
                else if (std::sscanf(tok.c_str(), "%d/%d", &vi, &ti) == 2) { ni = 0; }
                else if (std::sscanf(tok.c_str(), "%d", &vi) == 1) { ti = ni = 0; }
                else { bad = true; break; }
                // negative indices count back from the elements read so far: join on the resolved triple, not on the token
                auto fix = [](int i, size_t n) { return i < 0 ? (int)n + i : i - 1; };
                const int pv = fix(vi, P.size()), pt = ti ? fix(ti, T.size()) : -1, pn = ni ? fix(ni, N.size()) : -1;
                if (pv < 0 || pv >= (int)P.size() || (ti && (pt < 0 || pt >= (int)T.size())) || (ni && (pn < 0 || pn >= (int)N.size()))) { bad = true; break; }
                const std::string key = std::to_string(pv) + "/" + std::to_string(pt) + "/" + std::to_string(pn);
                auto it = g.join.find(key);
                if (it != g.join.end()) { poly.push_back(it->second); continue; }
                if (!ni) g.hasNormals = false;
                uint32_t id = g.mesh->addVertex(P[pv], ni ? N[pn] : Vec3f(0.0f), ti ? T[pt] : Vec2f{0, 0});
                g.join[key] = id;
                poly.push_back(id);
            }
            if (bad) { std::fprintf(stderr, "[Model] %s: face with an index out of range skipped: %s\n", path.c_str(), line.c_str()); continue; }
            for (size_t i = 2; i < poly.size(); i++) g.mesh->addTriangle(poly[0], poly[i - 1], poly[i]);
        }
    }
    auto model = std::make_shared<ModelInstance>();
    model->setPath(path);
    for (auto& name : mtlOrder) model->materials().push_back(mtlColors[name]);
    if (model->materials().empty()) model->materials().push_back(Material());
    for (auto& g : groups) {
        if (g.mesh->indices.empty()) continue;
        if (!g.hasNormals) {   // area-weighted smooth normals (aiProcess_GenSmoothNormals)
            for (auto& n : g.mesh->normals) n = Vec3f(0.0f);
            for (size_t i = 0; i + 2 < g.mesh->indices.size(); i += 3) {
                uint32_t a = g.mesh->indices[i], b = g.mesh->indices[i + 1], c = g.mesh->indices[i + 2];
                Vec3f fn = cross(g.mesh->positions[b] - g.mesh->positions[a], g.mesh->positions[c] - g.mesh->positions[a]);
                g.mesh->normals[a] = g.mesh->normals[a] + fn; g.mesh->normals[b] = g.mesh->normals[b] + fn; g.mesh->normals[c] = g.mesh->normals[c] + fn;
            }
            for (auto& n : g.mesh->normals) n = (dot(n, n) > 0.0f) ? normalize(n) : Vec3f(0, 1, 0);
        }
        auto mi = std::make_shared<MeshInstance>();
        mi->meshData = g.mesh;
        mi->matIndex = 0;
        for (size_t i = 0; i < mtlOrder.size(); i++) if (mtlOrder[i] == g.mtl) mi->matIndex = (int)i;
        auto tx = mtlTextures.find(g.mtl);
        if (tx != mtlTextures.end()) mi->texIndex = Resource::addImage(dirName(path) + tx->second);
        model->meshInstances().push_back(mi);
    }
    return model;
}

namespace Resource {
static std::map<std::string, ModelInstancePtr> gModels;
static std::map<std::string, int> gImageIndex;
static std::vector<ByteImagePtr> gImages;

ModelInstancePtr openModelInstance(const std::string& path, Vec3f pos, Vec3f scale, Vec3f rotation) {
    ModelInstancePtr raw;
    auto it = gModels.find(path);
    if (it != gModels.end()) raw = it->second;
    else {
        raw = (path.rfind("builtin:", 0) == 0) ? makeBuiltinModel(path.substr(8)) : loadObj(path);
        if (!raw) return nullptr;
        raw->setPath(path);
        gModels[path] = raw;
    }
    auto copy = raw->copy();
    copy->setPos(pos);
    copy->setScale(scale);
    copy->setRotation(rotation);
    return copy;
}

int addImage(ByteImagePtr img, const std::string& key) {
    auto it = gImageIndex.find(key);
    if (it != gImageIndex.end()) return it->second;
    gImageIndex[key] = (int)gImages.size();
    gImages.push_back(img);
    return (int)gImages.size() - 1;
}

int addImage(const std::string& path) {
    auto it = gImageIndex.find(path);
    if (it != gImageIndex.end()) return it->second;
    auto img = std::make_shared<ByteImage>();
    if (!loadByteImage(path, img->rgb, img->width, img->height)) return -1;
    return addImage(img, path);
}

const std::vector<ByteImagePtr>& getAllImages() { return gImages; }

void clear() { gModels.clear(); gImageIndex.clear(); gImages.clear(); }
}  // namespace Resource

}  // namespace zillum

