#include "Model.h"
#include "ImageIO.h"
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

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

// (position, texcoord, normal) indices of one face corner after resolving negative references; -1 = absent
struct JoinKey {
    int v, t, n;
    bool operator==(const JoinKey& o) const { return v == o.v && t == o.t && n == o.n; }
};
struct JoinHash {
    size_t operator()(const JoinKey& k) const {
        uint64_t h = (uint64_t)(uint32_t)k.v * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)k.t + 0x7F4A7C15u) * 0xC2B2AE3D27D4EB4Full;
        h ^= ((uint64_t)(uint32_t)k.n + 0x165667B1u) * 0xD6E8FEB86659FD93ull;
        return (size_t)(h ^ (h >> 29));
    }
};

static ModelInstancePtr loadObj(const std::string& path) {
    std::string text;
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) { std::fprintf(stderr, "[Model] cannot open %s\n", path.c_str()); return nullptr; }
        f.seekg(0, std::ios::end);
        const std::streamoff size = f.tellg();
        f.seekg(0, std::ios::beg);
        text.resize(size > 0 ? (size_t)size : 0);
        if (size > 0) f.read(&text[0], size);
        text.resize((size_t)std::max<std::streamsize>(f.gcount(), 0));
    }
    std::vector<Vec3f> P, N;
    std::vector<Vec2f> T;
    struct Group { std::string mtl; MeshDataPtr mesh; std::unordered_map<JoinKey, uint32_t, JoinHash> join; bool hasNormals = true; };
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
    // One pass over the bytes of the file (a Rungholt-sized OBJ is ~300 MB / 10 M lines: no stream object per line, no string keys):
    // tokens are runs of non-blank characters, numbers go through std::from_chars (correctly rounded, like the stream extraction
    // it replaces), vertices are joined on the resolved integer triple.
    const char* p = text.data();
    const char* const end = p + text.size();
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
    auto skipBlanks = [&](const char*& q, const char* e) { while (q < e && blank(*q)) q++; };
    auto token = [&](const char*& q, const char* e, const char*& t0, const char*& t1) {     // next token of the line [q, e)
        skipBlanks(q, e);
        t0 = q;
        while (q < e && !blank(*q)) q++;
        t1 = q;
        return t1 > t0;
    };
    auto number = [&](const char*& q, const char* e) {                                       // a missing or malformed component reads as 0
        const char *t0, *t1;
        float v = 0.0f;
        if (token(q, e, t0, t1)) {
            if (*t0 == '+') t0++;
            if (std::from_chars(t0, t1, v).ec != std::errc()) v = 0.0f;
        }
        return v;
    };
    auto integer = [](const char*& q, const char* e, int& v) {                                // optional sign + digits; false when there is no digit
        const char* s = q;
        if (s < e && (*s == '+' || *s == '-')) s++;
        if (s >= e || *s < '0' || *s > '9') return false;
        auto r = std::from_chars(*q == '+' ? q + 1 : q, e, v);
        if (r.ec != std::errc()) return false;
        q = r.ptr;
        return true;
    };
    std::vector<uint32_t> poly;
    while (p < end) {
        const char* eol = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char* q = p;
        const char* const lineStart = p;
        p = eol < end ? eol + 1 : end;
        const char *k0, *k1;
        if (!token(q, eol, k0, k1)) continue;
        const size_t kl = (size_t)(k1 - k0);
        auto word = [&]() { const char *t0, *t1; return token(q, eol, t0, t1) ? std::string(t0, t1) : std::string(); };
        if (kl == 1 && *k0 == 'v') { Vec3f v; v.x = number(q, eol); v.y = number(q, eol); v.z = number(q, eol); P.push_back(v); }
        else if (kl == 2 && k0[0] == 'v' && k0[1] == 'n') { Vec3f v; v.x = number(q, eol); v.y = number(q, eol); v.z = number(q, eol); N.push_back(v); }
        else if (kl == 2 && k0[0] == 'v' && k0[1] == 't') { Vec2f v; v.x = number(q, eol); v.y = number(q, eol); v.y = 1.0f - v.y; T.push_back(v); }
        else if (kl == 6 && !std::memcmp(k0, "mtllib", 6)) { std::string m = word(); for (char& c : m) if (c == '\\') c = '/'; loadMtl(m); }
        else if (kl == 6 && !std::memcmp(k0, "usemtl", 6)) { useGroup(word()); }
        else if (kl == 1 && *k0 == 'f') {
            if (cur < 0) useGroup("");
            Group& g = groups[cur];
            poly.clear();
            bool bad = false;
            const char *t0, *t1;
            while (token(q, eol, t0, t1)) {
                // v, v/vt, v//vn or v/vt/vn; an empty or unreadable trailing field counts as absent
                int vi = 0, ti = 0, ni = 0;
                const char* c = t0;
                if (!integer(c, t1, vi)) { bad = true; break; }
                if (c < t1 && *c == '/') {
                    c++;
                    if (c < t1 && *c == '/') { c++; if (!integer(c, t1, ni)) ni = 0; }
                    else if (integer(c, t1, ti)) { if (c < t1 && *c == '/') { c++; if (!integer(c, t1, ni)) ni = 0; } }
                    else ti = 0;
                }
                // negative indices count back from the elements read so far: join on the resolved triple, not on the token
                auto fix = [](int i, size_t n) { return i < 0 ? (int)n + i : i - 1; };
                const int pv = fix(vi, P.size()), pt = ti ? fix(ti, T.size()) : -1, pn = ni ? fix(ni, N.size()) : -1;
                if (pv < 0 || pv >= (int)P.size() || (ti && (pt < 0 || pt >= (int)T.size())) || (ni && (pn < 0 || pn >= (int)N.size()))) { bad = true; break; }
                const JoinKey key{pv, pt, pn};
                auto it = g.join.find(key);
                if (it != g.join.end()) { poly.push_back(it->second); continue; }
                if (!ni) g.hasNormals = false;
                uint32_t id = g.mesh->addVertex(P[pv], ni ? N[pn] : Vec3f(0.0f), ti ? T[pt] : Vec2f{0, 0});
                g.join.emplace(key, id);
                poly.push_back(id);
            }
            if (bad) {
                const char* le = eol;
                while (le > lineStart && (le[-1] == '\r')) le--;
                std::fprintf(stderr, "[Model] %s: face with an index out of range skipped: %.*s\n", path.c_str(), (int)(le - lineStart), lineStart);
                continue;
            }
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

