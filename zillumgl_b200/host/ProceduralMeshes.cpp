// "builtin:" procedural models.  The reference ships no assets (res/ holds only
// scene.xml), so the benchmark scenes are generated: the default scene's square / teapot,
// a Cornell box, a Sponza-class atrium with exactly 262,144 triangles and a Rungholt-class
// voxel city with exactly 6,291,456 triangles.  Generators are
// deterministic (fixed integer hashes, no libc rand).  Model space is Y-up like the
// reference's imported assets; ModelInstance::modelMatrix() turns it Z-up.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <sstream>
#include "Model.h"

namespace zillum {

static const float kPi = 3.14159265358979323846f;

static uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
static float hash01(uint32_t a, uint32_t b = 0) { return (mix32(a * 0x9e3779b9u + mix32(b)) >> 8) * (1.0f / 16777216.0f); }

// Builders think in the Z-up world frame; Z(...) converts to model space.
static Vec3f Z(float x, float y, float z) { return MeshData::fromZUp(x, y, z); }
static Vec3f Zv(Vec3f v) { return MeshData::fromZUp(v.x, v.y, v.z); }

// (nu x nv) quad grid over a parametric surface P(u,v), u,v in [0,1]; normal from N(u,v).
static void addSurface(MeshData& m, int nu, int nv, const std::function<Vec3f(float, float)>& P,
                       const std::function<Vec3f(float, float)>& N, float uvScaleU = 1.0f, float uvScaleV = 1.0f,
                       bool flip = false) {
    uint32_t base = (uint32_t)m.positions.size();
    for (int j = 0; j <= nv; j++)
        for (int i = 0; i <= nu; i++) {
            float u = (float)i / nu, v = (float)j / nv;
            m.addVertex(Zv(P(u, v)), Zv(N(u, v)), Vec2f{u * uvScaleU, v * uvScaleV});
        }
    for (int j = 0; j < nv; j++)
        for (int i = 0; i < nu; i++) {
            uint32_t a = base + j * (nu + 1) + i, b = a + 1, c = a + (nu + 1), d = c + 1;
            if (!flip) { m.addTriangle(a, b, d); m.addTriangle(a, d, c); }
            else { m.addTriangle(a, d, b); m.addTriangle(a, c, d); }
        }
}

static void addQuadGrid(MeshData& m, Vec3f o, Vec3f du, Vec3f dv, int nu, int nv, float uvU = 1.0f, float uvV = 1.0f) {
    Vec3f n = normalize(cross(du, dv));
    addSurface(m, nu, nv, [=](float u, float v) { return o + du * u + dv * v; }, [=](float, float) { return n; }, uvU, uvV);
}

// axis-aligned box with flat normals (24 vertices, 12 triangles), world frame
static void addBox(MeshData& m, Vec3f lo, Vec3f hi) {
    Vec3f d = hi - lo;
    addQuadGrid(m, Vec3f(lo.x, lo.y, hi.z), Vec3f(d.x, 0, 0), Vec3f(0, d.y, 0), 1, 1);   // +z
    addQuadGrid(m, Vec3f(lo.x, hi.y, lo.z), Vec3f(d.x, 0, 0), Vec3f(0, -d.y, 0), 1, 1);  // -z
    addQuadGrid(m, Vec3f(hi.x, lo.y, lo.z), Vec3f(0, d.y, 0), Vec3f(0, 0, d.z), 1, 1);   // +x
    addQuadGrid(m, Vec3f(lo.x, hi.y, lo.z), Vec3f(0, -d.y, 0), Vec3f(0, 0, d.z), 1, 1);  // -x
    addQuadGrid(m, Vec3f(hi.x, hi.y, lo.z), Vec3f(-d.x, 0, 0), Vec3f(0, 0, d.z), 1, 1);  // +y
    addQuadGrid(m, Vec3f(lo.x, lo.y, lo.z), Vec3f(d.x, 0, 0), Vec3f(0, 0, d.z), 1, 1);   // -y
}

// surface of revolution about the world Z axis through (cx, cy); profile(t) -> (radius, height)
static void addLathe(MeshData& m, float cx, float cy, int segs, int rings, const std::function<Vec2f(float)>& profile) {
    auto P = [=](float u, float v) {
        Vec2f p = profile(v);
        float a = u * 2.0f * kPi;
        return Vec3f(cx + p.x * std::cos(a), cy + p.x * std::sin(a), p.y);
    };
    auto N = [=](float u, float v) {
        const float e = 1e-3f;
        Vec2f p0 = profile(std::fmax(v - e, 0.0f)), p1 = profile(std::fmin(v + e, 1.0f));
        float dr = p1.x - p0.x, dz = p1.y - p0.y;
        float a = u * 2.0f * kPi;
        Vec3f n(dz * std::cos(a), dz * std::sin(a), -dr);
        float l = length(n);
        return l > 0 ? n / l : Vec3f(0, 0, 1);
    };
    addSurface(m, segs, rings, P, N, 4.0f, 2.0f);
}

static MeshInstancePtr makeInstance(MeshDataPtr mesh, int matIndex, int texIndex = -1) {
    auto mi = std::make_shared<MeshInstance>();
    mi->meshData = mesh;
    mi->matIndex = matIndex;
    mi->texIndex = texIndex;
    return mi;
}

static std::map<std::string, std::string> parseQuery(const std::string& spec, std::string& name) {
    std::map<std::string, std::string> kv;
    size_t q = spec.find('?');
    name = spec.substr(0, q);
    if (q == std::string::npos) return kv;
    std::stringstream ss(spec.substr(q + 1));
    std::string item;
    while (std::getline(ss, item, '&')) {
        size_t e = item.find('=');
        if (e != std::string::npos) kv[item.substr(0, e)] = item.substr(e + 1);
    }
    return kv;
}

static size_t triCount(const ModelInstancePtr& m) {
    size_t n = 0;
    for (auto& mi : m->meshInstances()) n += mi->meshData->indices.size() / 3;
    return n;
}

// ---- small models -------------------------------------------------------------------------
static ModelInstancePtr makeSquare() {         // unit square, normal +Y (model) = +Z (world)
    auto mesh = std::make_shared<MeshData>();
    addQuadGrid(*mesh, Vec3f(-0.5f, -0.5f, 0), Vec3f(1, 0, 0), Vec3f(0, 1, 0), 1, 1);
    auto model = std::make_shared<ModelInstance>();
    model->meshInstances().push_back(makeInstance(mesh, 0));
    model->materials().push_back(Material());
    return model;
}

static ModelInstancePtr makeCube() {
    auto mesh = std::make_shared<MeshData>();
    addBox(*mesh, Vec3f(-0.5f, -0.5f, -0.5f), Vec3f(0.5f, 0.5f, 0.5f));
    auto model = std::make_shared<ModelInstance>();
    model->meshInstances().push_back(makeInstance(mesh, 0));
    model->materials().push_back(Material());
    return model;
}

static ModelInstancePtr makeSphere(int segs, int rings) {
    auto mesh = std::make_shared<MeshData>();
    addSurface(*mesh, segs, rings,
        [](float u, float v) { float a = u * 2 * kPi, b = v * kPi; return Vec3f(std::sin(b) * std::cos(a), std::sin(b) * std::sin(a), -std::cos(b)); },
        [](float u, float v) { float a = u * 2 * kPi, b = v * kPi; return Vec3f(std::sin(b) * std::cos(a), std::sin(b) * std::sin(a), -std::cos(b)); }, 2.0f, 1.0f);
    auto model = std::make_shared<ModelInstance>();
    model->meshInstances().push_back(makeInstance(mesh, 0));
    model->materials().push_back(Material());
    return model;
}

// teapot-class body / lid of the default scene (res/scene.xml:24-39 uses teapot/Mesh003, Mesh004)
static ModelInstancePtr makeTeapotPart(bool cap) {
    auto mesh = std::make_shared<MeshData>();
    if (!cap) {
        addLathe(*mesh, 0, 0, 64, 40, [](float t) {
            // rounded pot: foot, belly, shoulder, rim
            float z = 2.2f * t;
            float r = 0.55f + 1.05f * std::sin(kPi * std::pow(t, 0.8f)) + 0.25f * (1.0f - t);
            if (t > 0.92f) r = 0.55f + 1.05f * std::sin(kPi * std::pow(0.92f, 0.8f)) + 0.02f + (t - 0.92f) * 0.6f;
            return Vec2f{r, z};
        });
    } else {
        addLathe(*mesh, 0, 0, 64, 10, [](float t) {
            float r = 0.9f * std::cos(t * kPi * 0.5f) + (t > 0.8f ? 0.18f * std::sin((t - 0.8f) * 5.0f * kPi) : 0.0f);
            return Vec2f{std::fmax(r, 0.0f), 2.25f + 0.55f * t};
        });
    }
    auto model = std::make_shared<ModelInstance>();
    model->meshInstances().push_back(makeInstance(mesh, 0));
    model->materials().push_back(Material());
    return model;
}

// Cornell box room: floor/ceiling/back wall white, left red, right green, a tall Principled
// block and a short rough-dielectric block.  Interior spans [-1,1]^2 x [0,2].
static ModelInstancePtr makeCornell() {
    auto model = std::make_shared<ModelInstance>();
    Material white; white.baseColor = Vec3f(0.73f, 0.73f, 0.73f);
    Material red; red.baseColor = Vec3f(0.65f, 0.05f, 0.05f);
    Material green; green.baseColor = Vec3f(0.12f, 0.45f, 0.15f);
    Material tall; tall.type = Material::Principled; tall.baseColor = Vec3f(0.8f, 0.75f, 0.6f); tall.roughness = 0.35f;
    tall.metallic = 0.2f; tall.specular = 0.8f; tall.clearcoat = 0.6f; tall.clearcoatGloss = 0.8f; tall.sheen = 0.1f;
    Material glass; glass.type = Material::Dielectric; glass.baseColor = Vec3f(0.95f, 0.97f, 1.0f); glass.ior = 1.5f; glass.roughness = 0.15f;
    model->materials() = {white, red, green, tall, glass};
    auto wWhite = std::make_shared<MeshData>(), wRed = std::make_shared<MeshData>(), wGreen = std::make_shared<MeshData>();
    auto bTall = std::make_shared<MeshData>(), bShort = std::make_shared<MeshData>();
    addQuadGrid(*wWhite, Vec3f(-1, -1, 0), Vec3f(2, 0, 0), Vec3f(0, 2, 0), 1, 1);    // floor (normal +z)
    addQuadGrid(*wWhite, Vec3f(-1, 1, 2), Vec3f(2, 0, 0), Vec3f(0, -2, 0), 1, 1);    // ceiling (normal -z)
    addQuadGrid(*wWhite, Vec3f(-1, 1, 0), Vec3f(2, 0, 0), Vec3f(0, 0, 2), 1, 1);     // back wall at y=+1 (normal -y)
    addQuadGrid(*wRed, Vec3f(-1, -1, 0), Vec3f(0, 2, 0), Vec3f(0, 0, 2), 1, 1);      // left wall x=-1 (normal +x)
    addQuadGrid(*wGreen, Vec3f(1, 1, 0), Vec3f(0, -2, 0), Vec3f(0, 0, 2), 1, 1);     // right wall x=+1 (normal -x)
    auto rotBox = [](MeshData& m, Vec3f c, Vec3f half, float angDeg) {
        size_t first = m.positions.size();
        addBox(m, Vec3f(-half.x, -half.y, 0), Vec3f(half.x, half.y, 2 * half.z));
        float a = angDeg * kPi / 180.0f, cs = std::cos(a), sn = std::sin(a);
        for (size_t i = first; i < m.positions.size(); i++) {
            // vertices are stored in model space (x, z, -y): rotate about world Z
            Vec3f p = m.positions[i], n = m.normals[i];
            float wx = p.x, wy = -p.z, nx = n.x, ny = -n.z;
            float rx = wx * cs - wy * sn + c.x, ry = wx * sn + wy * cs + c.y;
            float rnx = nx * cs - ny * sn, rny = nx * sn + ny * cs;
            m.positions[i] = Vec3f(rx, p.y + c.z, -ry);
            m.normals[i] = Vec3f(rnx, n.y, -rny);
        }
    };
    rotBox(*bTall, Vec3f(-0.35f, 0.3f, 0), Vec3f(0.3f, 0.3f, 0.6f), 18.0f);
    rotBox(*bShort, Vec3f(0.35f, -0.3f, 0), Vec3f(0.3f, 0.3f, 0.3f), -17.0f);
    model->meshInstances() = {makeInstance(wWhite, 0), makeInstance(wRed, 1), makeInstance(wGreen, 2), makeInstance(bTall, 3), makeInstance(bShort, 4)};
    return model;
}

// ---- Sponza-class atrium: exactly 262,144 triangles ---------------------------------------
static ByteImagePtr makeCheckerTexture(int size, int cells) {
    auto img = std::make_shared<ByteImage>();
    img->width = img->height = size;
    img->rgb.resize((size_t)size * size * 3);
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            int cx = x * cells / size, cy = y * cells / size;
            float n = hash01((uint32_t)(x / 4), (uint32_t)(y / 4)) * 0.12f;
            float base = ((cx + cy) & 1) ? 0.78f : 0.42f;
            unsigned char* p = &img->rgb[3 * ((size_t)y * size + x)];
            p[0] = (unsigned char)(255 * std::fmin(base + n, 1.0f));
            p[1] = (unsigned char)(255 * std::fmin(base * 0.93f + n, 1.0f));
            p[2] = (unsigned char)(255 * std::fmin(base * 0.82f + n, 1.0f));
        }
    return img;
}

static ModelInstancePtr makeSponza() {
    const size_t kTarget = 262144;
    auto model = std::make_shared<ModelInstance>();
    enum { MFloor, MWall, MColumn, MArch, MGallery, MDrapeR, MDrapeG, MDrapeB, MGold, MGlass, MFrieze, MCount };
    std::vector<Material> mats(MCount);
    mats[MFloor].type = Material::Principled; mats[MFloor].roughness = 0.45f; mats[MFloor].specular = 0.6f; mats[MFloor].baseColor = Vec3f(0.7f);
    mats[MWall].baseColor = Vec3f(0.72f, 0.66f, 0.55f);
    mats[MColumn].type = Material::Principled; mats[MColumn].baseColor = Vec3f(0.75f, 0.72f, 0.66f); mats[MColumn].roughness = 0.6f; mats[MColumn].specular = 0.4f;
    mats[MArch].baseColor = Vec3f(0.68f, 0.62f, 0.52f);
    mats[MGallery].baseColor = Vec3f(0.55f, 0.5f, 0.45f);
    Vec3f drape[3] = {Vec3f(0.7f, 0.08f, 0.08f), Vec3f(0.1f, 0.5f, 0.15f), Vec3f(0.1f, 0.15f, 0.6f)};
    for (int i = 0; i < 3; i++) {
        Material& d = mats[MDrapeR + i];
        d.type = Material::Principled; d.baseColor = drape[i]; d.roughness = 0.8f; d.sheen = 0.8f; d.sheenTint = 0.5f; d.specular = 0.2f; d.subsurface = 0.3f;
    }
    mats[MGold].type = Material::MetalWorkflow; mats[MGold].baseColor = Vec3f(1.0f, 0.77f, 0.34f); mats[MGold].metallic = 1.0f; mats[MGold].roughness = 0.25f;
    mats[MGlass].type = Material::Dielectric; mats[MGlass].baseColor = Vec3f(0.95f, 1.0f, 0.97f); mats[MGlass].ior = 1.5f; mats[MGlass].roughness = 0.0f;
    mats[MFrieze].baseColor = Vec3f(0.6f, 0.55f, 0.5f);
    model->materials() = mats;
    std::vector<MeshDataPtr> mesh(MCount);
    for (auto& m : mesh) m = std::make_shared<MeshData>();

    const float L = 20.0f, W = 8.0f, H = 12.0f;            // half-length, half-width, height
    // floor (textured) 128x64 quads
    addQuadGrid(*mesh[MFloor], Vec3f(-L, -W, 0), Vec3f(2 * L, 0, 0), Vec3f(0, 2 * W, 0), 128, 64, 20.0f, 8.0f);
    // four walls, slightly bumpy so the tessellation matters
    auto wall = [&](Vec3f o, Vec3f du, Vec3f dv, int nu, int nv, uint32_t seed) {
        Vec3f n = normalize(cross(du, dv));
        addSurface(*mesh[MWall], nu, nv,
            [=](float u, float v) { float b = 0.03f * std::sin(u * 97.0f + seed) * std::sin(v * 41.0f); return o + du * u + dv * v + n * b; },
            [=](float, float) { return n; }, 10.0f, 4.0f);
    };
    wall(Vec3f(-L, W, 0), Vec3f(2 * L, 0, 0), Vec3f(0, 0, H), 128, 32, 1);      // y=+W, normal -y
    wall(Vec3f(L, -W, 0), Vec3f(-2 * L, 0, 0), Vec3f(0, 0, H), 128, 32, 2);     // y=-W, normal +y
    wall(Vec3f(L, W, 0), Vec3f(0, -2 * W, 0), Vec3f(0, 0, H), 64, 32, 3);       // x=+L, normal -x
    wall(Vec3f(-L, -W, 0), Vec3f(0, 2 * W, 0), Vec3f(0, 0, H), 64, 32, 4);      // x=-L, normal +x
    // columns: 2 rows x 10 x 2 levels, fluted
    const int nCols = 10;
    auto colX = [&](int i) { return -L + 2.0f + i * (2 * L - 4.0f) / (nCols - 1); };
    for (int level = 0; level < 2; level++)
        for (int row = 0; row < 2; row++)
            for (int i = 0; i < nCols; i++) {
                float cx = colX(i), cy = (row ? 1.0f : -1.0f) * 4.5f, z0 = level * 6.0f, z1 = z0 + 4.2f;
                addLathe(*mesh[MColumn], cx, cy, 32, 16, [=](float t) {
                    float r = 0.32f + 0.10f * std::exp(-30.0f * t) + 0.12f * std::exp(-30.0f * (1 - t));
                    return Vec2f{r, z0 + (z1 - z0) * t};
                });
                // caps: fans as degenerate lathe rings (32 tris each)
                for (int capI = 0; capI < 2; capI++) {
                    MeshData& m = *mesh[MColumn];
                    float z = capI ? z1 : z0, r = 0.32f + 0.22f - (capI ? 0.10f : 0.12f) + 0.0f;
                    Vec3f n = capI ? Vec3f(0, 0, 1) : Vec3f(0, 0, -1);
                    uint32_t c = m.addVertex(Z(cx, cy, z), Zv(n), Vec2f{0.5f, 0.5f});
                    uint32_t first = (uint32_t)m.positions.size();
                    for (int s = 0; s < 32; s++) { float a = s * 2 * kPi / 32; m.addVertex(Z(cx + r * std::cos(a), cy + r * std::sin(a), z), Zv(n), Vec2f{std::cos(a), std::sin(a)}); }
                    for (int s = 0; s < 32; s++) { uint32_t a = first + s, b = first + (s + 1) % 32; if (capI) m.addTriangle(c, a, b); else m.addTriangle(c, b, a); }
                }
            }
    // arches between neighbouring columns: half torus tubes, 24 x 16 quads
    for (int level = 0; level < 2; level++)
        for (int row = 0; row < 2; row++)
            for (int i = 0; i + 1 < nCols; i++) {
                float x0 = colX(i), x1 = colX(i + 1), cy = (row ? 1.0f : -1.0f) * 4.5f, zc = level * 6.0f + 4.2f;
                float R = 0.5f * (x1 - x0), xm = 0.5f * (x0 + x1), r = 0.22f;
                addSurface(*mesh[MArch], 24, 16,
                    [=](float u, float v) { float a = u * kPi, b = v * 2 * kPi; float rr = R + r * std::cos(b); return Vec3f(xm - rr * std::cos(a), cy + r * std::sin(b), zc + rr * std::sin(a)); },
                    [=](float u, float v) { float a = u * kPi, b = v * 2 * kPi; return Vec3f(-std::cos(b) * std::cos(a), std::sin(b), std::cos(b) * std::sin(a)); }, 4.0f, 1.0f, true);
            }
    // gallery walkways (top + bottom faces), both sides
    for (int row = 0; row < 2; row++) {
        float y0 = row ? 4.0f : -W, y1 = row ? W : -4.0f;
        addQuadGrid(*mesh[MGallery], Vec3f(-L, y0, 6.0f), Vec3f(2 * L, 0, 0), Vec3f(0, y1 - y0, 0), 128, 8, 20.0f, 2.0f);
        addQuadGrid(*mesh[MGallery], Vec3f(-L, y1, 5.7f), Vec3f(2 * L, 0, 0), Vec3f(0, y0 - y1, 0), 128, 8, 20.0f, 2.0f);
    }
    // drapes: 12 hanging wavy sheets, 64x64 quads each
    for (int d = 0; d < 12; d++) {
        float cx = -L + 3.5f + (d % 6) * 6.6f, cy = (d < 6 ? -1.0f : 1.0f) * 3.6f;
        uint32_t seed = 17 + d * 31;
        float ph = hash01(seed) * 6.28f;
        auto P = [=](float u, float v) {
            float x = cx + (u - 0.5f) * 3.0f;
            float sway = 0.25f * std::sin(u * 18.0f + ph) * (0.3f + 0.7f * (1 - v)) + 0.08f * std::sin(v * 9.0f + u * 5.0f);
            return Vec3f(x, cy + sway, 5.5f + 5.0f * v);
        };
        auto Nf = [=](float u, float v) {
            const float e = 1e-3f;
            Vec3f du = P(std::fmin(u + e, 1.0f), v) - P(std::fmax(u - e, 0.0f), v), dv = P(u, std::fmin(v + e, 1.0f)) - P(u, std::fmax(v - e, 0.0f));
            Vec3f n = cross(du, dv); float l = length(n);
            return l > 0 ? n / l : Vec3f(0, 1, 0);
        };
        addSurface(*mesh[MDrapeR + d % 3], 64, 64, P, Nf, 2.0f, 2.0f);
    }
    // vases: 8 lathes (gold / glass alternating), 64 x 32 quads
    for (int i = 0; i < 8; i++) {
        float cx = -L + 5.0f + i * 4.3f, cy = (i & 1) ? 1.6f : -1.6f;
        addLathe(*mesh[(i & 1) ? MGlass : MGold], cx, cy, 64, 32, [](float t) {
            float r = 0.18f + 0.45f * std::sin(kPi * std::pow(t, 0.7f)) * (1.0f - 0.35f * t) + 0.1f * (t > 0.9f ? (t - 0.9f) * 8.0f : 0.0f);
            return Vec2f{r, 1.6f * t};
        });
    }
    // frieze band: whatever is left, as an (n x 64) bumpy strip along the +y wall
    size_t used = 0;
    for (auto& m : mesh) used += m->indices.size() / 3;
    if (used > kTarget || (kTarget - used) % 128 != 0) { std::fprintf(stderr, "[sponza] generator budget broken: %zu\n", used); std::abort(); }
    int nu = (int)((kTarget - used) / 128);
    if (nu > 0) {
        Vec3f o(-L, W - 0.15f, H - 1.5f), du(2 * L, 0, 0), dv(0, 0, 1.2f), n(0, -1, 0);
        addSurface(*mesh[MFrieze], nu, 64,
            [=](float u, float v) { float b = 0.06f * std::sin(u * 240.0f) * std::sin(v * 25.0f); return o + du * u + dv * v + n * b; },
            [=](float u, float v) { float gx = 0.06f * 240.0f * std::cos(u * 240.0f) * std::sin(v * 25.0f) / 40.0f; Vec3f m2(gx, -1, 0); return m2 / length(m2); }, 40.0f, 1.0f, true);
    }
    int checker = Resource::addImage(makeCheckerTexture(256, 8), "builtin:checker256");
    for (int i = 0; i < MCount; i++)
        if (!mesh[i]->indices.empty()) model->meshInstances().push_back(makeInstance(mesh[i], i, i == MFloor ? checker : -1));
    if (triCount(model) != kTarget) { std::fprintf(stderr, "[sponza] %zu triangles, expected %zu\n", triCount(model), kTarget); std::abort(); }
    return model;
}

// ---- Rungholt-class voxel city: nx*ny boxes x 12 triangles (default 1024x512 = 6,291,456) ----
static ModelInstancePtr makeRungholt(int nx, int ny) {
    auto model = std::make_shared<ModelInstance>();
    const int nMat = 8;
    std::vector<Material> mats(nMat);
    Vec3f palette[6] = {Vec3f(0.35f, 0.55f, 0.25f), Vec3f(0.6f, 0.55f, 0.45f), Vec3f(0.7f, 0.7f, 0.72f), Vec3f(0.55f, 0.3f, 0.22f), Vec3f(0.85f, 0.82f, 0.7f), Vec3f(0.3f, 0.32f, 0.38f)};
    for (int i = 0; i < 6; i++) mats[i].baseColor = palette[i];
    mats[6].type = Material::Principled; mats[6].baseColor = Vec3f(0.2f, 0.35f, 0.6f); mats[6].roughness = 0.2f; mats[6].specular = 1.0f; mats[6].clearcoat = 1.0f; mats[6].clearcoatGloss = 0.9f;
    mats[7].type = Material::MetalWorkflow; mats[7].baseColor = Vec3f(0.9f, 0.9f, 0.92f); mats[7].metallic = 1.0f; mats[7].roughness = 0.3f;
    model->materials() = mats;
    std::vector<MeshDataPtr> mesh(nMat);
    for (auto& m : mesh) { m = std::make_shared<MeshData>(); }
    // fractal value noise for the terrain + city blocks with towers
    auto vnoise = [](float x, float y, uint32_t seed) {
        int xi = (int)std::floor(x), yi = (int)std::floor(y);
        float fx = x - xi, fy = y - yi;
        fx = fx * fx * (3 - 2 * fx); fy = fy * fy * (3 - 2 * fy);
        float a = hash01((uint32_t)xi * 73856093u ^ (uint32_t)yi * 19349663u, seed), b = hash01((uint32_t)(xi + 1) * 73856093u ^ (uint32_t)yi * 19349663u, seed);
        float c = hash01((uint32_t)xi * 73856093u ^ (uint32_t)(yi + 1) * 19349663u, seed), d = hash01((uint32_t)(xi + 1) * 73856093u ^ (uint32_t)(yi + 1) * 19349663u, seed);
        return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy;
    };
    // rows are generated in chunks, in parallel, each chunk into its own eight meshes; the chunks are then concatenated in row order, so the
    // meshes are the ones a single j, i loop appends (half a million boxes = 12.6 M vertices: 0.7 s on one core)
    const int chunks = std::max(1, std::min(ny, 64));
    std::vector<std::vector<MeshData>> part(chunks, std::vector<MeshData>(nMat));
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < chunks; c++)
    for (int j = (int)((long)c * ny / chunks); j < (int)((long)(c + 1) * ny / chunks); j++)
        for (int i = 0; i < nx; i++) {
            float x = (float)i, y = (float)j;
            float terrain = 6.0f * vnoise(x / 96.0f, y / 96.0f, 1) + 3.0f * vnoise(x / 31.0f, y / 31.0f, 2) + 1.0f * vnoise(x / 9.0f, y / 9.0f, 3);
            int bx = i / 12, by = j / 12;                    // city blocks of 12x12 cells with 2-cell streets
            bool street = (i % 12) < 2 || (j % 12) < 2;
            float urban = vnoise(bx / 6.0f, by / 6.0f, 7);
            float h = std::floor(terrain) + 1.0f;
            int mat = (int)(hash01((uint32_t)bx * 9176u + by, 11) * 3.999f);
            if (urban > 0.45f) {
                if (street) { h = std::floor(terrain * 0.3f) + 1.0f; mat = 5; }
                else {
                    float tower = hash01((uint32_t)bx * 7919u + by, 5);
                    h = std::floor(terrain * 0.3f) + 2.0f + std::floor(tower * tower * 40.0f * (urban - 0.3f));
                    int lot = ((i % 12) / 5) + 2 * ((j % 12) / 5);
                    h = std::fmax(h - std::floor(hash01((uint32_t)(bx * 131 + by) * 4 + lot, 9) * 6.0f), 2.0f);
                    float r = hash01((uint32_t)bx * 31337u + by, 13);
                    mat = r < 0.12f ? 6 : (r < 0.2f ? 7 : 1 + (int)(r * 4.999f) % 4);
                }
            } else if (h < 3.0f) mat = 5; else if (h > 8.0f) mat = 2; else mat = 0;
            addBox(part[c][mat], Vec3f(x - nx * 0.5f, y - ny * 0.5f, 0.0f), Vec3f(x + 1 - nx * 0.5f, y + 1 - ny * 0.5f, h));
        }
    for (int m = 0; m < nMat; m++) {
        std::vector<size_t> vOff(chunks + 1, 0), iOff(chunks + 1, 0);
        for (int c = 0; c < chunks; c++) { vOff[c + 1] = vOff[c] + part[c][m].positions.size(); iOff[c + 1] = iOff[c] + part[c][m].indices.size(); }
        MeshData& out = *mesh[m];
        out.positions.resize(vOff[chunks]); out.normals.resize(vOff[chunks]); out.texcoords.resize(vOff[chunks]); out.indices.resize(iOff[chunks]);
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < chunks; c++) {
            MeshData& in = part[c][m];
            std::copy(in.positions.begin(), in.positions.end(), out.positions.begin() + vOff[c]);
            std::copy(in.normals.begin(), in.normals.end(), out.normals.begin() + vOff[c]);
            std::copy(in.texcoords.begin(), in.texcoords.end(), out.texcoords.begin() + vOff[c]);
            const uint32_t base = (uint32_t)vOff[c];
            for (size_t k = 0; k < in.indices.size(); k++) out.indices[iOff[c] + k] = in.indices[k] + base;
            in = MeshData();
        }
    }
    for (int i = 0; i < nMat; i++)
        if (!mesh[i]->indices.empty()) model->meshInstances().push_back(makeInstance(mesh[i], i));
    return model;
}

ModelInstancePtr makeBuiltinModel(const std::string& spec) {
    std::string name;
    auto kv = parseQuery(spec, name);
    auto geti = [&](const char* k, int def) { auto it = kv.find(k); return it == kv.end() ? def : std::atoi(it->second.c_str()); };
    if (name == "square") return makeSquare();
    if (name == "cube") return makeCube();
    if (name == "sphere") return makeSphere(geti("segs", 64), geti("rings", 32));
    if (name == "teapotBody") return makeTeapotPart(false);
    if (name == "teapotCap") return makeTeapotPart(true);
    if (name == "cornell") return makeCornell();
    if (name == "sponza") return makeSponza();
    if (name == "rungholt") return makeRungholt(geti("nx", 1024), geti("ny", 512));
    std::fprintf(stderr, "[Model] unknown builtin model '%s'\n", spec.c_str());
    return nullptr;
}

}  // namespace zillum
