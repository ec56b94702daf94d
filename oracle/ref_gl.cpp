// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref.
//
// ref_gl.cpp — "libGL" for the reference's own GL-object classes.  src/core/{Buffer,Texture,Image}.cpp are compiled
// unmodified, where they lie; the OpenGL 4.5 direct-state-access calls they make land here, on a store of host-memory
// objects.  Uploads copy and convert like a driver would for the formats the reference uses: RGB16F rounds binary32
// texels to binary16 (kept as binary32 values), integer and float formats are stored as they are, GL_SRGB arrays keep
// their 8-bit texels (decoded at sampling time), a NULL data pointer allocates zeroed storage.
#include "ref_gl.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include "../include/zl_libm.h"
#include "thirdparty/stb_image/stb_image.h"

namespace refgl {

static std::map<GLuint, Object>& store() { static std::map<GLuint, Object> s; return s; }
static GLuint nextName = 1;

Object* object(GLuint name) {
    auto it = store().find(name);
    return it == store().end() ? nullptr : &it->second;
}
static Object& need(GLuint name, const char* who) {
    Object* o = object(name);
    if (!o) { std::fprintf(stderr, "[refgl] %s: no object %u\n", who, name); std::abort(); }
    return *o;
}

float roundToHalf(float f) {
    uint32_t x = zl_f2u(f), sign = x & 0x80000000u, ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return f;
    if (ax >= 0x477ff000u) return zl_u2f(sign | 0x7f800000u);
    if (ax < 0x33000001u) return zl_u2f(sign);
    if (ax < 0x38800000u) {
        float q = zl_u2f(ax) * 16777216.0f;
        float r = std::nearbyint(q);
        return zl_u2f(sign | zl_f2u(r * (1.0f / 16777216.0f)));
    }
    uint32_t rem = ax & 0x1fffu, base = ax & ~0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (base & 0x2000u))) base += 0x2000u;
    return zl_u2f(sign | base);
}

bool formatInfo(GLenum f, int* comps, int* kind, bool* half) {
    *half = false;
    switch (f) {
    case GL_R32F: *comps = 1; *kind = K_FLOAT; return true;
    case GL_RG32F: *comps = 2; *kind = K_FLOAT; return true;
    case GL_RGB32F: *comps = 3; *kind = K_FLOAT; return true;
    case GL_RGBA32F: *comps = 4; *kind = K_FLOAT; return true;
    case GL_R16F: *comps = 1; *kind = K_FLOAT; *half = true; return true;
    case GL_RG16F: *comps = 2; *kind = K_FLOAT; *half = true; return true;
    case GL_RGB16F: *comps = 3; *kind = K_FLOAT; *half = true; return true;
    case GL_RGBA16F: *comps = 4; *kind = K_FLOAT; *half = true; return true;
    case GL_R32I: case GL_R32UI: *comps = 1; *kind = K_INT; return true;
    case GL_RG32I: case GL_RG32UI: *comps = 2; *kind = K_INT; return true;
    case GL_RGB32I: case GL_RGB32UI: *comps = 3; *kind = K_INT; return true;
    case GL_RGBA32I: case GL_RGBA32UI: *comps = 4; *kind = K_INT; return true;
    case GL_SRGB: *comps = 3; *kind = K_SRGB8; return true;
    default: return false;
    }
}

const uint8_t* texels(const Object& t, size_t* byteSize) {
    const Object* o = &t;
    if (t.target == GL_TEXTURE_BUFFER) o = &need(t.buffer, "texels(buffer texture)");
    if (byteSize) *byteSize = o->bytes.size();
    return o->bytes.empty() ? nullptr : o->bytes.data();
}

}  // namespace refgl
using namespace refgl;

void glCreateBuffers(GLsizei n, GLuint* ids) { for (int i = 0; i < n; i++) { ids[i] = nextName++; store()[ids[i]] = Object(); } }
void glDeleteBuffers(GLsizei n, const GLuint* ids) { for (int i = 0; i < n; i++) store().erase(ids[i]); }
void glNamedBufferData(GLuint b, GLsizeiptr size, const void* data, GLenum) {
    Object& o = need(b, "glNamedBufferData");
    o.bytes.assign((size_t)size, 0);
    if (data && size > 0) std::memcpy(o.bytes.data(), data, (size_t)size);
}
void glNamedBufferSubData(GLuint b, GLintptr off, GLsizeiptr size, const void* data) {
    Object& o = need(b, "glNamedBufferSubData");
    if ((size_t)(off + size) <= o.bytes.size()) std::memcpy(o.bytes.data() + off, data, (size_t)size);
}
void glGetNamedBufferSubData(GLuint b, GLintptr off, GLsizeiptr size, void* data) {
    Object& o = need(b, "glGetNamedBufferSubData");
    if ((size_t)(off + size) <= o.bytes.size()) std::memcpy(data, o.bytes.data() + off, (size_t)size);
}
void glCreateTextures(GLenum target, GLsizei n, GLuint* ids) {
    // src/core/Texture.cpp:27-31 creates the name in Texture::Texture and TextureBuffered creates a second one (:178): each is an object
    for (int i = 0; i < n; i++) { ids[i] = nextName++; Object o; o.target = target; store()[ids[i]] = o; }
}
void glDeleteTextures(GLsizei n, const GLuint* ids) { for (int i = 0; i < n; i++) store().erase(ids[i]); }
void glTextureParameteri(GLuint t, GLenum pname, GLint param) {
    Object& o = need(t, "glTextureParameteri");
    if (pname == GL_TEXTURE_MIN_FILTER) o.minFilter = param;
    else if (pname == GL_TEXTURE_MAG_FILTER) o.magFilter = param;
    else if (pname == GL_TEXTURE_WRAP_S) o.wrapS = param;
    else if (pname == GL_TEXTURE_WRAP_T) o.wrapT = param;
}
void glClearTexImage(GLuint t, GLint, GLenum, GLenum, const void* data) {
    Object& o = need(t, "glClearTexImage");
    if (data) { std::fprintf(stderr, "[refgl] glClearTexImage with a value is not implemented\n"); std::abort(); }
    std::fill(o.bytes.begin(), o.bytes.end(), 0);
}
void glTextureBuffer(GLuint t, GLenum internalformat, GLuint buffer) {
    Object& o = need(t, "glTextureBuffer");
    bool half;
    if (!formatInfo(internalformat, &o.comps, &o.kind, &half) || half) { std::fprintf(stderr, "[refgl] glTextureBuffer: format %u\n", internalformat); std::abort(); }
    o.target = GL_TEXTURE_BUFFER; o.internalFormat = internalformat; o.buffer = buffer;
    o.width = (int)(need(buffer, "glTextureBuffer").bytes.size() / (4 * (size_t)o.comps));
}
static int sourceComps(GLenum format) {
    switch (format) {
    case GL_RED: case GL_RED_INTEGER: return 1;
    case GL_RG: case GL_RG_INTEGER: return 2;
    case GL_RGB: case GL_RGB_INTEGER: return 3;
    case GL_RGBA: case GL_RGBA_INTEGER: return 4;
    default: return 0;
    }
}
void glTextureImage2DEXT(GLuint t, GLenum, GLint level, GLint internalformat, GLsizei w, GLsizei h, GLint, GLenum format, GLenum type, const void* pixels) {
    Object& o = need(t, "glTextureImage2DEXT");
    bool half;
    if (level != 0 || !formatInfo((GLenum)internalformat, &o.comps, &o.kind, &half)) { std::fprintf(stderr, "[refgl] glTextureImage2DEXT: format %d\n", internalformat); std::abort(); }
    o.internalFormat = (GLenum)internalformat; o.width = w; o.height = h; o.layers = 1;
    const size_t n = (size_t)w * h;
    o.bytes.assign(n * o.comps * 4, 0);
    if (!pixels) return;
    const int sc = sourceComps(format);
    if (sc < o.comps) { std::fprintf(stderr, "[refgl] glTextureImage2DEXT: %d source components for %d\n", sc, o.comps); std::abort(); }
    if (o.kind == K_FLOAT && type == GL_FLOAT) {
        const float* src = (const float*)pixels; float* dst = (float*)o.bytes.data();
        for (size_t i = 0; i < n; i++) for (int c = 0; c < o.comps; c++) { float v = src[i * sc + c]; dst[i * o.comps + c] = half ? roundToHalf(v) : v; }
    } else if (o.kind == K_INT && (type == GL_INT || type == GL_UNSIGNED_INT)) {
        const int32_t* src = (const int32_t*)pixels; int32_t* dst = (int32_t*)o.bytes.data();
        for (size_t i = 0; i < n; i++) for (int c = 0; c < o.comps; c++) dst[i * o.comps + c] = src[i * sc + c];
    } else { std::fprintf(stderr, "[refgl] glTextureImage2DEXT: source type %u for kind %d\n", type, o.kind); std::abort(); }
}
void glTextureImage3DEXT(GLuint t, GLenum, GLint, GLint internalformat, GLsizei w, GLsizei h, GLsizei d, GLint, GLenum, GLenum, const void* pixels) {
    Object& o = need(t, "glTextureImage3DEXT");
    bool half;
    if (!formatInfo((GLenum)internalformat, &o.comps, &o.kind, &half) || o.kind != K_SRGB8 || pixels) { std::fprintf(stderr, "[refgl] glTextureImage3DEXT: only an empty GL_SRGB array\n"); std::abort(); }
    o.internalFormat = (GLenum)internalformat; o.width = w; o.height = h; o.layers = d;
    o.bytes.assign((size_t)w * h * d * 3, 0);
}
void glTextureSubImage3D(GLuint t, GLint, GLint x0, GLint y0, GLint z0, GLsizei w, GLsizei h, GLsizei d, GLenum format, GLenum type, const void* pixels) {
    Object& o = need(t, "glTextureSubImage3D");
    if (o.kind != K_SRGB8 || format != GL_RGB || type != GL_UNSIGNED_BYTE || d != 1) { std::fprintf(stderr, "[refgl] glTextureSubImage3D: RGB8 layers only\n"); std::abort(); }
    const uint8_t* src = (const uint8_t*)pixels;
    for (int y = 0; y < h; y++)        // GL_UNPACK_ALIGNMENT is left at 4 by the reference; the stand-in images have w * 3 % 4 == 0 or are tightly packed
        std::memcpy(&o.bytes[(((size_t)z0 * o.height + y0 + y) * o.width + x0) * 3], src + (size_t)y * w * 3, (size_t)w * 3);
}
void glGetTextureImage(GLuint t, GLint, GLenum format, GLenum type, GLsizei bufSize, void* pixels) {
    Object& o = need(t, "glGetTextureImage");
    if (o.kind != K_FLOAT || format != GL_RGB || type != GL_UNSIGNED_BYTE) { std::fprintf(stderr, "[refgl] glGetTextureImage: float -> RGB8 only\n"); std::abort(); }
    const float* src = (const float*)o.bytes.data(); uint8_t* dst = (uint8_t*)pixels;
    const size_t n = (size_t)o.width * o.height;
    if ((size_t)bufSize < n * 3) return;
    for (size_t i = 0; i < n; i++) for (int c = 0; c < 3; c++) {
        float v = c < o.comps ? src[i * o.comps + c] : 0.0f;
        v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        dst[i * 3 + c] = (uint8_t)std::nearbyint(v * 255.0f);
    }
}

// ---- stb_image stand-in over registered in-memory images ----
namespace {
struct MemImage { int w, h, ch; std::vector<float> f; std::vector<uint8_t> b; };
std::map<std::string, MemImage>& images() { static std::map<std::string, MemImage> m; return m; }
}
extern "C" void zr_register_image(const char* path, int w, int h, int channels, const float* dataF, const uint8_t* data8) {
    MemImage im; im.w = w; im.h = h; im.ch = channels;
    const size_t n = (size_t)w * h * channels;
    if (dataF) im.f.assign(dataF, dataF + n);
    if (data8) im.b.assign(data8, data8 + n);
    images()[path] = std::move(im);
}
extern "C" void zr_clear_images(void) { images().clear(); }
unsigned char* stbi_load(const char* path, int* w, int* h, int* ch, int desired) {
    auto it = images().find(path);
    if (it == images().end() || it->second.b.empty() || it->second.ch != desired) return nullptr;
    *w = it->second.w; *h = it->second.h; *ch = it->second.ch;
    unsigned char* p = (unsigned char*)std::malloc(it->second.b.size());
    std::memcpy(p, it->second.b.data(), it->second.b.size());
    return p;
}
float* stbi_loadf(const char* path, int* w, int* h, int* ch, int desired) {
    auto it = images().find(path);
    if (it == images().end() || it->second.f.empty() || it->second.ch != desired) return nullptr;
    *w = it->second.w; *h = it->second.h; *ch = it->second.ch;
    float* p = (float*)std::malloc(it->second.f.size() * sizeof(float));
    std::memcpy(p, it->second.f.data(), it->second.f.size() * sizeof(float));
    return p;
}
void stbi_image_free(void* p) { std::free(p); }
