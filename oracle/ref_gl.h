// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref.
// ref_gl.h — the host-memory object store behind the GL entry points of ref_shim/include/glad/glad.h.
#pragma once
#include <cstdint>
#include <vector>
#include <glad/glad.h>

namespace refgl {

enum Kind { K_FLOAT = 0, K_INT = 1, K_SRGB8 = 2 };        // same numbering as glsl::TexBinding::kind

struct Object {
    GLenum target = 0;                 // 0 = buffer, else GL_TEXTURE_*
    std::vector<uint8_t> bytes;        // buffers and 2-D / array textures own their storage
    GLuint buffer = 0;                 // buffer textures: the buffer whose bytes they view
    GLenum internalFormat = 0;
    int comps = 0, kind = K_FLOAT, width = 0, height = 1, layers = 1;
    GLint minFilter = GL_LINEAR, magFilter = GL_LINEAR, wrapS = GL_REPEAT, wrapT = GL_REPEAT;
};

Object* object(GLuint name);           // nullptr when the name is not alive
// storage of a texture as the samplers see it: buffer textures resolve to their buffer
const uint8_t* texels(const Object& t, size_t* byteSize);
bool formatInfo(GLenum internalFormat, int* comps, int* kind, bool* half);
float roundToHalf(float f);

}  // namespace refgl

extern "C" {
// in-memory "files" for stbi_load / stbi_loadf (src/core/Image.cpp:12-19): float RGB or 8-bit RGB, row 0 first
void zr_register_image(const char* path, int w, int h, int channels, const float* dataF, const uint8_t* data8);
void zr_clear_images(void);
}
