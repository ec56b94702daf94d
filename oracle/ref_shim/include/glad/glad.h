// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref.
// glad/glad.h — stand-in for the OpenGL loader header the reference includes everywhere.  No GL exists in this
// image: the reference's GL-object classes (Buffer, Texture*, Shader, Pipeline) are given host-memory
// implementations in oracle/ref_host.cpp instead of their GL ones, so only the TYPES and the enumerant NAMES the
// reference's headers mention are needed.  The values are arbitrary distinct numbers (nothing here talks to a driver);
// barrier / clear bits are distinct powers of two because the reference ORs them.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <algorithm>
typedef unsigned int GLenum; typedef unsigned int GLuint; typedef int GLint; typedef int GLsizei; typedef float GLfloat;
typedef double GLdouble; typedef short GLshort; typedef unsigned short GLushort; typedef signed char GLbyte; typedef unsigned char GLubyte;
typedef unsigned char GLboolean; typedef unsigned int GLbitfield;
enum : unsigned int {
    GL_NO_ERROR = 0, GL_TRUE = 1, GL_FALSE = 0,
    GL_SRGB = 0x0f00, GL_TEXTURE_MIN_FILTER, GL_TEXTURE_MAG_FILTER, GL_TEXTURE_WRAP_S, GL_TEXTURE_WRAP_T, GL_TEXTURE_WRAP_R,
    GL_BACK = 4096,
    GL_BYTE = 4097,
    GL_CLAMP_TO_EDGE = 4098,
    GL_DEPTH24_STENCIL8 = 4099,
    GL_DEPTH32F_STENCIL8 = 4100,
    GL_DEPTH_COMPONENT = 4101,
    GL_DEPTH_STENCIL = 4102,
    GL_DOUBLE = 4103,
    GL_DYNAMIC_COPY = 4104,
    GL_DYNAMIC_DRAW = 4105,
    GL_DYNAMIC_READ = 4106,
    GL_FILL = 4107,
    GL_FLOAT = 4108,
    GL_FRONT = 4109,
    GL_FRONT_AND_BACK = 4110,
    GL_HALF_FLOAT = 4111,
    GL_INT = 4112,
    GL_INVALID_ENUM = 4113,
    GL_INVALID_FRAMEBUFFER_OPERATION = 4114,
    GL_INVALID_OPERATION = 4115,
    GL_INVALID_VALUE = 4116,
    GL_LINE = 4117,
    GL_LINEAR = 4118,
    GL_LINES = 4119,
    GL_LINE_LOOP = 4120,
    GL_LINE_STRIP = 4121,
    GL_NEAREST = 4122,
    GL_OUT_OF_MEMORY = 4123,
    GL_POINT = 4124,
    GL_POINTS = 4125,
    GL_R16 = 4126,
    GL_R16F = 4127,
    GL_R16I = 4128,
    GL_R16UI = 4129,
    GL_R32F = 4130,
    GL_R32I = 4131,
    GL_R32UI = 4132,
    GL_R8 = 4133,
    GL_R8I = 4134,
    GL_R8UI = 4135,
    GL_READ_ONLY = 4136,
    GL_READ_WRITE = 4137,
    GL_RED = 4138,
    GL_RED_INTEGER = 4139,
    GL_REPEAT = 4140,
    GL_RG = 4141,
    GL_RG16 = 4142,
    GL_RG16F = 4143,
    GL_RG16I = 4144,
    GL_RG16UI = 4145,
    GL_RG32F = 4146,
    GL_RG32I = 4147,
    GL_RG32UI = 4148,
    GL_RG8 = 4149,
    GL_RG8I = 4150,
    GL_RG8UI = 4151,
    GL_RGB = 4152,
    GL_RGB12 = 4153,
    GL_RGB16 = 4154,
    GL_RGB16F = 4155,
    GL_RGB16I = 4156,
    GL_RGB16UI = 4157,
    GL_RGB32F = 4158,
    GL_RGB32I = 4159,
    GL_RGB32UI = 4160,
    GL_RGB8 = 4161,
    GL_RGB8I = 4162,
    GL_RGB8UI = 4163,
    GL_RGBA = 4164,
    GL_RGBA12 = 4165,
    GL_RGBA16 = 4166,
    GL_RGBA16F = 4167,
    GL_RGBA16I = 4168,
    GL_RGBA16UI = 4169,
    GL_RGBA32F = 4170,
    GL_RGBA32I = 4171,
    GL_RGBA32UI = 4172,
    GL_RGBA8 = 4173,
    GL_RGBA8I = 4174,
    GL_RGBA8UI = 4175,
    GL_RGBA_INTEGER = 4176,
    GL_RGB_INTEGER = 4177,
    GL_RG_INTEGER = 4178,
    GL_SHORT = 4179,
    GL_STACK_OVERFLOW = 4180,
    GL_STACK_UNDERFLOW = 4181,
    GL_STATIC_COPY = 4182,
    GL_STATIC_DRAW = 4183,
    GL_STATIC_READ = 4184,
    GL_STREAM_COPY = 4185,
    GL_STREAM_DRAW = 4186,
    GL_STREAM_READ = 4187,
    GL_TEXTURE_1D = 4188,
    GL_TEXTURE_1D_ARRAY = 4189,
    GL_TEXTURE_2D = 4190,
    GL_TEXTURE_2D_ARRAY = 4191,
    GL_TEXTURE_3D = 4192,
    GL_TEXTURE_BUFFER = 4193,
    GL_TRIANGLES = 4194,
    GL_TRIANGLE_FAN = 4195,
    GL_TRIANGLE_STRIP = 4196,
    GL_UNSIGNED_BYTE = 4197,
    GL_UNSIGNED_INT = 4198,
    GL_UNSIGNED_SHORT = 4199,
    GL_WRITE_ONLY = 4200,
    GL_ATOMIC_COUNTER_BARRIER_BIT = 1u,
    GL_BUFFER_UPDATE_BARRIER_BIT = 2u,
    GL_COLOR_BUFFER_BIT = 4u,
    GL_COMMAND_BARRIER_BIT = 8u,
    GL_DEPTH_BUFFER_BIT = 16u,
    GL_ELEMENT_ARRAY_BARRIER_BIT = 32u,
    GL_FRAMEBUFFER_BARRIER_BIT = 64u,
    GL_PIXEL_BUFFER_BARRIER_BIT = 128u,
    GL_SHADER_IMAGE_ACCESS_BARRIER_BIT = 256u,
    GL_SHADER_STORAGE_BARRIER_BIT = 512u,
    GL_STENCIL_BUFFER_BIT = 1024u,
    GL_TEXTURE_FETCH_BARRIER_BIT = 2048u,
    GL_TEXTURE_UPDATE_BARRIER_BIT = 4096u,
    GL_TRANSFORM_FEEDBACK_BARRIER_BIT = 8192u,
    GL_UNIFORM_BARRIER_BIT = 16384u,
    GL_VERTEX_ATTRIB_ARRAY_BARRIER_BIT = 32768u,
};
inline GLenum glGetError() { return GL_NO_ERROR; }
typedef std::ptrdiff_t GLintptr; typedef std::ptrdiff_t GLsizeiptr;
// The GL entry points that src/core/{Buffer,Texture}.cpp call, implemented over host memory in oracle/ref_gl.cpp
// (an object store: name -> bytes + geometry + formats).  Semantics: OpenGL 4.5 direct-state-access, data copied at the call.
void glCreateBuffers(GLsizei n, GLuint* ids);
void glDeleteBuffers(GLsizei n, const GLuint* ids);
void glNamedBufferData(GLuint buffer, GLsizeiptr size, const void* data, GLenum usage);
void glNamedBufferSubData(GLuint buffer, GLintptr offset, GLsizeiptr size, const void* data);
void glGetNamedBufferSubData(GLuint buffer, GLintptr offset, GLsizeiptr size, void* data);
void glCreateTextures(GLenum target, GLsizei n, GLuint* ids);
void glDeleteTextures(GLsizei n, const GLuint* ids);
void glTextureParameteri(GLuint texture, GLenum pname, GLint param);
void glClearTexImage(GLuint texture, GLint level, GLenum format, GLenum type, const void* data);
void glTextureBuffer(GLuint texture, GLenum internalformat, GLuint buffer);
void glTextureImage2DEXT(GLuint texture, GLenum target, GLint level, GLint internalformat, GLsizei width, GLsizei height,
                         GLint border, GLenum format, GLenum type, const void* pixels);
void glTextureImage3DEXT(GLuint texture, GLenum target, GLint level, GLint internalformat, GLsizei width, GLsizei height,
                         GLsizei depth, GLint border, GLenum format, GLenum type, const void* pixels);
void glTextureSubImage3D(GLuint texture, GLint level, GLint xoffset, GLint yoffset, GLint zoffset, GLsizei width, GLsizei height,
                         GLsizei depth, GLenum format, GLenum type, const void* pixels);
void glGetTextureImage(GLuint texture, GLint level, GLenum format, GLenum type, GLsizei bufSize, void* pixels);
