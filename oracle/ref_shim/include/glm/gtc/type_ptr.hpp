// ORACLE — test infrastructure only (see glm/glm.hpp in this directory).
#pragma once
#include "../glm.hpp"
namespace glm {
inline const float* value_ptr(const vec2& v) { return &v.x; }
inline const float* value_ptr(const vec3& v) { return &v.x; }
inline const float* value_ptr(const vec4& v) { return &v.x; }
inline const float* value_ptr(const mat3& m) { return &m.c[0].x; }
inline const float* value_ptr(const mat4& m) { return &m.c[0].x; }
}  // namespace glm
