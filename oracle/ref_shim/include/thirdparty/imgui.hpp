// ORACLE — test infrastructure only.  Stand-in for Dear ImGui (absent): the settings panels of
// src/integrator/{NaivePath,LightPath,TriplePath}.cpp compile against these no-ops (headless: no widget ever changes a value).
#pragma once
#include <string>
struct ImVec2 { float x, y; ImVec2(float a = 0, float b = 0) : x(a), y(b) {} };
#define IM_ARRAYSIZE(a) ((int)(sizeof(a) / sizeof(*(a))))
namespace ImGui {
inline void SetNextItemWidth(float) {}
inline void SameLine() {}
inline void Text(const char*, ...) {}
inline bool InputInt(const char*, int*, int = 1, int = 100) { return false; }
inline bool Checkbox(const char*, bool*) { return false; }
inline bool Combo(const char*, int*, const char* const[], int) { return false; }
inline bool SliderFloat(const char*, float*, float, float) { return false; }
inline void ProgressBar(float, const ImVec2& = ImVec2(-1, 0), const char* = nullptr) {}
}  // namespace ImGui
