// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref.
// GLFW/glfw3.h — stand-in: the key names src/core/Camera.cpp:16-52 switches on (no window system here).
#pragma once
enum { GLFW_KEY_SPACE = 32, GLFW_KEY_A = 65, GLFW_KEY_D = 68, GLFW_KEY_E = 69, GLFW_KEY_Q = 81, GLFW_KEY_R = 82,
       GLFW_KEY_S = 83, GLFW_KEY_W = 87, GLFW_KEY_LEFT_SHIFT = 340 };
struct GLFWwindow;
