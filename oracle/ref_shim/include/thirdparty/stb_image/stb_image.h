// ORACLE — test infrastructure only.  Stand-in for stb_image (absent): src/core/Image.h only needs the include to resolve;
// Image::createFromFile is implemented over in-memory images in oracle/ref_host.cpp.
#pragma once
#include <memory>
#include <map>
#include <string>
// image files do not exist here: "paths" name in-memory images registered with zr_register_image (oracle/ref_gl.cpp)
unsigned char* stbi_load(const char* path, int* w, int* h, int* channelsInFile, int desiredChannels);
float* stbi_loadf(const char* path, int* w, int* h, int* channelsInFile, int desiredChannels);
void stbi_image_free(void* p);
