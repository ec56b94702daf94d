// Headless image output/input: linear-radiance PFM and OpenEXR (scanline, uncompressed,
// 32-bit float) writers replace the reference's tone-mapped PNG screenshot
// (src/Application.cpp:371-380); PFM and Radiance .hdr (RGBE) readers stand in for
// stbi_loadf (src/core/Image.cpp:10-34).
#pragma once
#include <string>
#include <vector>

namespace zillum {

// rgba: width*height*4 floats, row 0 = bottom of the image (film convention).
bool writePFM(const std::string& path, const float* rgba, int width, int height);
bool writeEXR(const std::string& path, const float* rgba, int width, int height);
// 8-bit RGB PNG (stored deflate blocks: no compression library in the image), stands in for stbi_write_png with
// stbi_flip_vertically_on_write(true) (src/Application.cpp:371-380): rgb8 rows are in film order (row 0 = bottom)
// and are written top row first.
bool writePNG(const std::string& path, const unsigned char* rgb8, int width, int height);
// rgb out: width*height*3 floats, row 0 = top of the image (stb convention).  PFM, Radiance .hdr, or any 8-bit format of
// loadByteImage converted like stbi_loadf does (pow(v / 255, 2.2)).
bool loadFloatImage(const std::string& path, std::vector<float>& rgb, int& width, int& height);
// 8-bit RGB, row 0 = top; stands in for stbi_load(path, ..., 3) on albedo textures (ImageDecode.cpp): PNG (all colour
// types and bit depths, Adam7), sequential and progressive JPEG, TGA (incl. RLE and palettes), BMP, binary PPM;
// the format is recognised by content, not by extension.  false: unreadable or unsupported (e.g. CMYK or arithmetic-coded JPEG).
bool loadByteImage(const std::string& path, std::vector<unsigned char>& rgb, int& width, int& height);

}  // namespace zillum
