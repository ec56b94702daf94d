#include "ImageIO.h"
#include <new>
#include <algorithm>
#include <cstdint>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace zillum {

bool writePFM(const std::string& path, const float* rgba, int width, int height) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "PF\n%d %d\n-1.0\n", width, height);   // negative scale = little endian; rows bottom-to-top
    std::vector<float> row((size_t)width * 3);
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++)
            for (int c = 0; c < 3; c++) row[3 * x + c] = rgba[4 * ((size_t)y * width + x) + c];
        std::fwrite(row.data(), sizeof(float), row.size(), f);
    }
    std::fclose(f);
    return true;
}

// Minimal OpenEXR 2 writer: single-part scanline image, NO_COMPRESSION, channels B,G,R as FLOAT.
bool writeEXR(const std::string& path, const float* rgba, int width, int height) {
    std::vector<unsigned char> out;
    auto put = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; out.insert(out.end(), b, b + n); };
    auto putStr = [&](const char* s) { put(s, std::strlen(s) + 1); };
    auto putI32 = [&](int32_t v) { put(&v, 4); };
    auto putF32 = [&](float v) { put(&v, 4); };
    auto attr = [&](const char* name, const char* type, int32_t size) { putStr(name); putStr(type); putI32(size); };
    putI32(20000630); putI32(2);                                  // magic, version 2 / no flags
    attr("channels", "chlist", 3 * (2 + 16) + 1);
    for (const char* ch : {"B", "G", "R"}) { putStr(ch); putI32(2 /*FLOAT*/); unsigned char lin[4] = {0, 0, 0, 0}; put(lin, 4); putI32(1); putI32(1); }
    out.push_back(0);
    attr("compression", "compression", 1); out.push_back(0);
    attr("dataWindow", "box2i", 16); putI32(0); putI32(0); putI32(width - 1); putI32(height - 1);
    attr("displayWindow", "box2i", 16); putI32(0); putI32(0); putI32(width - 1); putI32(height - 1);
    attr("lineOrder", "lineOrder", 1); out.push_back(0);          // INCREASING_Y
    attr("pixelAspectRatio", "float", 4); putF32(1.0f);
    attr("screenWindowCenter", "v2f", 8); putF32(0.0f); putF32(0.0f);
    attr("screenWindowWidth", "float", 4); putF32(1.0f);
    out.push_back(0);                                             // end of header
    const size_t lineBytes = (size_t)width * 3 * 4;
    uint64_t offset = out.size() + (uint64_t)height * 8;
    for (int y = 0; y < height; y++) { put(&offset, 8); offset += 8 + lineBytes; }
    std::vector<float> line((size_t)width * 3);
    for (int y = 0; y < height; y++) {
        int filmRow = height - 1 - y;                             // EXR row 0 = top, film row 0 = bottom
        for (int c = 0; c < 3; c++)                               // channel order B, G, R
            for (int x = 0; x < width; x++) line[(size_t)c * width + x] = rgba[4 * ((size_t)filmRow * width + x) + (2 - c)];
        putI32(y); putI32((int32_t)lineBytes); put(line.data(), lineBytes);
    }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    return true;
}

// PNG = signature, IHDR, one IDAT holding a zlib stream of "stored" deflate blocks, IEND.
bool writePNG(const std::string& path, const unsigned char* rgb8, int width, int height) {
    if (width <= 0 || height <= 0 || !rgb8) return false;
    static uint32_t crcTable[256];
    static bool crcReady = false;
    if (!crcReady) {
        for (uint32_t n = 0; n < 256; n++) { uint32_t c = n; for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1; crcTable[n] = c; }
        crcReady = true;
    }
    std::vector<unsigned char> raw;                                   // filter byte 0 + RGB per scanline, top row first
    raw.reserve((size_t)height * (1 + (size_t)width * 3));
    for (int y = 0; y < height; y++) {
        raw.push_back(0);
        const unsigned char* row = rgb8 + (size_t)(height - 1 - y) * width * 3;
        raw.insert(raw.end(), row, row + (size_t)width * 3);
    }
    std::vector<unsigned char> z;
    z.push_back(0x78); z.push_back(0x01);                             // zlib header: deflate, 32 K window, no preset dictionary
    uint32_t a = 1, b = 0;                                            // Adler-32 of the raw data
    for (size_t off = 0; off < raw.size();) {
        const size_t len = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + len == raw.size() ? 1 : 0);                 // BFINAL, BTYPE = 00 (stored)
        z.push_back((unsigned char)(len & 0xff)); z.push_back((unsigned char)(len >> 8));
        z.push_back((unsigned char)(~len & 0xff)); z.push_back((unsigned char)((~len >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
        for (size_t i = off; i < off + len; i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
        off += len;
    }
    const uint32_t adler = (b << 16) | a;
    for (int s = 24; s >= 0; s -= 8) z.push_back((unsigned char)(adler >> s));
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    auto chunk = [&](const char* type, const std::vector<unsigned char>& data) {
        unsigned char len[4] = {(unsigned char)(data.size() >> 24), (unsigned char)(data.size() >> 16), (unsigned char)(data.size() >> 8), (unsigned char)data.size()};
        std::fwrite(len, 1, 4, f);
        uint32_t c = 0xffffffffu;
        for (int i = 0; i < 4; i++) c = crcTable[(c ^ (unsigned char)type[i]) & 0xff] ^ (c >> 8);
        for (unsigned char d : data) c = crcTable[(c ^ d) & 0xff] ^ (c >> 8);
        c ^= 0xffffffffu;
        std::fwrite(type, 1, 4, f);
        if (!data.empty()) std::fwrite(data.data(), 1, data.size(), f);
        unsigned char crc[4] = {(unsigned char)(c >> 24), (unsigned char)(c >> 16), (unsigned char)(c >> 8), (unsigned char)c};
        std::fwrite(crc, 1, 4, f);
    };
    const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr = {(unsigned char)(width >> 24), (unsigned char)(width >> 16), (unsigned char)(width >> 8), (unsigned char)width,
                                       (unsigned char)(height >> 24), (unsigned char)(height >> 16), (unsigned char)(height >> 8), (unsigned char)height,
                                       8, 2, 0, 0, 0};                // 8 bits, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    chunk("IHDR", ihdr);
    chunk("IDAT", z);
    chunk("IEND", {});
    return std::fclose(f) == 0;
}

static bool endsWith(const std::string& s, const char* suf) {
    size_t n = std::strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; i++) if (std::tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
    return true;
}

static bool loadPFM(const std::string& path, std::vector<float>& rgb, int& width, int& height) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string magic; float scale;
    f >> magic >> width >> height >> scale;
    f.get();
    int ch = magic == "PF" ? 3 : (magic == "Pf" ? 1 : 0);
    if (!f || !ch || width <= 0 || height <= 0 || (size_t)width * (size_t)height > ((size_t)1 << 28)) return false;
    std::vector<float> raw((size_t)width * height * ch);
    f.read((char*)raw.data(), raw.size() * 4);
    if (!f) return false;
    if (scale > 0) for (auto& v : raw) { unsigned char* b = (unsigned char*)&v; std::swap(b[0], b[3]); std::swap(b[1], b[2]); }
    rgb.resize((size_t)width * height * 3);
    for (int y = 0; y < height; y++)                              // PFM rows are bottom-to-top
        for (int x = 0; x < width; x++)
            for (int c = 0; c < 3; c++)
                rgb[3 * ((size_t)y * width + x) + c] = raw[((size_t)(height - 1 - y) * width + x) * ch + (ch == 3 ? c : 0)];
    return true;
}

static bool loadHDR(const std::string& path, std::vector<float>& rgb, int& width, int& height) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char line[256];
    bool ok = false;
    while (std::fgets(line, sizeof line, f)) {
        if (line[0] == '\n') break;
        if (std::strstr(line, "FORMAT=32-bit_rle_rgbe")) ok = true;
    }
    if (!ok || !std::fgets(line, sizeof line, f) || std::sscanf(line, "-Y %d +X %d", &height, &width) != 2 ||
        width <= 0 || height <= 0 || (size_t)width * (size_t)height > ((size_t)1 << 28)) { std::fclose(f); return false; }
    rgb.resize((size_t)width * height * 3);
    std::vector<unsigned char> scan((size_t)width * 4);
    for (int y = 0; y < height; y++) {
        unsigned char h[4];
        if (std::fread(h, 1, 4, f) != 4) { std::fclose(f); return false; }
        if (h[0] == 2 && h[1] == 2 && !(h[2] & 0x80) && ((h[2] << 8) | h[3]) == width) {   // new-style RLE
            for (int c = 0; c < 4; c++)
                for (int x = 0; x < width;) {
                    int n = std::fgetc(f);
                    if (n <= 0) { std::fclose(f); return false; }            // end of file or a zero-length run: corrupt
                    if (n > 128) { int v = std::fgetc(f); n -= 128; while (n-- && x < width) scan[4 * x++ + c] = (unsigned char)v; }
                    else while (n-- && x < width) scan[4 * x++ + c] = (unsigned char)std::fgetc(f);
                }
        } else {                                                                           // flat
            std::memcpy(scan.data(), h, 4);
            if (std::fread(scan.data() + 4, 1, (size_t)width * 4 - 4, f) != (size_t)width * 4 - 4) { std::fclose(f); return false; }
        }
        for (int x = 0; x < width; x++) {
            const unsigned char* p = &scan[4 * x];
            float s = p[3] ? std::ldexp(1.0f, (int)p[3] - (128 + 8)) : 0.0f;
            for (int c = 0; c < 3; c++) rgb[3 * ((size_t)y * width + x) + c] = p[c] * s;
        }
    }
    std::fclose(f);
    return true;
}

bool loadFloatImage(const std::string& path, std::vector<float>& rgb, int& width, int& height) {
    try {
        if (endsWith(path, ".pfm")) return loadPFM(path, rgb, width, height);
        if (endsWith(path, ".hdr")) return loadHDR(path, rgb, width, height);
    } catch (const std::bad_alloc&) { return false; }
    // an 8-bit file as a float image (the commented res/scene.xml names a .png environment map): stbi_loadf's
    // LDR -> HDR conversion, pow(v / 255, 2.2) with the quotient in float and the power in double
    std::vector<unsigned char> ldr;
    if (!loadByteImage(path, ldr, width, height)) return false;
    float lut[256];
    for (int v = 0; v < 256; v++) lut[v] = (float)std::pow((double)((float)v / 255.0f), 2.2);
    rgb.resize(ldr.size());
    for (size_t i = 0; i < ldr.size(); i++) rgb[i] = lut[ldr[i]];
    return true;
}

}  // namespace zillum
