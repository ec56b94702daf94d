// 8-bit image decoders for albedo textures: PNG, TGA, BMP and JPEG (sequential and progressive), written against the
// file-format specifications (RFC 1950 / 1951 / 2083, ITU T.81, the Truevision and BMP headers).  They stand in for
// stbi_load(path, &w, &h, &n, 3) in the reference's texture path (src/core/Image.cpp:10-34, src/core/Texture.cpp:134-171):
// the result is always 3 channels of 8 bits, row 0 = top of the image; grey is replicated, alpha is dropped, 16-bit
// samples keep their high byte, palettes are expanded.  No third-party code and no compression library is used.
#include "ImageIO.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <new>
#include <stdexcept>

namespace zillum {
namespace {

using Bytes = std::vector<unsigned char>;

bool readFile(const std::string& path, Bytes& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    std::streamoff n = f.tellg();
    if (n <= 0) return false;
    f.seekg(0);
    out.resize((size_t)n);
    f.read((char*)out.data(), n);
    return (bool)f;
}

// ---------------------------------------------------------------------------------------------
// inflate (RFC 1951) inside a zlib wrapper (RFC 1950)
// ---------------------------------------------------------------------------------------------
struct BitReader {
    const unsigned char* p; size_t n, pos = 0; uint32_t acc = 0; int cnt = 0; bool bad = false;
    BitReader(const unsigned char* p_, size_t n_) : p(p_), n(n_) {}
    uint32_t bits(int k) {            // k <= 16, LSB first
        while (cnt < k) {
            if (pos >= n) { bad = true; return 0; }
            acc |= (uint32_t)p[pos++] << cnt; cnt += 8;
        }
        uint32_t v = acc & ((1u << k) - 1u);
        acc >>= k; cnt -= k;
        return v;
    }
    void alignByte() { acc = 0; cnt = 0; }
};
struct Huffman {            // canonical code, decoded bit by bit with per-length first-code tables
    uint16_t count[16] = {0}, symbol[320] = {0};
    bool build(const unsigned char* lengths, int n) {
        std::memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        int left = 1;
        for (int len = 1; len < 16; len++) { left <<= 1; left -= count[len]; if (left < 0) return false; }
        uint16_t offs[16]; offs[1] = 0;
        for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + count[len];
        for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; len++) {
            code |= (int)br.bits(1);
            if (br.bad) return -1;
            int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};
bool inflateZlib(const unsigned char* src, size_t n, Bytes& out) {
    if (n < 6 || (src[0] & 0x0f) != 8 || ((src[0] << 8) | src[1]) % 31 != 0 || (src[1] & 0x20)) return false;
    BitReader br(src + 2, n - 2);
    static const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint16_t lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint16_t distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    bool last = false;
    while (!last) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (br.bad) return false;
        if (type == 0) {
            br.alignByte();
            if (br.pos + 4 > br.n) return false;
            const unsigned len = br.p[br.pos] | (br.p[br.pos + 1] << 8), nlen = br.p[br.pos + 2] | (br.p[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xffffu) != nlen || br.pos + len > br.n) return false;
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
            continue;
        }
        if (type == 3) return false;
        Huffman lit, dist;
        unsigned char lengths[320];
        if (type == 1) {
            int i = 0;
            for (; i < 144; i++) lengths[i] = 8;
            for (; i < 256; i++) lengths[i] = 9;
            for (; i < 280; i++) lengths[i] = 7;
            for (; i < 288; i++) lengths[i] = 8;
            lit.build(lengths, 288);
            for (i = 0; i < 30; i++) lengths[i] = 5;
            dist.build(lengths, 30);
        } else {
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            if (br.bad || nlen > 286 || ndist > 30) return false;
            static const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            unsigned char cl[19] = {0};
            for (int i = 0; i < ncode; i++) cl[order[i]] = (unsigned char)br.bits(3);
            Huffman clh;
            if (!clh.build(cl, 19)) return false;
            int i = 0;
            while (i < nlen + ndist) {
                const int sym = clh.decode(br);
                if (sym < 0) return false;
                if (sym < 16) { lengths[i++] = (unsigned char)sym; continue; }
                int rep = 0; unsigned char val = 0;
                if (sym == 16) { if (i == 0) return false; val = lengths[i - 1]; rep = 3 + (int)br.bits(2); }
                else if (sym == 17) rep = 3 + (int)br.bits(3);
                else rep = 11 + (int)br.bits(7);
                if (i + rep > nlen + ndist) return false;
                while (rep--) lengths[i++] = val;
            }
            if (!lit.build(lengths, nlen) || !dist.build(lengths + nlen, ndist)) return false;
        }
        while (true) {
            const int sym = lit.decode(br);
            if (sym < 0) return false;
            if (sym < 256) { out.push_back((unsigned char)sym); continue; }
            if (sym == 256) break;
            if (sym > 285) return false;
            const int len = lenBase[sym - 257] + (int)br.bits(lenExtra[sym - 257]);
            const int ds = dist.decode(br);
            if (ds < 0 || ds > 29) return false;
            const size_t d = distBase[ds] + br.bits(distExtra[ds]);
            if (br.bad || d > out.size()) return false;
            const size_t from = out.size() - d;
            for (int k = 0; k < len; k++) out.push_back(out[from + k]);
        }
    }
    return !br.bad;
}

// ---------------------------------------------------------------------------------------------
// PNG (RFC 2083): all colour types, 1-16 bits, tRNS ignored, non-interlaced and Adam7
// ---------------------------------------------------------------------------------------------
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// un-filter one pass of `h` scanlines of `rowBytes` bytes (each preceded by its filter byte), in place -> tightly packed rows
bool pngUnfilter(const unsigned char* src, size_t avail, int h, size_t rowBytes, int bpp, Bytes& rows) {
    if (avail < (size_t)h * (rowBytes + 1)) return false;
    rows.assign((size_t)h * rowBytes, 0);
    for (int y = 0; y < h; y++) {
        const unsigned char* in = src + (size_t)y * (rowBytes + 1);
        unsigned char* cur = rows.data() + (size_t)y * rowBytes;
        const unsigned char* up = y ? cur - rowBytes : nullptr;
        const int ft = in[0];
        in++;
        for (size_t i = 0; i < rowBytes; i++) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
            int pred = 0;
            switch (ft) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) >> 1; break;
            case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                      pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
            default: return false;
            }
            cur[i] = (unsigned char)(in[i] + pred);
        }
    }
    return true;
}
bool loadPNG(const Bytes& file, Bytes& rgb, int& width, int& height) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || std::memcmp(file.data(), sig, 8) != 0) return false;
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    Bytes idat, palette;
    bool haveHdr = false;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        const unsigned char* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) return false;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) return false;
            width = (int)be32(data); height = (int)be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] != 0 || data[11] != 0 || interlace > 1) return false;
            haveHdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!haveHdr || width <= 0 || height <= 0 || width > (1 << 15) || height > (1 << 15)) return false;
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels) return false;
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) return false;
    if ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) return false;
    if (ctype == 3 && (depth == 16 || palette.size() < 3)) return false;
    Bytes raw;
    if (!inflateZlib(idat.data(), idat.size(), raw)) return false;
    const int bitsPerPixel = channels * depth, bpp = std::max(1, bitsPerPixel / 8);
    rgb.assign((size_t)width * height * 3, 0);
    // sample c of pixel x in an un-filtered row -> 8 bits (16-bit: high byte; < 8 bits grey: scaled to 0..255; palette: index)
    auto sample = [&](const unsigned char* row, int x, int c) -> int {
        if (depth == 8) return row[(size_t)x * channels + c];
        if (depth == 16) return row[((size_t)x * channels + c) * 2];
        const int bit = x * depth, v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
        return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
    };
    auto put = [&](const unsigned char* row, int xs, int px, int py) {
        unsigned char* o = &rgb[((size_t)py * width + px) * 3];
        if (ctype == 3) {
            const size_t idx = (size_t)sample(row, xs, 0) * 3;
            if (idx + 2 < palette.size()) { o[0] = palette[idx]; o[1] = palette[idx + 1]; o[2] = palette[idx + 2]; }
        } else if (channels <= 2) { o[0] = o[1] = o[2] = (unsigned char)sample(row, xs, 0); }
        else { o[0] = (unsigned char)sample(row, xs, 0); o[1] = (unsigned char)sample(row, xs, 1); o[2] = (unsigned char)sample(row, xs, 2); }
    };
    Bytes rows;
    if (!interlace) {
        const size_t rowBytes = ((size_t)width * bitsPerPixel + 7) / 8;
        if (!pngUnfilter(raw.data(), raw.size(), height, rowBytes, bpp, rows)) return false;
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) put(rows.data() + (size_t)y * rowBytes, x, x, y);
        return true;
    }
    static const int x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t off = 0;
    for (int p = 0; p < 7; p++) {
        const int pw = (width - x0[p] + dx[p] - 1) / dx[p], ph = (height - y0[p] + dy[p] - 1) / dy[p];
        if (pw <= 0 || ph <= 0) continue;
        const size_t rowBytes = ((size_t)pw * bitsPerPixel + 7) / 8;
        if (off > raw.size() || !pngUnfilter(raw.data() + off, raw.size() - off, ph, rowBytes, bpp, rows)) return false;
        off += (size_t)ph * (rowBytes + 1);
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++) put(rows.data() + (size_t)y * rowBytes, x, x0[p] + x * dx[p], y0[p] + y * dy[p]);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// TGA: types 1/2/3 and their RLE forms 9/10/11; 8 (grey or palette), 15/16, 24, 32 bits
// ---------------------------------------------------------------------------------------------
bool loadTGA(const Bytes& f, Bytes& rgb, int& width, int& height) {
    if (f.size() < 18) return false;
    const int idLen = f[0], cmapType = f[1], type = f[2], cmapFirst = f[3] | (f[4] << 8), cmapLen = f[5] | (f[6] << 8), cmapBits = f[7];
    width = f[12] | (f[13] << 8); height = f[14] | (f[15] << 8);
    const int bits = f[16], desc = f[17];
    const bool rle = type >= 9;
    const int base = rle ? type - 8 : type;
    if (width <= 0 || height <= 0 || base < 1 || base > 3 || cmapType > 1) return false;
    if (base == 1 && (!cmapType || bits != 8)) return false;
    if (base == 2 && !(bits == 15 || bits == 16 || bits == 24 || bits == 32)) return false;
    if (base == 3 && bits != 8 && bits != 16) return false;
    size_t pos = 18 + (size_t)idLen;
    const int cmapBytes = (cmapBits + 7) / 8;
    const unsigned char* cmap = nullptr;
    if (cmapType) { cmap = f.data() + pos; pos += (size_t)cmapLen * cmapBytes; if (pos > f.size()) return false; }
    const int pixBytes = (bits + 7) / 8;
    auto toRgb = [&](const unsigned char* p, int nbits, unsigned char* o) {
        if (nbits == 24 || nbits == 32) { o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; }
        else if (nbits == 15 || nbits == 16) {
            const int v = p[0] | (p[1] << 8);
            o[0] = (unsigned char)(((v >> 10) & 31) * 255 / 31); o[1] = (unsigned char)(((v >> 5) & 31) * 255 / 31); o[2] = (unsigned char)((v & 31) * 255 / 31);
        } else { o[0] = o[1] = o[2] = p[0]; }
    };
    rgb.assign((size_t)width * height * 3, 0);
    const size_t total = (size_t)width * height;
    size_t i = 0;
    unsigned char px[4] = {0, 0, 0, 0};
    int run = 0; bool runRaw = true;
    while (i < total) {
        if (rle) {
            if (run == 0) {
                if (pos >= f.size()) return false;
                const int h = f[pos++];
                run = (h & 127) + 1; runRaw = !(h & 128);
                if (!runRaw) { if (pos + pixBytes > f.size()) return false; std::memcpy(px, &f[pos], pixBytes); pos += pixBytes; }
            }
            if (runRaw) { if (pos + pixBytes > f.size()) return false; std::memcpy(px, &f[pos], pixBytes); pos += pixBytes; }
            run--;
        } else {
            if (pos + pixBytes > f.size()) return false;
            std::memcpy(px, &f[pos], pixBytes); pos += pixBytes;
        }
        const int x = (int)(i % width), yFile = (int)(i / width);
        const int y = (desc & 0x20) ? yFile : height - 1 - yFile;          // bit 5 clear: bottom-up file
        const int xx = (desc & 0x10) ? width - 1 - x : x;
        unsigned char* o = &rgb[((size_t)y * width + xx) * 3];
        if (base == 1) {
            const int idx = px[0] - cmapFirst;
            if (idx < 0 || idx >= cmapLen) return false;
            toRgb(cmap + (size_t)idx * cmapBytes, cmapBits, o);
        } else if (base == 3) { o[0] = o[1] = o[2] = px[0]; }
        else toRgb(px, bits, o);
        i++;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// BMP: BITMAPINFOHEADER family, uncompressed 8 (palette), 24 and 32 bits
// ---------------------------------------------------------------------------------------------
bool loadBMP(const Bytes& f, Bytes& rgb, int& width, int& height) {
    if (f.size() < 54 || f[0] != 'B' || f[1] != 'M') return false;
    auto le32 = [&](size_t o) { return (uint32_t)f[o] | ((uint32_t)f[o + 1] << 8) | ((uint32_t)f[o + 2] << 16) | ((uint32_t)f[o + 3] << 24); };
    const uint32_t dataOff = le32(10), hdr = le32(14);
    if (hdr < 40) return false;
    width = (int)le32(18);
    int h = (int)le32(22);
    const int bits = f[28] | (f[29] << 8);
    const uint32_t comp = le32(30);
    const bool topDown = h < 0;
    height = topDown ? -h : h;
    if (width <= 0 || height <= 0 || !(comp == 0 || (comp == 3 && bits == 32)) || !(bits == 8 || bits == 24 || bits == 32)) return false;
    const size_t stride = (((size_t)width * bits + 31) / 32) * 4;
    if ((size_t)dataOff + stride * height > f.size()) return false;
    const unsigned char* pal = f.data() + 14 + hdr;
    rgb.assign((size_t)width * height * 3, 0);
    for (int y = 0; y < height; y++) {
        const unsigned char* row = f.data() + dataOff + stride * (size_t)(topDown ? y : height - 1 - y);
        for (int x = 0; x < width; x++) {
            unsigned char* o = &rgb[((size_t)y * width + x) * 3];
            if (bits == 8) { const unsigned char* p = pal + 4 * (size_t)row[x]; if (p + 3 > f.data() + f.size()) return false; o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; }
            else { const unsigned char* p = row + (size_t)x * (bits / 8); o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// JPEG (ITU T.81): baseline, extended sequential and progressive DCT, Huffman coding, 8-bit, 1 or 3 components, any sampling
// factors, interleaved and single-component scans, restart intervals.  Floating-point IDCT, bilinear ("triangle") chroma up-sampling for 2x factors, JFIF YCbCr -> RGB.
// ---------------------------------------------------------------------------------------------
struct JpegHuff {
    unsigned char bits[17] = {0}, vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool present = false;
    void prepare() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};
struct JpegBits {
    const unsigned char* p; size_t n, pos; uint32_t acc = 0; int cnt = 0; bool marker = false;
    int bit() {
        if (cnt == 0) {
            int b = 0;
            if (!marker && pos < n) {
                b = p[pos++];
                if (b == 0xff) {
                    const int b2 = pos < n ? p[pos] : 0xd9;
                    if (b2 == 0) pos++;
                    else { marker = true; pos--; b = 0; }      // a marker ends the entropy-coded segment: feed zeros
                }
            }
            acc = (uint32_t)b; cnt = 8;
        }
        cnt--;
        return (acc >> cnt) & 1;
    }
    int receive(int s) { int v = 0; while (s--) v = (v << 1) | bit(); return v; }
    void reset() { acc = 0; cnt = 0; marker = false; }
};
int jpegDecodeSym(JpegBits& br, const JpegHuff& h) {
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    return -1;
}
int jpegExtend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
void jpegIdct(const int* coef, const uint16_t* q, unsigned char* out, int stride) {
    static float cosT[8][8]; static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; x++) for (int u = 0; u < 8; u++) cosT[x][u] = (float)((u == 0 ? std::sqrt(0.5) : 1.0) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
        init = true;
    }
    static const int zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                               35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    float blk[64], tmp[64];
    for (int i = 0; i < 64; i++) blk[i] = 0.0f;
    for (int i = 0; i < 64; i++) blk[zz[i]] = (float)(coef[i] * (int)q[i]);
    for (int y = 0; y < 8; y++)          // rows: tmp[y][x] = sum_u C(u) blk[y][u] cos
        for (int x = 0; x < 8; x++) { float s = 0; for (int u = 0; u < 8; u++) s += blk[y * 8 + u] * cosT[x][u]; tmp[y * 8 + x] = s; }
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++) {
            float s = 0; for (int v = 0; v < 8; v++) s += tmp[v * 8 + x] * cosT[y][v];
            const int val = (int)std::floor(s * 0.25f + 128.5f);
            out[y * stride + x] = (unsigned char)(val < 0 ? 0 : val > 255 ? 255 : val);
        }
}
// One decoder for sequential (SOF0 / SOF1) and progressive (SOF2) files: every scan decodes into per-component coefficient
// arrays (zig-zag order), and the blocks are de-quantised and transformed once after the last scan.  Scans may be interleaved
// (MCU = h x v blocks of each component) or hold a single component (blocks in that component's own raster, T.81 A.2.3).
struct JpegComp {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0;
    int blocksW = 0, blocksH = 0;       // allocated block grid (padded to whole MCUs)
    int stride = 0, hgt = 0;            // sample plane
    std::vector<int16_t> coef;
    Bytes data;
};
bool loadJPEG(const Bytes& f, Bytes& rgb, int& width, int& height) {
    if (f.size() < 4 || f[0] != 0xff || f[1] != 0xd8) return false;
    uint16_t qt[4][64] = {{0}};
    JpegHuff dc[4], ac[4];
    JpegComp comp[3];
    int ncomp = 0, hmax = 1, vmax = 1, restart = 0, mcux = 0, mcuy = 0;
    bool haveFrame = false, progressive = false, haveScan = false;
    size_t pos = 2;
    while (pos + 4 <= f.size()) {
        if (f[pos] != 0xff) { pos++; continue; }
        const int m = f[pos + 1];
        if (m == 0xff) { pos++; continue; }
        pos += 2;
        if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01 || m == 0x00) continue;
        if (m == 0xd9) break;
        if (pos + 2 > f.size()) return false;
        const size_t len = ((size_t)f[pos] << 8) | f[pos + 1];
        if (len < 2 || pos + len > f.size()) return false;
        const unsigned char* d = &f[pos + 2];
        const size_t dn = len - 2;
        if (m == 0xdb) {
            size_t i = 0;
            while (i < dn) {
                const int pq = d[i] >> 4, tq = d[i] & 15; i++;
                if (tq > 3 || i + (pq ? 128 : 64) > dn) return false;
                for (int k = 0; k < 64; k++) { qt[tq][k] = pq ? (uint16_t)((d[i] << 8) | d[i + 1]) : d[i]; i += pq ? 2 : 1; }
            }
        } else if (m == 0xc4) {
            size_t i = 0;
            while (i + 17 <= dn) {
                const int tc = d[i] >> 4, th = d[i] & 15; i++;
                if (th > 3 || tc > 1) return false;
                JpegHuff& h = tc ? ac[th] : dc[th];
                int total = 0;
                for (int l = 1; l <= 16; l++) { h.bits[l] = d[i++]; total += h.bits[l]; }
                if (total > 256 || i + total > dn) return false;
                std::memcpy(h.vals, d + i, total); i += total;
                h.prepare(); h.present = true;
            }
        } else if (m == 0xc0 || m == 0xc1 || m == 0xc2) {
            if (haveFrame || dn < 6 || d[0] != 8) return false;
            progressive = m == 0xc2;
            height = (d[1] << 8) | d[2]; width = (d[3] << 8) | d[4]; ncomp = d[5];
            if ((ncomp != 1 && ncomp != 3) || dn < 6 + 3 * (size_t)ncomp || width <= 0 || height <= 0) return false;
            for (int c = 0; c < ncomp; c++) {
                comp[c].id = d[6 + 3 * c]; comp[c].h = d[7 + 3 * c] >> 4; comp[c].v = d[7 + 3 * c] & 15; comp[c].tq = d[8 + 3 * c];
                if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) return false;
                hmax = std::max(hmax, comp[c].h); vmax = std::max(vmax, comp[c].v);
            }
            mcux = (width + 8 * hmax - 1) / (8 * hmax); mcuy = (height + 8 * vmax - 1) / (8 * vmax);
            for (int c = 0; c < ncomp; c++) {
                comp[c].blocksW = mcux * comp[c].h; comp[c].blocksH = mcuy * comp[c].v;
                comp[c].coef.assign((size_t)comp[c].blocksW * comp[c].blocksH * 64, 0);
            }
            haveFrame = true;
        } else if (m >= 0xc3 && m <= 0xcf && m != 0xc8 && m != 0xcc) {
            return false;                               // lossless, hierarchical, arithmetic coding: not supported
        } else if (m == 0xdd) {
            if (dn < 2) return false;
            restart = (d[0] << 8) | d[1];
        } else if (m == 0xda) {
            if (!haveFrame || dn < 1) return false;
            const int ns = d[0];
            if (ns < 1 || ns > ncomp || dn < 1 + 2 * (size_t)ns + 3) return false;
            int sc[3];
            for (int c = 0; c < ns; c++) {
                int k = -1;
                for (int j = 0; j < ncomp; j++) if (comp[j].id == d[1 + 2 * c]) k = j;
                if (k < 0) return false;
                comp[k].td = d[2 + 2 * c] >> 4; comp[k].ta = d[2 + 2 * c] & 15;
                if (comp[k].td > 3 || comp[k].ta > 3) return false;
                sc[c] = k;
            }
            const int Ss = d[1 + 2 * ns], Se = d[2 + 2 * ns], Ah = d[3 + 2 * ns] >> 4, Al = d[3 + 2 * ns] & 15;
            if (!progressive) { if (Ss != 0 || Se != 63 || Ah != 0 || Al != 0) return false; }
            else if (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13) return false;
            for (int c = 0; c < ns; c++) {
                if ((Ss == 0 && Ah == 0 && !dc[comp[sc[c]].td].present) || (Se > 0 && !ac[comp[sc[c]].ta].present)) return false;
                comp[sc[c]].pred = 0;
            }
            JpegBits br{f.data(), f.size(), pos + len};
            int eobrun = 0, untilRestart = restart;
            // block of one MCU element; T.81 F.2.2 (sequential), G.1.2 (progressive)
            auto decodeBlock = [&](JpegComp& c, int16_t* blk) -> bool {
                if (!progressive) {
                    const int t = jpegDecodeSym(br, dc[c.td]);
                    if (t < 0 || t > 15) return false;
                    c.pred += jpegExtend(br.receive(t), t);
                    blk[0] = (int16_t)c.pred;
                    for (int k = 1; k < 64;) {
                        const int rs = jpegDecodeSym(br, ac[c.ta]);
                        if (rs < 0) return false;
                        const int r = rs >> 4, s = rs & 15;
                        if (s == 0) { if (r == 15) { k += 16; continue; } break; }
                        k += r;
                        if (k > 63) return false;
                        blk[k++] = (int16_t)jpegExtend(br.receive(s), s);
                    }
                    return true;
                }
                if (Ss == 0) {                                         // DC scan
                    if (Ah == 0) {
                        const int t = jpegDecodeSym(br, dc[c.td]);
                        if (t < 0 || t > 15) return false;
                        c.pred += jpegExtend(br.receive(t), t);
                        blk[0] = (int16_t)(c.pred * (1 << Al));
                    } else if (br.bit()) blk[0] = (int16_t)(blk[0] | (1 << Al));
                    return true;
                }
                const int p1 = 1 << Al, m1 = -(1 << Al);
                if (Ah == 0) {                                         // AC, first pass of the band
                    if (eobrun > 0) { eobrun--; return true; }
                    for (int k = Ss; k <= Se; k++) {
                        const int rs = jpegDecodeSym(br, ac[c.ta]);
                        if (rs < 0) return false;
                        const int r = rs >> 4, s = rs & 15;
                        if (s) {
                            k += r;
                            if (k > Se) return false;
                            blk[k] = (int16_t)(jpegExtend(br.receive(s), s) * p1);
                        } else if (r == 15) k += 15;
                        else { eobrun = (1 << r) - 1; if (r) eobrun += br.receive(r); break; }
                    }
                    return true;
                }
                // AC refinement: correction bits for the coefficients that are already non-zero, new +-1 coefficients in between
                auto refine = [&](int16_t& v) { if (br.bit() && (v & p1) == 0) v = (int16_t)(v + (v >= 0 ? p1 : m1)); };
                int k = Ss;
                if (eobrun == 0) {
                    for (; k <= Se; k++) {
                        const int rs = jpegDecodeSym(br, ac[c.ta]);
                        if (rs < 0) return false;
                        int r = rs >> 4, s = rs & 15, val = 0;
                        if (s) val = br.bit() ? p1 : m1;
                        else if (r != 15) { eobrun = 1 << r; if (r) eobrun += br.receive(r); break; }
                        for (; k <= Se; k++) {
                            if (blk[k] != 0) refine(blk[k]);
                            else if (--r < 0) break;
                        }
                        if (s && k <= Se) blk[k] = (int16_t)val;
                    }
                }
                if (eobrun > 0) {
                    for (; k <= Se; k++) if (blk[k] != 0) refine(blk[k]);
                    eobrun--;
                }
                return true;
            };
            auto restartIfDue = [&]() -> bool {
                if (!restart || untilRestart > 0) return true;
                size_t p = br.pos;
                while (p + 1 < f.size() && !(f[p] == 0xff && f[p + 1] >= 0xd0 && f[p + 1] <= 0xd7)) p++;
                if (p + 1 >= f.size()) return false;
                br.pos = p + 2; br.reset();
                for (int c = 0; c < ncomp; c++) comp[c].pred = 0;
                eobrun = 0;
                untilRestart = restart;
                return true;
            };
            if (ns == 1) {                                             // single-component scan: the component's own block raster
                JpegComp& c = comp[sc[0]];
                const int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
                const int bw = (cw + 7) / 8, bh = (ch + 7) / 8;
                for (int by = 0; by < bh; by++)
                    for (int bx = 0; bx < bw; bx++) {
                        if (!restartIfDue()) return false;
                        if (!decodeBlock(c, &c.coef[((size_t)by * c.blocksW + bx) * 64])) return false;
                        untilRestart--;
                    }
            } else {
                for (int my = 0; my < mcuy; my++)
                    for (int mx = 0; mx < mcux; mx++) {
                        if (!restartIfDue()) return false;
                        for (int s2 = 0; s2 < ns; s2++) {
                            JpegComp& c = comp[sc[s2]];
                            for (int by = 0; by < c.v; by++)
                                for (int bx = 0; bx < c.h; bx++)
                                    if (!decodeBlock(c, &c.coef[((size_t)(my * c.v + by) * c.blocksW + (mx * c.h + bx)) * 64])) return false;
                        }
                        untilRestart--;
                    }
            }
            haveScan = true;
            pos = br.pos;                                              // at the marker that ended the entropy-coded segment
            continue;
        }
        pos += len;
    }
    if (!haveFrame || !haveScan) return false;
    // de-quantise + inverse DCT
    for (int c = 0; c < ncomp; c++) {
        JpegComp& cc = comp[c];
        cc.stride = cc.blocksW * 8; cc.hgt = cc.blocksH * 8;
        cc.data.assign((size_t)cc.stride * cc.hgt, 128);
        int coef[64];
        for (int by = 0; by < cc.blocksH; by++)
            for (int bx = 0; bx < cc.blocksW; bx++) {
                const int16_t* blk = &cc.coef[((size_t)by * cc.blocksW + bx) * 64];
                for (int k = 0; k < 64; k++) coef[k] = blk[k];
                jpegIdct(coef, qt[cc.tq], &cc.data[(size_t)by * 8 * cc.stride + (size_t)bx * 8], cc.stride);
            }
    }
    // up-sample and convert
    rgb.assign((size_t)width * height * 3, 0);
    auto sampleAt = [&](const JpegComp& c, int x, int y) -> float {
        const int fx = hmax / c.h, fy = vmax / c.v;
        if (fx * c.h != hmax || fy * c.v != vmax || (fx == 1 && fy == 1))
            return c.data[(size_t)std::min(y * c.v / vmax, c.hgt - 1) * c.stride + std::min(x * c.h / hmax, c.stride - 1)];
        // centre-aligned bilinear interpolation of the sub-sampled plane
        const int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
        const float sx = (x + 0.5f) / fx - 0.5f, sy = (y + 0.5f) / fy - 0.5f;
        int x0 = (int)std::floor(sx), y0 = (int)std::floor(sy);
        const float ax = sx - x0, ay = sy - y0;
        auto at = [&](int xx, int yy) { xx = std::min(std::max(xx, 0), cw - 1); yy = std::min(std::max(yy, 0), ch - 1); return (float)c.data[(size_t)yy * c.stride + xx]; };
        return (at(x0, y0) * (1 - ax) + at(x0 + 1, y0) * ax) * (1 - ay) + (at(x0, y0 + 1) * (1 - ax) + at(x0 + 1, y0 + 1) * ax) * ay;
    };
    auto clamp8 = [](float v) { const int i = (int)std::floor(v + 0.5f); return (unsigned char)(i < 0 ? 0 : i > 255 ? 255 : i); };
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            unsigned char* o = &rgb[((size_t)y * width + x) * 3];
            const float Y = sampleAt(comp[0], x, y);
            if (ncomp == 1) { o[0] = o[1] = o[2] = clamp8(Y); continue; }
            const float cb = sampleAt(comp[1], x, y) - 128.0f, cr = sampleAt(comp[2], x, y) - 128.0f;
            o[0] = clamp8(Y + 1.402f * cr); o[1] = clamp8(Y - 0.344136f * cb - 0.714136f * cr); o[2] = clamp8(Y + 1.772f * cb);
        }
    return true;
}

bool loadPPM(const Bytes& f, Bytes& rgb, int& width, int& height) {
    // "P6" <ws> width <ws> height <ws> maxval <single ws> data; '#' comments allowed in the header
    size_t pos = 2;
    auto number = [&](int& v) {
        while (pos < f.size()) {
            if (f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') pos++; }
            else if (std::isspace(f[pos])) pos++;
            else break;
        }
        if (pos >= f.size() || !std::isdigit(f[pos])) return false;
        v = 0;
        while (pos < f.size() && std::isdigit(f[pos])) v = v * 10 + (f[pos++] - '0');
        return true;
    };
    int maxv = 0;
    if (!number(width) || !number(height) || !number(maxv) || maxv != 255 || width <= 0 || height <= 0) return false;
    pos++;
    const size_t need = (size_t)width * height * 3;
    if (pos + need > f.size()) return false;
    rgb.assign(f.begin() + pos, f.begin() + pos + need);
    return true;
}

}  // namespace

bool loadByteImage(const std::string& path, std::vector<unsigned char>& rgb, int& width, int& height) {
    Bytes f;
    if (!readFile(path, f) || f.size() < 4) return false;
    width = height = 0;
    try {       // a corrupt header may ask for gigabytes: a failed allocation is "unreadable", not fatal
        bool ok;
        if (f[0] == 'P' && f[1] == '6') ok = loadPPM(f, rgb, width, height);
        else if (f[0] == 0x89 && f[1] == 'P') ok = loadPNG(f, rgb, width, height);
        else if (f[0] == 0xff && f[1] == 0xd8) ok = loadJPEG(f, rgb, width, height);
        else if (f[0] == 'B' && f[1] == 'M') ok = loadBMP(f, rgb, width, height);
        else ok = loadTGA(f, rgb, width, height);          // TGA has no magic number: last
        return ok && rgb.size() == (size_t)width * height * 3;
    } catch (const std::bad_alloc&) {
        return false;
    } catch (const std::length_error&) {
        return false;
    }
}

}  // namespace zillum
