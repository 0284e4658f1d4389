// depth_io.cpp -- the input side of the path: 16-bit depth images as the reference reads them.
//
// The reference loads its frames with stbi_load_16("assets/T0.png", ...) (Application.cpp:28-29; stb_image.h is a
// vendored third-party header) and hands the raw uint16 samples (TUM RGB-D convention, 5000 units per metre,
// common.h) to preProcess.  This is a self-contained reader for that format -- PNG, grey or colour, 8 or 16 bit,
// non-interlaced, all five scanline filters, CRC-checked -- built on zlib's inflate (the only dependency), plus a
// writer (fixtures, dumps) and binary PGM (P5, maxval 65535) for tools without a PNG encoder.
// stbi_load_16's conventions are kept: big-endian 16-bit samples become native uint16; 8-bit samples are widened
// as v * 257; of a multi-channel image the first channel is returned (the reference asks for the file's own
// channel count and reads the buffer as one uint16 per pixel, which is only meaningful for 1-channel files).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

#include "vh/abi.h"

namespace {

thread_local std::string g_ioError;
int iofail(const char* what) {
    g_ioError = what;
    return VH_ERR_INVALID;
}

bool readFile(const char* path, std::vector<unsigned char>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    buf.resize((size_t)n);
    bool ok = n == 0 || fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void putBe32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16)); v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}

inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

int decodePng(const std::vector<unsigned char>& file, uint16_t** out, int* w, int* h) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) return iofail("not a PNG file");
    size_t pos = 8;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = -1;
    bool haveHdr = false, ended = false;
    std::vector<unsigned char> idat;
    while (pos + 12 <= file.size() && !ended) {
        const uint32_t len = be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        if (len > file.size() - pos - 12) return iofail("PNG: truncated chunk");
        const unsigned char* data = &file[pos + 8];
        const uint32_t crc = be32(&file[pos + 8 + len]);
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4) != crc) return iofail("PNG: chunk CRC mismatch");
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return iofail("PNG: bad IHDR");
            W = be32(data); H = be32(data + 4);
            depth = data[8]; ctype = data[9];
            if (data[10] != 0 || data[11] != 0) return iofail("PNG: unknown compression / filter method");
            if (data[12] != 0) return iofail("PNG: interlaced images are not supported");
            if (!(depth == 8 || depth == 16)) return iofail("PNG: only 8- and 16-bit samples are supported");
            if (!(ctype == 0 || ctype == 2 || ctype == 4 || ctype == 6)) return iofail("PNG: palette images are not supported");
            if (W == 0 || H == 0 || W > 16384 || H > 16384) return iofail("PNG: unreasonable image size");
            haveHdr = true;
        } else if (!memcmp(type, "IDAT", 4)) {
            if (!haveHdr) return iofail("PNG: IDAT before IHDR");
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            ended = true;
        } else if (!(type[0] & 0x20)) {
            return iofail("PNG: unknown critical chunk");
        }
        pos += 12 + (size_t)len;
    }
    if (!haveHdr || !ended) return iofail("PNG: truncated file");
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : 4;
    const size_t bpp = (size_t)channels * (depth / 8), row = (size_t)W * bpp;
    std::vector<unsigned char> raw((row + 1) * H);
    uLongf got = (uLongf)raw.size();
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw.size())
        return iofail("PNG: inflate failed or size mismatch");
    std::vector<unsigned char> prev(row, 0), cur(row);
    uint16_t* img = static_cast<uint16_t*>(malloc(sizeof(uint16_t) * (size_t)W * H));
    if (!img) return iofail("out of memory");
    for (uint32_t y = 0; y < H; ++y) {
        const unsigned char* s = &raw[(row + 1) * y];
        const int ft = s[0];
        ++s;
        for (size_t i = 0; i < row; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v;
            switch (ft) {
                case 0: v = s[i]; break;
                case 1: v = s[i] + a; break;
                case 2: v = s[i] + b; break;
                case 3: v = s[i] + ((a + b) >> 1); break;
                case 4: v = s[i] + paeth(a, b, c); break;
                default: free(img); return iofail("PNG: bad scanline filter");
            }
            cur[i] = (unsigned char)v;
        }
        for (uint32_t x = 0; x < W; ++x) {
            const unsigned char* p = &cur[x * bpp];
            img[(size_t)y * W + x] = depth == 16 ? (uint16_t)((p[0] << 8) | p[1]) : (uint16_t)(p[0] * 257);
        }
        prev.swap(cur);
    }
    *out = img; *w = (int)W; *h = (int)H;
    return VH_OK;
}

// binary PGM: "P5" <w> <h> <maxval> <single whitespace> samples (big-endian when maxval > 255); '#' comments allowed
int decodePgm(const std::vector<unsigned char>& f, uint16_t** out, int* w, int* h) {
    size_t pos = 2;
    long vals[3];
    for (int k = 0; k < 3; ++k) {
        for (;;) {
            while (pos < f.size() && isspace(f[pos])) ++pos;
            if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') ++pos; continue; }
            break;
        }
        if (pos >= f.size() || !isdigit(f[pos])) return iofail("PGM: bad header");
        long v = 0;
        while (pos < f.size() && isdigit(f[pos])) { v = v * 10 + (f[pos] - '0'); if (v > 1000000) return iofail("PGM: bad header"); ++pos; }
        vals[k] = v;
    }
    if (pos >= f.size() || !isspace(f[pos])) return iofail("PGM: bad header");
    ++pos;
    const long W = vals[0], H = vals[1], maxv = vals[2];
    if (W <= 0 || H <= 0 || W > 16384 || H > 16384 || maxv <= 0 || maxv > 65535) return iofail("PGM: bad header");
    const size_t bps = maxv > 255 ? 2 : 1, need = (size_t)W * H * bps;
    if (f.size() - pos < need) return iofail("PGM: truncated file");
    uint16_t* img = static_cast<uint16_t*>(malloc(sizeof(uint16_t) * (size_t)W * H));
    if (!img) return iofail("out of memory");
    for (size_t i = 0; i < (size_t)W * H; ++i) img[i] = bps == 2 ? (uint16_t)((f[pos + 2 * i] << 8) | f[pos + 2 * i + 1]) : f[pos + i];
    *out = img; *w = (int)W; *h = (int)H;
    return VH_OK;
}

void putChunk(std::vector<unsigned char>& o, const char* type, const std::vector<unsigned char>& data) {
    putBe32(o, (uint32_t)data.size());
    const size_t start = o.size();
    o.insert(o.end(), type, type + 4);
    o.insert(o.end(), data.begin(), data.end());
    putBe32(o, (uint32_t)crc32(crc32(0L, Z_NULL, 0), &o[start], (uInt)(o.size() - start)));
}

}  // namespace

extern "C" {

const char* vh_depth_last_error(void) { return g_ioError.c_str(); }

int vh_depth_read(const char* path, uint16_t** out, int* width, int* height) {
    if (!path || !out || !width || !height) return iofail("vh_depth_read: null argument");
    std::vector<unsigned char> file;
    if (!readFile(path, file)) return iofail("vh_depth_read: cannot read file");
    if (file.size() >= 2 && file[0] == 'P' && file[1] == '5') return decodePgm(file, out, width, height);
    return decodePng(file, out, width, height);
}

void vh_depth_free(uint16_t* data) { free(data); }

// 16-bit grey PNG.  filter: 0..4 = that scanline filter on every row (all five are exercised by the tests)
int vh_depth_write_png(const char* path, const uint16_t* data, int width, int height, int filter) {
    if (!path || !data || width <= 0 || height <= 0 || filter < 0 || filter > 4) return iofail("vh_depth_write_png: bad argument");
    const size_t row = (size_t)width * 2, bpp = 2;
    std::vector<unsigned char> raw((row + 1) * height), prev(row, 0), cur(row);
    for (int y = 0; y < height; ++y) {
        for (int x = 0; x < width; ++x) { cur[2 * x] = (unsigned char)(data[(size_t)y * width + x] >> 8); cur[2 * x + 1] = (unsigned char)data[(size_t)y * width + x]; }
        unsigned char* d = &raw[(row + 1) * y];
        *d++ = (unsigned char)filter;
        for (size_t i = 0; i < row; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            const int pred = filter == 0 ? 0 : filter == 1 ? a : filter == 2 ? b : filter == 3 ? ((a + b) >> 1) : paeth(a, b, c);
            d[i] = (unsigned char)(cur[i] - pred);
        }
        prev = cur;
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(clen);
    if (compress2(z.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return iofail("vh_depth_write_png: deflate failed");
    z.resize(clen);
    std::vector<unsigned char> o = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, hdr;
    putBe32(hdr, (uint32_t)width); putBe32(hdr, (uint32_t)height);
    hdr.push_back(16); hdr.push_back(0); hdr.push_back(0); hdr.push_back(0); hdr.push_back(0);
    putChunk(o, "IHDR", hdr);
    putChunk(o, "IDAT", z);
    putChunk(o, "IEND", {});
    FILE* f = fopen(path, "wb");
    if (!f) return iofail("vh_depth_write_png: cannot open file");
    const bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
    fclose(f);
    return ok ? VH_OK : iofail("vh_depth_write_png: short write");
}

}  // extern "C"
