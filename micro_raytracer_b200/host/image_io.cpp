// image_io.cpp — the encoders/decoders the front-ends need, standing in for the `image` crate
// (Cargo.toml:18) at its three call sites OUTSIDE the hot path: saving the final image
// (src/cli.rs:168,174: format by extension), the JPEG q90 body of the HTTP answer
// (src/http.rs:121-122) and texture files (src/parser.rs:660-672: RGB8 only).  PNG via zlib,
// binary PPM, baseline JPEG (4:4:4, the standard Annex K tables scaled to the quality).
#include "image_io.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace mrt_host {

static std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(path + ": No such file or directory (os error 2)");
    return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void write_file(const std::string& path, const std::string& data) {
    std::ofstream f(path, std::ios::binary);
    if (!f || !f.write(data.data(), (std::streamsize)data.size())) throw Error(path + ": cannot write");
}
static void be32(std::string& o, uint32_t v) {
    o += (char)(v >> 24); o += (char)(v >> 16); o += (char)(v >> 8); o += (char)v;
}
static uint32_t rd32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// ----------------------------------------------------------------------------- PNG
static void png_chunk(std::string& o, const char* type, const std::string& data) {
    be32(o, (uint32_t)data.size());
    std::string body(type, 4);
    body += data;
    o += body;
    be32(o, (uint32_t)crc32(0L, reinterpret_cast<const Bytef*>(body.data()), (uInt)body.size()));
}
std::string encode_png(const Image& im) {
    std::string raw;
    raw.reserve((size_t)im.h * (im.w * 3 + 1));
    for (uint32_t y = 0; y < im.h; y++) {
        raw += '\0';  // filter type None
        raw.append(reinterpret_cast<const char*>(im.rgb.data() + (size_t)y * im.w * 3), (size_t)im.w * 3);
    }
    uLongf n = compressBound((uLong)raw.size());
    std::string z(n, '\0');
    // level 1: a path-traced image is noise to deflate — level 6 takes 4x the time of the whole 8-GPU headline render to
    // make the file 3 % smaller (the image crate's own default for PNG is its fast setting too)
    if (compress2(reinterpret_cast<Bytef*>(&z[0]), &n, reinterpret_cast<const Bytef*>(raw.data()), (uLong)raw.size(), 1) != Z_OK)
        throw Error("png: deflate failed");
    z.resize(n);
    std::string o("\x89PNG\r\n\x1a\n", 8);
    std::string ihdr;
    be32(ihdr, im.w); be32(ihdr, im.h);
    ihdr += (char)8; ihdr += (char)2; ihdr += (char)0; ihdr += (char)0; ihdr += (char)0;  // 8-bit RGB, no interlace
    png_chunk(o, "IHDR", ihdr);
    png_chunk(o, "IDAT", z);
    png_chunk(o, "IEND", "");
    return o;
}
Image decode_png_rgb8(const std::string& d) {
    if (d.size() < 8 || std::memcmp(d.data(), "\x89PNG\r\n\x1a\n", 8) != 0) throw Error("Format error decoding Png: Invalid PNG signature.");
    const unsigned char* p = reinterpret_cast<const unsigned char*>(d.data());
    size_t i = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::string idat;
    bool end = false;
    while (!end && i + 12 <= d.size()) {
        uint32_t len = rd32(p + i);
        if (i + 12 + (size_t)len > d.size()) throw Error("Format error decoding Png: unexpected end of file");
        const char* type = d.data() + i + 4;
        const unsigned char* body = p + i + 8;
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            w = rd32(body); h = rd32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.append(reinterpret_cast<const char*>(body), len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        i += 12 + (size_t)len;
    }
    if (w == 0 || h == 0) throw Error("Format error decoding Png: missing IHDR");
    if (depth != 8 || ctype != 2) throw Error("is not rgb888 image!");  // parser.rs:664
    if (interlace != 0) throw Error("png: interlaced images are not supported");
    const size_t stride = (size_t)w * 3;
    std::string raw((stride + 1) * h, '\0');
    uLongf n = (uLongf)raw.size();
    if (uncompress(reinterpret_cast<Bytef*>(&raw[0]), &n, reinterpret_cast<const Bytef*>(idat.data()), (uLong)idat.size()) != Z_OK || n != raw.size())
        throw Error("Format error decoding Png: corrupt deflate stream");
    Image im;
    im.w = w; im.h = h;
    im.rgb.resize(stride * h);
    std::vector<uint8_t> zero(stride, 0);
    for (uint32_t y = 0; y < h; y++) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(raw.data()) + (size_t)y * (stride + 1);
        uint8_t* cur = im.rgb.data() + (size_t)y * stride;
        const uint8_t* up = y ? cur - stride : zero.data();
        const int ft = src[0];
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= 3 ? cur[x - 3] : 0, b = up[x], c = x >= 3 ? up[x - 3] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw Error("Format error decoding Png: bad filter type");
            }
            cur[x] = (uint8_t)(src[1 + x] + pred);
        }
    }
    return im;
}

// ----------------------------------------------------------------------------- PPM (P6)
std::string encode_ppm(const Image& im) {
    std::string o = "P6\n" + std::to_string(im.w) + " " + std::to_string(im.h) + "\n255\n";
    o.append(reinterpret_cast<const char*>(im.rgb.data()), im.rgb.size());
    return o;
}
static Image decode_ppm(const std::string& d) {
    size_t i = 2;
    auto next_int = [&]() -> uint32_t {
        for (;;) {
            while (i < d.size() && std::isspace((unsigned char)d[i])) i++;
            if (i < d.size() && d[i] == '#') { while (i < d.size() && d[i] != '\n') i++; continue; }
            break;
        }
        uint32_t v = 0; bool any = false;
        while (i < d.size() && d[i] >= '0' && d[i] <= '9') { v = v * 10 + (uint32_t)(d[i] - '0'); i++; any = true; }
        if (!any) throw Error("Format error decoding Pnm: bad header");
        return v;
    };
    Image im;
    im.w = next_int(); im.h = next_int();
    if (next_int() != 255) throw Error("is not rgb888 image!");
    i++;  // the single whitespace after maxval
    const size_t n = (size_t)im.w * im.h * 3;
    if (d.size() < i + n) throw Error("Format error decoding Pnm: unexpected end of file");
    im.rgb.assign(d.begin() + (long)i, d.begin() + (long)(i + n));
    return im;
}

Image load_image_rgb8(const std::string& path) {
    const std::string d = read_file(path);
    if (d.size() >= 2 && d[0] == 'P' && d[1] == '6') return decode_ppm(d);
    return decode_png_rgb8(d);
}

// ----------------------------------------------------------------------------- JPEG (baseline, 4:4:4)
namespace {
const uint8_t kZig[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t kQL[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                         18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kQC[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                         99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// Annex K.3 Huffman tables: BITS (codes per length 1..16) + HUFFVAL
const uint8_t kDcLBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcCBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVal[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLVal[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcCBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcCVal[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct Huff { uint16_t code[256]; uint8_t len[256]; };
Huff make_huff(const uint8_t* bits, const uint8_t* vals) {
    Huff h{};
    uint32_t code = 0;
    size_t k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1]; i++) { h.code[vals[k]] = (uint16_t)code++; h.len[vals[k]] = (uint8_t)l; k++; }
        code <<= 1;
    }
    return h;
}
struct BitWriter {
    std::string& o;
    uint32_t acc = 0;
    int n = 0;
    void put(uint32_t code, int len) {
        acc = (acc << len) | (code & ((1u << len) - 1u));
        n += len;
        while (n >= 8) {
            const uint8_t b = (uint8_t)(acc >> (n - 8));
            o += (char)b;
            if (b == 0xFF) o += '\0';
            n -= 8;
        }
    }
    void flush() { if (n > 0) put(0x7F, 8 - n); }
};
void marker(std::string& o, uint8_t m, const std::string& body) {
    o += (char)0xFF; o += (char)m;
    const uint32_t len = (uint32_t)body.size() + 2;
    o += (char)(len >> 8); o += (char)len;
    o += body;
}
void encode_block(BitWriter& bw, const float* px /*64 level-shifted samples*/, const uint8_t* q, int& dc_prev, const Huff& hdc, const Huff& hac) {
    // separable 8x8 DCT-II
    static float C[8][8];
    static bool init = false;
    if (!init) {
        for (int u = 0; u < 8; u++)
            for (int x = 0; x < 8; x++) C[u][x] = (u == 0 ? std::sqrt(0.125f) : 0.5f) * std::cos((2 * x + 1) * u * 3.14159265358979323846f / 16.0f);
        init = true;
    }
    float tmp[64], out[64];
    for (int y = 0; y < 8; y++)
        for (int u = 0; u < 8; u++) { float s = 0; for (int x = 0; x < 8; x++) s += C[u][x] * px[y * 8 + x]; tmp[y * 8 + u] = s; }
    for (int v = 0; v < 8; v++)
        for (int u = 0; u < 8; u++) { float s = 0; for (int y = 0; y < 8; y++) s += C[v][y] * tmp[y * 8 + u]; out[v * 8 + u] = s; }
    int zz[64];
    for (int i = 0; i < 64; i++) zz[i] = (int)std::lround(out[kZig[i]] / (float)q[kZig[i]]);
    auto category = [](int v) { int a = v < 0 ? -v : v, n = 0; while (a) { n++; a >>= 1; } return n; };
    auto bits_of = [](int v, int n) { return (uint32_t)(v < 0 ? v + (1 << n) - 1 : v); };
    const int diff = zz[0] - dc_prev;
    dc_prev = zz[0];
    int n = category(diff);
    bw.put(hdc.code[n], hdc.len[n]);
    if (n) bw.put(bits_of(diff, n), n);
    int run = 0;
    for (int i = 1; i < 64; i++) {
        if (zz[i] == 0) { run++; continue; }
        while (run > 15) { bw.put(hac.code[0xF0], hac.len[0xF0]); run -= 16; }
        n = category(zz[i]);
        const int sym = (run << 4) | n;
        bw.put(hac.code[sym], hac.len[sym]);
        bw.put(bits_of(zz[i], n), n);
        run = 0;
    }
    if (run) bw.put(hac.code[0], hac.len[0]);  // EOB
}
}  // namespace

std::string encode_jpeg(const Image& im, int quality) {
    quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    uint8_t ql[64], qc[64];
    for (int i = 0; i < 64; i++) {
        int a = (kQL[i] * scale + 50) / 100, b = (kQC[i] * scale + 50) / 100;
        ql[i] = (uint8_t)(a < 1 ? 1 : a > 255 ? 255 : a);
        qc[i] = (uint8_t)(b < 1 ? 1 : b > 255 ? 255 : b);
    }
    std::string o;
    o += (char)0xFF; o += (char)0xD8;
    marker(o, 0xE0, std::string("JFIF\0\x01\x01\0\0\x01\0\x01\0\0", 14));
    for (int t = 0; t < 2; t++) {
        std::string b(1, (char)t);
        const uint8_t* q = t ? qc : ql;
        for (int i = 0; i < 64; i++) b += (char)q[kZig[i]];
        marker(o, 0xDB, b);
    }
    {
        std::string b;
        b += (char)8; b += (char)(im.h >> 8); b += (char)im.h; b += (char)(im.w >> 8); b += (char)im.w; b += (char)3;
        for (int c = 0; c < 3; c++) { b += (char)(c + 1); b += (char)0x11; b += (char)(c ? 1 : 0); }
        marker(o, 0xC0, b);
    }
    auto dht = [&](int cls, int id, const uint8_t* bits, const uint8_t* vals, size_t nvals) {
        std::string b(1, (char)((cls << 4) | id));
        b.append(reinterpret_cast<const char*>(bits), 16);
        b.append(reinterpret_cast<const char*>(vals), nvals);
        marker(o, 0xC4, b);
    };
    dht(0, 0, kDcLBits, kDcVal, 12); dht(1, 0, kAcLBits, kAcLVal, 162);
    dht(0, 1, kDcCBits, kDcVal, 12); dht(1, 1, kAcCBits, kAcCVal, 162);
    {
        std::string b;
        b += (char)3;
        for (int c = 0; c < 3; c++) { b += (char)(c + 1); b += (char)(c ? 0x11 : 0x00); }
        b += (char)0; b += (char)63; b += (char)0;
        marker(o, 0xDA, b);
    }
    const Huff hdl = make_huff(kDcLBits, kDcVal), hal = make_huff(kAcLBits, kAcLVal);
    const Huff hdc = make_huff(kDcCBits, kDcVal), hac = make_huff(kAcCBits, kAcCVal);
    BitWriter bw{o};
    int dcy = 0, dcb = 0, dcr = 0;
    float Y[64], Cb[64], Cr[64];
    for (uint32_t by = 0; by < im.h; by += 8) {
        for (uint32_t bx = 0; bx < im.w; bx += 8) {
            for (int y = 0; y < 8; y++) {
                for (int x = 0; x < 8; x++) {
                    const uint32_t sx = bx + (uint32_t)x < im.w ? bx + (uint32_t)x : im.w - 1;  // edge replication
                    const uint32_t sy = by + (uint32_t)y < im.h ? by + (uint32_t)y : im.h - 1;
                    const uint8_t* p = im.rgb.data() + ((size_t)sy * im.w + sx) * 3;
                    const float r = p[0], g = p[1], b = p[2];
                    Y[y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
                    Cb[y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
                    Cr[y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
                }
            }
            encode_block(bw, Y, ql, dcy, hdl, hal);
            encode_block(bw, Cb, qc, dcb, hdc, hac);
            encode_block(bw, Cr, qc, dcr, hdc, hac);
        }
    }
    bw.flush();
    o += (char)0xFF; o += (char)0xD9;
    return o;
}

// ≙ RgbImage::save (cli.rs:168,174): the format follows the extension
void save_image(const Image& im, const std::string& path) {
    const size_t dot = path.rfind('.');
    std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
    for (char& ch : ext) ch = (char)std::tolower((unsigned char)ch);
    if (ext == "png") write_file(path, encode_png(im));
    else if (ext == "ppm" || ext == "pnm") write_file(path, encode_ppm(im));
    else if (ext == "jpg" || ext == "jpeg") write_file(path, encode_jpeg(im, 75));  // image 0.24's default quality for save()
    else throw Error("The image format could not be determined");
}

}  // namespace mrt_host
