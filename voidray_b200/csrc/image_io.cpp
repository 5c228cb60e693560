// image_io.cpp — host-side image decoders (see image_io.h). The reference decodes through the `image` crate
// (image 0.24.3 with png 0.17.6, jpeg-decoder 0.2.6, tiff 0.7.3, exr 1.5.0; Cargo.lock), none of which is
// under /root/reference: the formats are restated here from their published specifications (PNG 1.2 / RFC
// 2083, ITU-T T.81, TIFF 6.0, the Radiance picture format, the OpenEXR file layout). PNG, TIFF, HDR and EXR
// decode losslessly, so the samples equal any conforming decoder's; a JPEG decoder is only defined up to
// its IDCT and colour rounding (±1–2 levels between libjpeg, stb and jpeg-decoder), stated in DESIGN.md.
#include "image_io.h"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>

namespace vr {
namespace {

struct Reader {
    const uint8_t* p;
    size_t n;
    bool big = true;
    bool has(size_t off, size_t len) const { return off <= n && len <= n - off; }
    uint8_t u8(size_t o) const { return p[o]; }
    uint16_t u16(size_t o) const { return big ? (uint16_t)(p[o] << 8 | p[o + 1]) : (uint16_t)(p[o + 1] << 8 | p[o]); }
    uint32_t u32(size_t o) const {
        return big ? ((uint32_t)p[o] << 24 | (uint32_t)p[o + 1] << 16 | (uint32_t)p[o + 2] << 8 | p[o + 3])
                   : ((uint32_t)p[o + 3] << 24 | (uint32_t)p[o + 2] << 16 | (uint32_t)p[o + 1] << 8 | p[o]);
    }
    uint64_t u64le(size_t o) const { return (uint64_t)u32le(o) | (uint64_t)u32le(o + 4) << 32; }
    uint32_t u32le(size_t o) const {
        return (uint32_t)p[o + 3] << 24 | (uint32_t)p[o + 2] << 16 | (uint32_t)p[o + 1] << 8 | p[o];
    }
};

// Decoded size sanity: refuses headers whose pixel count cannot possibly come out of a file this small (a
// corrupt or hostile header must not make the library allocate gigabytes). ratio = the codec's best case.
bool plausible_size(uint64_t w, uint64_t h, uint64_t bytes_per_pixel, size_t file_size, uint64_t ratio, const char* fmt,
                    std::string& err) {
    if (w == 0 || h == 0 || w * h > (1ull << 30)) {
        err = std::string(fmt) + ": image dimensions out of range";
        return false;
    }
    if (w * h * bytes_per_pixel > (uint64_t)file_size * ratio + 65536) {
        err = std::string(fmt) + ": header announces more pixels than the file can hold";
        return false;
    }
    return true;
}

bool zlib_inflate(const uint8_t* src, size_t n_src, uint8_t* dst, size_t n_dst, size_t* produced, std::string& err) {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) {
        err = "zlib: inflateInit failed";
        return false;
    }
    zs.next_in = const_cast<Bytef*>(src);
    zs.avail_in = (uInt)n_src;
    zs.next_out = dst;
    zs.avail_out = (uInt)n_dst;
    const int rc = inflate(&zs, Z_FINISH);
    *produced = zs.total_out;
    inflateEnd(&zs);
    // Z_BUF_ERROR with the output full is accepted: encoders may leave trailing bytes in the stream
    if (rc != Z_STREAM_END && !(rc == Z_BUF_ERROR && zs.avail_out == 0)) {
        err = "zlib: corrupt deflate stream";
        return false;
    }
    return true;
}

// ===================================================================================== PNG
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool png_unfilter(uint8_t* rows, uint32_t n_rows, size_t row_bytes, size_t bpp, std::string& err) {
    // rows: n_rows x (1 filter byte + row_bytes); reconstructed in place
    const uint8_t* prev = nullptr;
    for (uint32_t y = 0; y < n_rows; ++y) {
        uint8_t* line = rows + (size_t)y * (row_bytes + 1);
        const uint8_t ft = line[0];
        uint8_t* cur = line + 1;
        switch (ft) {
            case 0: break;
            case 1:
                for (size_t i = bpp; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
                break;
            case 2:
                if (prev)
                    for (size_t i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + prev[i]);
                break;
            case 3:
                for (size_t i = 0; i < row_bytes; ++i) {
                    const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0;
                    cur[i] = (uint8_t)(cur[i] + ((a + b) >> 1));
                }
                break;
            case 4:
                for (size_t i = 0; i < row_bytes; ++i) {
                    const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0,
                              c = (prev && i >= bpp) ? prev[i - bpp] : 0;
                    cur[i] = (uint8_t)(cur[i] + paeth(a, b, c));
                }
                break;
            default: err = "png: unknown filter type"; return false;
        }
        prev = cur;
    }
    return true;
}

bool decode_png(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, true};
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool seen_end = false;
    while (!seen_end) {
        if (!r.has(pos, 12)) {
            err = "png: truncated chunk";
            return false;
        }
        const uint32_t len = r.u32(pos);
        const uint8_t* type = data + pos + 4;
        if (!r.has(pos + 8, (size_t)len + 4)) {
            err = "png: truncated chunk data";
            return false;
        }
        const uint8_t* body = data + pos + 8;
        if ((uint32_t)crc32(crc32(0, type, 4), body, len) != r.u32(pos + 8 + len)) {
            err = "png: chunk CRC mismatch";
            return false;
        }
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) {
                err = "png: bad IHDR";
                return false;
            }
            w = r.u32(pos + 8);
            h = r.u32(pos + 12);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
            if (body[10] != 0 || body[11] != 0 || interlace > 1) {
                err = "png: unsupported compression / filter / interlace method";
                return false;
            }
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            seen_end = true;
        }
        pos += 12 + (size_t)len;
    }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "png: bad colour type"; return false;
    }
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (w == 0 || h == 0 || !depth_ok) {
        err = "png: bad header";
        return false;
    }
    if (ctype == 3 && plte.size() < 3) {
        err = "png: palette image without PLTE";
        return false;
    }
    const size_t bits_pp = (size_t)channels * depth;
    const size_t bpp = std::max<size_t>(1, bits_pp / 8);
    // pass geometry: non-interlaced = one pass over everything; Adam7 = 7 reduced images
    struct Pass {
        uint32_t x0, y0, dx, dy, pw, ph;
    };
    std::vector<Pass> passes;
    if (!interlace) {
        passes.push_back({0, 0, 1, 1, w, h});
    } else {
        static const uint32_t X0[7] = {0, 4, 0, 2, 0, 1, 0}, Y0[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1},
                              DY[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int i = 0; i < 7; ++i) {
            const uint32_t pw = w > X0[i] ? (w - X0[i] + DX[i] - 1) / DX[i] : 0, ph = h > Y0[i] ? (h - Y0[i] + DY[i] - 1) / DY[i] : 0;
            if (pw && ph) passes.push_back({X0[i], Y0[i], DX[i], DY[i], pw, ph});
        }
    }
    if (!plausible_size(w, h, std::max<size_t>(1, bits_pp / 8), idat.size(), 1100, "png", err)) return false;
    size_t raw_size = 0;
    for (const Pass& p : passes) raw_size += (size_t)p.ph * (1 + ((size_t)p.pw * bits_pp + 7) / 8);
    std::vector<uint8_t> raw(raw_size);
    size_t produced = 0;
    if (!zlib_inflate(idat.data(), idat.size(), raw.data(), raw.size(), &produced, err)) return false;
    if (produced != raw_size) {
        err = "png: image data too short";
        return false;
    }
    out.w = w;
    out.h = h;
    out.format = "png";
    out.bits = depth == 16 ? 16 : 8;
    if (out.bits == 8) out.u8.assign((size_t)3 * w * h, 0);
    else out.u16.assign((size_t)3 * w * h, 0);
    const int maxv = (1 << depth) - 1;
    size_t off = 0;
    for (const Pass& p : passes) {
        const size_t row_bytes = ((size_t)p.pw * bits_pp + 7) / 8;
        uint8_t* rows = raw.data() + off;
        if (!png_unfilter(rows, p.ph, row_bytes, bpp, err)) return false;
        for (uint32_t py = 0; py < p.ph; ++py) {
            const uint8_t* line = rows + (size_t)py * (row_bytes + 1) + 1;
            const uint32_t y = p.y0 + py * p.dy;
            for (uint32_t px = 0; px < p.pw; ++px) {
                const uint32_t x = p.x0 + px * p.dx;
                const size_t o = 3 * ((size_t)y * w + x);
                if (depth == 16) {
                    const uint8_t* s = line + (size_t)px * channels * 2;
                    auto rd = [&](int c) { return (uint16_t)(s[2 * c] << 8 | s[2 * c + 1]); };
                    if (channels <= 2) out.u16[o] = out.u16[o + 1] = out.u16[o + 2] = rd(0);
                    else
                        for (int c = 0; c < 3; ++c) out.u16[o + c] = rd(c);
                } else if (depth == 8 && ctype != 3) {
                    const uint8_t* s = line + (size_t)px * channels;
                    if (channels <= 2) out.u8[o] = out.u8[o + 1] = out.u8[o + 2] = s[0];
                    else
                        for (int c = 0; c < 3; ++c) out.u8[o + c] = s[c];
                } else {  // packed grey (scaled to 8 bits) or palette index
                    const size_t bit = (size_t)px * depth;
                    const int v = (line[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
                    if (ctype == 3) {
                        if ((size_t)3 * v + 2 >= plte.size()) {
                            err = "png: palette index out of range";
                            return false;
                        }
                        for (int c = 0; c < 3; ++c) out.u8[o + c] = plte[3 * v + c];
                    } else {
                        out.u8[o] = out.u8[o + 1] = out.u8[o + 2] = (uint8_t)(v * 255 / maxv);
                    }
                }
            }
        }
        off += (size_t)p.ph * (row_bytes + 1);
    }
    return true;
}

// ===================================================================================== JPEG (ITU-T T.81)
const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    uint8_t fast_len[512];
    uint8_t fast_sym[512];
    int32_t maxcode[18];  // per length, -1 if none
    int32_t valptr[17];
    int32_t mincode[17];
    uint8_t symbols[256];
    bool build(const uint8_t counts[16], const uint8_t* syms, int n_syms) {
        std::memset(fast_len, 0, sizeof fast_len);
        std::memcpy(symbols, syms, (size_t)n_syms);
        int code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k;
            mincode[len] = code;
            for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
                if (len <= 9) {
                    const int first = code << (9 - len), count = 1 << (9 - len);
                    if (first + count > 512) return false;
                    for (int j = 0; j < count; ++j) {
                        fast_len[first + j] = (uint8_t)len;
                        fast_sym[first + j] = syms[k];
                    }
                }
            }
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            if (code > (1 << len)) return false;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
        return true;
    }
};

struct JpegComponent {
    int id = 0, h = 1, v = 1, tq = 0;
    int td = 0, ta = 0;         // tables selected by the current scan
    int blocks_w = 0, blocks_h = 0;    // blocks covering the component (non-interleaved scan extent)
    int stride_blocks = 0, rows_blocks = 0;  // padded to whole MCUs
    int pred = 0;
    std::vector<int16_t> coef;  // stride_blocks * rows_blocks * 64, natural order
    std::vector<uint8_t> plane;  // (stride_blocks*8) x (rows_blocks*8)
};

struct JpegBits {
    const uint8_t* p;
    size_t n, pos;
    uint32_t acc = 0;
    int cnt = 0;
    int marker = 0;  // pending marker (0xD0..), 0 if none
    void fill() {
        while (cnt <= 24) {
            int b = 0;
            if (!marker && pos < n) {
                b = p[pos];
                if (b == 0xFF) {
                    int b2 = pos + 1 < n ? p[pos + 1] : 0xD9;
                    if (b2 == 0) {
                        pos += 2;
                    } else {
                        // skip fill bytes FF FF ... then a marker
                        size_t q = pos + 1;
                        while (q < n && p[q] == 0xFF) ++q;
                        marker = q < n ? p[q] : 0xD9;
                        pos = q + 1;
                        b = 0;
                    }
                } else {
                    ++pos;
                }
            }
            acc |= (uint32_t)b << (24 - cnt);
            cnt += 8;
        }
    }
    int peek(int nbits) {
        if (cnt < nbits) fill();
        return (int)(acc >> (32 - nbits));
    }
    void skip(int nbits) {
        acc <<= nbits;
        cnt -= nbits;
    }
    int get(int nbits) {
        if (nbits == 0) return 0;
        const int v = peek(nbits);
        skip(nbits);
        return v;
    }
    int bit() { return get(1); }
    void reset() {
        acc = 0;
        cnt = 0;
        marker = 0;
    }
};

inline int jpeg_extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

inline int huff_decode(JpegBits& br, const HuffTable& ht) {
    const int look = br.peek(9);
    const int len = ht.fast_len[look];
    if (len) {
        br.skip(len);
        return ht.fast_sym[look];
    }
    int code = br.peek(16);
    for (int l = 10; l <= 16; ++l) {
        const int c = code >> (16 - l);
        if (ht.maxcode[l] >= 0 && c <= ht.maxcode[l] && c >= ht.mincode[l]) {
            br.skip(l);
            return ht.symbols[ht.valptr[l] + c - ht.mincode[l]];
        }
    }
    return -1;
}

// 8x8 inverse DCT, 12-bit fixed point even/odd decomposition (the arithmetic jidctint / stb_image /
// jpeg-decoder share), level shift + clamp folded into the row pass.
inline uint8_t clamp_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
#define VR_F2F(x) ((int)((x)*4096.0 + 0.5))
#define VR_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                     \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                            \
    p2 = s2;                                                                           \
    p3 = s6;                                                                           \
    p1 = (p2 + p3) * VR_F2F(0.5411961);                                                \
    t2 = p1 + p3 * (-VR_F2F(1.847759065));                                             \
    t3 = p1 + p2 * VR_F2F(0.765366865);                                                \
    p2 = s0;                                                                           \
    p3 = s4;                                                                           \
    t0 = (p2 + p3) * 4096;                                                             \
    t1 = (p2 - p3) * 4096;                                                             \
    x0 = t0 + t3;                                                                      \
    x3 = t0 - t3;                                                                      \
    x1 = t1 + t2;                                                                      \
    x2 = t1 - t2;                                                                      \
    t0 = s7;                                                                           \
    t1 = s5;                                                                           \
    t2 = s3;                                                                           \
    t3 = s1;                                                                           \
    p3 = t0 + t2;                                                                      \
    p4 = t1 + t3;                                                                      \
    p1 = t0 + t3;                                                                      \
    p2 = t1 + t2;                                                                      \
    p5 = (p3 + p4) * VR_F2F(1.175875602);                                              \
    t0 = t0 * VR_F2F(0.298631336);                                                     \
    t1 = t1 * VR_F2F(2.053119869);                                                     \
    t2 = t2 * VR_F2F(3.072711026);                                                     \
    t3 = t3 * VR_F2F(1.501321110);                                                     \
    p1 = p5 + p1 * (-VR_F2F(0.899976223));                                             \
    p2 = p5 + p2 * (-VR_F2F(2.562915447));                                             \
    p3 = p3 * (-VR_F2F(1.961570560));                                                  \
    p4 = p4 * (-VR_F2F(0.390180644));                                                  \
    t3 += p1 + p4;                                                                     \
    t2 += p2 + p3;                                                                     \
    t1 += p2 + p4;                                                                     \
    t0 += p1 + p3;

void idct_block(const int16_t* coef, const uint16_t* q, uint8_t* dst, int stride) {
    int tmp[64];
    for (int i = 0; i < 8; ++i) {
        const int d0 = coef[i] * q[i], d1 = coef[8 + i] * q[8 + i], d2 = coef[16 + i] * q[16 + i],
                  d3 = coef[24 + i] * q[24 + i], d4 = coef[32 + i] * q[32 + i], d5 = coef[40 + i] * q[40 + i],
                  d6 = coef[48 + i] * q[48 + i], d7 = coef[56 + i] * q[56 + i];
        if (!(d1 | d2 | d3 | d4 | d5 | d6 | d7)) {
            const int dc = d0 * 4;
            for (int k = 0; k < 8; ++k) tmp[8 * k + i] = dc;
            continue;
        }
        VR_IDCT_1D(d0, d1, d2, d3, d4, d5, d6, d7)
        x0 += 512;
        x1 += 512;
        x2 += 512;
        x3 += 512;
        tmp[i] = (x0 + t3) >> 10;
        tmp[56 + i] = (x0 - t3) >> 10;
        tmp[8 + i] = (x1 + t2) >> 10;
        tmp[48 + i] = (x1 - t2) >> 10;
        tmp[16 + i] = (x2 + t1) >> 10;
        tmp[40 + i] = (x2 - t1) >> 10;
        tmp[24 + i] = (x3 + t0) >> 10;
        tmp[32 + i] = (x3 - t0) >> 10;
    }
    for (int i = 0; i < 8; ++i) {
        const int* v = tmp + 8 * i;
        uint8_t* o = dst + (size_t)i * stride;
        VR_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        const int bias = 65536 + (128 << 17);
        x0 += bias;
        x1 += bias;
        x2 += bias;
        x3 += bias;
        o[0] = clamp_u8((x0 + t3) >> 17);
        o[7] = clamp_u8((x0 - t3) >> 17);
        o[1] = clamp_u8((x1 + t2) >> 17);
        o[6] = clamp_u8((x1 - t2) >> 17);
        o[2] = clamp_u8((x2 + t1) >> 17);
        o[5] = clamp_u8((x2 - t1) >> 17);
        o[3] = clamp_u8((x3 + t0) >> 17);
        o[4] = clamp_u8((x3 - t0) >> 17);
    }
}

struct JpegDecoder {
    const uint8_t* data;
    size_t size;
    std::string& err;
    JpegDecoder(const uint8_t* d, size_t n, std::string& e) : data(d), size(n), err(e) {}
    uint16_t qt[4][64];  // natural order
    bool qt_present[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    std::vector<JpegComponent> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, mcus_w = 0, mcus_h = 0;
    bool progressive = false, have_frame = false;
    int restart_interval = 0;
    int adobe_transform = -1;
    bool jfif = false;
    uint32_t eobrun = 0;

    bool fail(const char* m) {
        err = std::string("jpeg: ") + m;
        return false;
    }

    bool parse_dqt(size_t o, size_t len) {
        size_t end = o + len;
        while (o < end) {
            const int pq = data[o] >> 4, tq = data[o] & 15;
            ++o;
            if (tq > 3 || pq > 1) return fail("bad DQT");
            if (o + (pq ? 128 : 64) > end) return fail("truncated DQT");
            for (int i = 0; i < 64; ++i) {
                qt[tq][ZIGZAG[i]] = pq ? (uint16_t)(data[o] << 8 | data[o + 1]) : data[o];
                o += pq ? 2 : 1;
            }
            qt_present[tq] = true;
        }
        return true;
    }
    bool parse_dht(size_t o, size_t len) {
        size_t end = o + len;
        while (o < end) {
            if (o + 17 > end) return fail("truncated DHT");
            const int tc = data[o] >> 4, th = data[o] & 15;
            if (tc > 1 || th > 3) return fail("bad DHT");
            const uint8_t* counts = data + o + 1;
            int n = 0;
            for (int i = 0; i < 16; ++i) n += counts[i];
            if (n > 256 || o + 17 + n > end) return fail("bad DHT size");
            if (!(tc ? ac[th] : dc[th]).build(counts, data + o + 17, n)) return fail("invalid Huffman table");
            o += 17 + n;
        }
        return true;
    }
    bool parse_sof(size_t o, size_t len, bool prog) {
        if (have_frame) return fail("multiple frames");
        if (len < 6) return fail("bad SOF");
        if (data[o] != 8) return fail("only 8-bit precision is supported");
        height = data[o + 1] << 8 | data[o + 2];
        width = data[o + 3] << 8 | data[o + 4];
        const int nf = data[o + 5];
        if (width == 0 || height == 0) return fail("zero-sized frame (DNL is not supported)");
        if (nf != 1 && nf != 3) return fail("only 1- and 3-component images are supported");
        if (!plausible_size((uint64_t)width, (uint64_t)height, (uint64_t)nf, size, 2048, "jpeg", err)) return false;
        if (len < (size_t)6 + 3 * nf) return fail("truncated SOF");
        comps.resize(nf);
        for (int i = 0; i < nf; ++i) {
            JpegComponent& c = comps[i];
            c.id = data[o + 6 + 3 * i];
            c.h = data[o + 7 + 3 * i] >> 4;
            c.v = data[o + 7 + 3 * i] & 15;
            c.tq = data[o + 8 + 3 * i];
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("bad component spec");
            hmax = std::max(hmax, c.h);
            vmax = std::max(vmax, c.v);
        }
        if (nf == 1) comps[0].h = comps[0].v = hmax = vmax = 1;  // a single component is never interleaved
        mcus_w = (width + 8 * hmax - 1) / (8 * hmax);
        mcus_h = (height + 8 * vmax - 1) / (8 * vmax);
        for (JpegComponent& c : comps) {
            const int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
            c.blocks_w = (cw + 7) / 8;
            c.blocks_h = (ch + 7) / 8;
            c.stride_blocks = mcus_w * c.h;
            c.rows_blocks = mcus_h * c.v;
            c.coef.assign((size_t)c.stride_blocks * c.rows_blocks * 64, 0);
        }
        progressive = prog;
        have_frame = true;
        return true;
    }

    bool decode_block(JpegBits& br, JpegComponent& c, int16_t* blk, int ss, int se, int ah, int al) {
        if (!progressive) {
            const int t = huff_decode(br, dc[c.td]);
            if (t < 0 || t > 15) return fail("bad DC code");
            c.pred += t ? jpeg_extend(br.get(t), t) : 0;
            blk[0] = (int16_t)c.pred;
            const HuffTable& h = ac[c.ta];
            for (int k = 1; k < 64;) {
                const int rs = huff_decode(br, h);
                if (rs < 0) return fail("bad AC code");
                const int r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r != 15) break;
                    k += 16;
                    continue;
                }
                k += r;
                if (k > 63) return fail("AC run past the block");
                blk[ZIGZAG[k]] = (int16_t)jpeg_extend(br.get(s), s);
                ++k;
            }
            return true;
        }
        if (ss == 0) {
            if (ah == 0) {  // DC first
                const int t = huff_decode(br, dc[c.td]);
                if (t < 0 || t > 15) return fail("bad DC code");
                c.pred += t ? jpeg_extend(br.get(t), t) : 0;
                blk[0] = (int16_t)(c.pred * (1 << al));
            } else if (br.bit()) {  // DC refinement
                blk[0] = (int16_t)(blk[0] | (1 << al));
            }
            return true;
        }
        const HuffTable& h = ac[c.ta];
        if (ah == 0) {  // AC first
            if (eobrun) {
                --eobrun;
                return true;
            }
            for (int k = ss; k <= se;) {
                const int rs = huff_decode(br, h);
                if (rs < 0) return fail("bad AC code");
                const int r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) {
                        eobrun = (1u << r) - 1;
                        if (r) eobrun += (uint32_t)br.get(r);
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) return fail("AC run past the block");
                    blk[ZIGZAG[k]] = (int16_t)(jpeg_extend(br.get(s), s) * (1 << al));
                    ++k;
                }
            }
            return true;
        }
        // AC refinement (T.81 G.1.2.3)
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se; ++k) {
                const int rs = huff_decode(br, h);
                if (rs < 0) return fail("bad AC code");
                int r = rs >> 4, s = rs & 15;
                if (s) {
                    s = br.bit() ? p1 : m1;
                } else if (r != 15) {
                    eobrun = 1u << r;
                    if (r) eobrun += (uint32_t)br.get(r);
                    break;
                }
                while (k <= se) {
                    int16_t* co = blk + ZIGZAG[k];
                    if (*co != 0) {
                        if (br.bit() && (*co & p1) == 0) *co = (int16_t)(*co + (*co >= 0 ? p1 : m1));
                    } else if (--r < 0) {
                        break;
                    }
                    ++k;
                }
                if (s && k <= se) blk[ZIGZAG[k]] = (int16_t)s;
            }
        }
        if (eobrun > 0) {
            for (; k <= se; ++k) {
                int16_t* co = blk + ZIGZAG[k];
                if (*co != 0 && br.bit() && (*co & p1) == 0) *co = (int16_t)(*co + (*co >= 0 ? p1 : m1));
            }
            --eobrun;
        }
        return true;
    }

    // returns the offset of the first byte after the scan's entropy-coded data (at a marker)
    bool decode_scan(size_t header, size_t len, size_t* next) {
        if (!have_frame) return fail("SOS before SOF");
        const int ns = data[header];
        if (ns < 1 || ns > (int)comps.size() || len < (size_t)4 + 2 * ns) return fail("bad SOS");
        JpegComponent* sc[4];
        for (int i = 0; i < ns; ++i) {
            const int cs = data[header + 1 + 2 * i];
            sc[i] = nullptr;
            for (JpegComponent& c : comps)
                if (c.id == cs) sc[i] = &c;
            if (!sc[i]) return fail("SOS names an unknown component");
            sc[i]->td = data[header + 2 + 2 * i] >> 4;
            sc[i]->ta = data[header + 2 + 2 * i] & 15;
            if (sc[i]->td > 3 || sc[i]->ta > 3) return fail("bad table selector");
        }
        const int ss = data[header + 1 + 2 * ns], se = data[header + 2 + 2 * ns], ah = data[header + 3 + 2 * ns] >> 4,
                  al = data[header + 3 + 2 * ns] & 15;
        if (progressive) {
            if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) return fail("bad progressive scan parameters");
        } else if (ss != 0 || se != 63 || ah != 0 || al != 0) {
            return fail("bad sequential scan parameters");
        }
        for (int i = 0; i < ns; ++i) {
            const bool need_dc = !progressive || (ss == 0 && ah == 0), need_ac = !progressive || ss > 0;
            if (need_dc && !dc[sc[i]->td].present) return fail("scan uses an undefined DC table");
            if (need_ac && !ac[sc[i]->ta].present) return fail("scan uses an undefined AC table");
            sc[i]->pred = 0;
        }
        eobrun = 0;
        JpegBits br{data, size, header + len};
        int n_units_x, n_units_y;
        if (ns == 1) {
            n_units_x = sc[0]->blocks_w;
            n_units_y = sc[0]->blocks_h;
        } else {
            n_units_x = mcus_w;
            n_units_y = mcus_h;
        }
        int until_restart = restart_interval;
        int expected_rst = 0;
        for (int uy = 0; uy < n_units_y; ++uy) {
            for (int ux = 0; ux < n_units_x; ++ux) {
                if (restart_interval && until_restart == 0) {
                    // byte-align, then the RSTn marker
                    br.acc = 0;
                    br.cnt = 0;
                    if (!br.marker) {
                        // the marker has not been reached by the bit reader yet: scan for it
                        while (br.pos + 1 < size && !(data[br.pos] == 0xFF && data[br.pos + 1] != 0 && data[br.pos + 1] != 0xFF)) ++br.pos;
                        if (br.pos + 1 >= size) return fail("missing restart marker");
                        br.marker = data[br.pos + 1];
                        br.pos += 2;
                    }
                    if (br.marker != 0xD0 + expected_rst) return fail("unexpected marker inside a scan");
                    expected_rst = (expected_rst + 1) & 7;
                    br.marker = 0;
                    for (int i = 0; i < ns; ++i) sc[i]->pred = 0;
                    eobrun = 0;
                    until_restart = restart_interval;
                }
                if (ns == 1) {
                    JpegComponent& c = *sc[0];
                    if (!decode_block(br, c, c.coef.data() + ((size_t)uy * c.stride_blocks + ux) * 64, ss, se, ah, al)) return false;
                } else {
                    for (int i = 0; i < ns; ++i) {
                        JpegComponent& c = *sc[i];
                        for (int v = 0; v < c.v; ++v)
                            for (int hh = 0; hh < c.h; ++hh) {
                                const size_t b = (size_t)(uy * c.v + v) * c.stride_blocks + (size_t)ux * c.h + hh;
                                if (!decode_block(br, c, c.coef.data() + b * 64, ss, se, ah, al)) return false;
                            }
                    }
                }
                if (restart_interval) --until_restart;
            }
        }
        // position of the next marker
        if (br.marker) {
            *next = br.pos - 2;
        } else {
            size_t q = br.pos;
            // bytes already pulled into the accumulator are entropy data; find the next marker after them
            while (q + 1 < size && !(data[q] == 0xFF && data[q + 1] != 0 && data[q + 1] != 0xFF && !(data[q + 1] >= 0xD0 && data[q + 1] <= 0xD7))) ++q;
            *next = q;
        }
        return true;
    }

    bool finish(DecodedImage& out) {
        for (JpegComponent& c : comps) {
            if (!qt_present[c.tq]) return fail("component uses an undefined quantisation table");
            const int pw = c.stride_blocks * 8;
            c.plane.assign((size_t)pw * c.rows_blocks * 8, 0);
            for (int by = 0; by < c.rows_blocks; ++by)
                for (int bx = 0; bx < c.stride_blocks; ++bx)
                    idct_block(c.coef.data() + ((size_t)by * c.stride_blocks + bx) * 64, qt[c.tq],
                               c.plane.data() + (size_t)by * 8 * pw + (size_t)bx * 8, pw);
            std::vector<int16_t>().swap(c.coef);
        }
        out.w = (uint32_t)width;
        out.h = (uint32_t)height;
        out.bits = 8;
        out.format = "jpeg";
        out.u8.assign((size_t)3 * width * height, 0);
        if (comps.size() == 1) {
            const JpegComponent& c = comps[0];
            const int pw = c.stride_blocks * 8;
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x) {
                    const uint8_t v = c.plane[(size_t)y * pw + x];
                    uint8_t* o = &out.u8[3 * ((size_t)y * width + x)];
                    o[0] = o[1] = o[2] = v;
                }
            return true;
        }
        // upsample every component to full resolution
        std::vector<uint8_t> full[3];
        for (int ci = 0; ci < 3; ++ci) {
            const JpegComponent& c = comps[ci];
            const int pw = c.stride_blocks * 8;
            const int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
            full[ci].resize((size_t)width * height);
            const int fx = hmax / c.h, fy = vmax / c.v;
            if (c.h == hmax && c.v == vmax) {
                for (int y = 0; y < height; ++y) std::memcpy(&full[ci][(size_t)y * width], &c.plane[(size_t)y * pw], (size_t)width);
            } else if (fx == 2 && hmax == 2 * c.h && (fy == 1 || fy == 2) && vmax == fy * c.v) {
                // triangle ("fancy") filter: 3/4 nearest + 1/4 next-nearest on each subsampled axis
                std::vector<int> row(cw);
                for (int y = 0; y < height; ++y) {
                    if (fy == 1) {
                        const uint8_t* s = &c.plane[(size_t)y * pw];
                        for (int x = 0; x < cw; ++x) row[x] = 4 * s[x];
                    } else {
                        const int cy = y >> 1;
                        const int oy = (y & 1) ? std::min(cy + 1, ch - 1) : std::max(cy - 1, 0);
                        const uint8_t *s0 = &c.plane[(size_t)cy * pw], *s1 = &c.plane[(size_t)oy * pw];
                        for (int x = 0; x < cw; ++x) row[x] = 3 * s0[x] + s1[x];
                    }
                    uint8_t* d = &full[ci][(size_t)y * width];
                    for (int x = 0; x < width; ++x) {
                        const int cx = x >> 1;
                        const int ox = (x & 1) ? std::min(cx + 1, cw - 1) : std::max(cx - 1, 0);
                        d[x] = (uint8_t)((3 * row[cx] + row[ox] + ((x & 1) ? 7 : 8)) >> 4);
                    }
                }
            } else {
                if (hmax % c.h || vmax % c.v) return fail("fractional sampling ratios are not supported");
                for (int y = 0; y < height; ++y)
                    for (int x = 0; x < width; ++x) full[ci][(size_t)y * width + x] = c.plane[(size_t)(y / fy) * pw + x / fx];
            }
        }
        bool ycc = true;
        if (adobe_transform >= 0) ycc = adobe_transform == 1;
        else if (!jfif && comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B') ycc = false;
        const size_t n = (size_t)width * height;
        if (!ycc) {
            for (size_t i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) out.u8[3 * i + c] = full[c][i];
            return true;
        }
        // ITU-R BT.601 full range, 20-bit fixed point
        auto fx = [](double x) { return (int)(x * (double)(1 << 20) + 0.5); };
        const int c_rv = fx(1.40200), c_gu = fx(0.34414), c_gv = fx(0.71414), c_bu = fx(1.77200);
        for (size_t i = 0; i < n; ++i) {
            const int y = (full[0][i] << 20) + (1 << 19), cb = full[1][i] - 128, cr = full[2][i] - 128;
            out.u8[3 * i + 0] = clamp_u8((y + c_rv * cr) >> 20);
            out.u8[3 * i + 1] = clamp_u8((y - c_gu * cb - c_gv * cr) >> 20);
            out.u8[3 * i + 2] = clamp_u8((y + c_bu * cb) >> 20);
        }
        return true;
    }

    bool run(DecodedImage& out) {
        size_t pos = 2;
        for (;;) {
            // next marker
            while (pos < size && data[pos] != 0xFF) ++pos;
            while (pos < size && data[pos] == 0xFF) ++pos;
            if (pos >= size) return fail("no EOI marker");
            const int m = data[pos++];
            if (m == 0xD9) break;
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (pos + 2 > size) return fail("truncated segment");
            const size_t len = (size_t)(data[pos] << 8 | data[pos + 1]);
            if (len < 2 || pos + len > size) return fail("truncated segment");
            const size_t body = pos + 2, blen = len - 2;
            switch (m) {
                case 0xDB:
                    if (!parse_dqt(body, blen)) return false;
                    break;
                case 0xC4:
                    if (!parse_dht(body, blen)) return false;
                    break;
                case 0xC0:
                case 0xC1:
                    if (!parse_sof(body, blen, false)) return false;
                    break;
                case 0xC2:
                    if (!parse_sof(body, blen, true)) return false;
                    break;
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    return fail("lossless, hierarchical and arithmetic-coded JPEG are not supported");
                case 0xDD:
                    if (blen < 2) return fail("bad DRI");
                    restart_interval = data[body] << 8 | data[body + 1];
                    break;
                case 0xE0:
                    if (blen >= 5 && !std::memcmp(data + body, "JFIF\0", 5)) jfif = true;
                    break;
                case 0xEE:
                    if (blen >= 12 && !std::memcmp(data + body, "Adobe", 5)) adobe_transform = data[body + 11];
                    break;
                case 0xDA: {
                    size_t next = 0;
                    if (!decode_scan(body, blen, &next)) return false;
                    pos = next;
                    continue;
                }
                default: break;
            }
            pos += len;
        }
        if (!have_frame) return fail("no frame header");
        return finish(out);
    }
};

bool decode_jpeg(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    JpegDecoder d(data, size, err);
    return d.run(out);
}

// ===================================================================================== TIFF 6.0
bool tiff_lzw(const uint8_t* src, size_t n, std::vector<uint8_t>& dst, size_t expect, std::string& err) {
    // MSB-first codes, 9..12 bits, ClearCode 256, EOI 257, "early change"
    struct Entry {
        int32_t prev;
        uint8_t first, last;
        uint16_t len;
    };
    static thread_local std::vector<Entry> table(4096);
    for (int i = 0; i < 256; ++i) table[i] = {-1, (uint8_t)i, (uint8_t)i, 1};
    dst.clear();
    dst.reserve(expect);
    int next = 258, width = 9, prev = -1;
    uint32_t acc = 0;
    int cnt = 0;
    size_t pos = 0;
    while (dst.size() < expect) {
        while (cnt < width && pos < n) {
            acc = (acc << 8) | src[pos++];
            cnt += 8;
        }
        if (cnt < width) break;
        const int code = (int)((acc >> (cnt - width)) & ((1u << width) - 1));
        cnt -= width;
        if (code == 257) break;
        if (code == 256) {
            next = 258;
            width = 9;
            prev = -1;
            continue;
        }
        if (prev < 0) {
            if (code > 255) {
                err = "tiff: corrupt LZW stream";
                return false;
            }
            dst.push_back((uint8_t)code);
            prev = code;
            continue;
        }
        if (code > next || (code == next && next >= 4096)) {
            err = "tiff: corrupt LZW stream";
            return false;
        }
        if (next < 4096) {  // new entry = string(prev) + first byte of string(code); string(next) starts like string(prev)
            const uint8_t tail = code < next ? table[code].first : table[prev].first;
            table[next] = {prev, table[prev].first, tail, (uint16_t)(table[prev].len + 1)};
        }
        // emit the string of `code`
        const size_t len = table[code].len, at = dst.size();
        dst.resize(at + len);
        int c = code;
        for (size_t i = len; i-- > 0;) {
            dst[at + i] = table[c].last;
            c = table[c].prev;
        }
        if (next < 4096) ++next;
        if (next + 1 >= (1 << width) && width < 12) ++width;
        prev = code;
    }
    return true;
}

bool tiff_packbits(const uint8_t* src, size_t n, std::vector<uint8_t>& dst, size_t expect) {
    dst.clear();
    dst.reserve(expect);
    size_t pos = 0;
    while (pos < n && dst.size() < expect) {
        const int8_t c = (int8_t)src[pos++];
        if (c >= 0) {
            const size_t k = (size_t)c + 1;
            if (pos + k > n) return false;
            dst.insert(dst.end(), src + pos, src + pos + k);
            pos += k;
        } else if (c != -128) {
            if (pos >= n) return false;
            dst.insert(dst.end(), (size_t)(1 - c), src[pos++]);
        }
    }
    return true;
}

bool decode_tiff(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, data[0] == 'M'};
    if (r.u16(2) != 42) {
        err = "tiff: not a classic TIFF (BigTIFF is not supported)";
        return false;
    }
    const size_t ifd = r.u32(4);
    if (!r.has(ifd, 2)) {
        err = "tiff: bad IFD offset";
        return false;
    }
    const int n_entries = r.u16(ifd);
    if (!r.has(ifd + 2, (size_t)12 * n_entries)) {
        err = "tiff: truncated IFD";
        return false;
    }
    uint32_t w = 0, h = 0, compression = 1, photometric = 2, spp = 1, rows_per_strip = 0xFFFFFFFFu, planar = 1,
             predictor = 1, sample_format = 1, tile_w = 0, tile_l = 0;
    std::vector<uint32_t> bits, offsets, counts, colormap;
    bool tiled = false;
    auto read_array = [&](size_t e, std::vector<uint32_t>& v) -> bool {
        const int type = r.u16(e + 2);
        const uint32_t n = r.u32(e + 4);
        const size_t esz = type == 3 ? 2 : (type == 4 ? 4 : (type == 1 ? 1 : 0));
        if (!esz) return false;
        size_t at = e + 8;
        if ((size_t)n * esz > 4) at = r.u32(e + 8);
        if (!r.has(at, (size_t)n * esz)) return false;
        v.resize(n);
        for (uint32_t i = 0; i < n; ++i) v[i] = esz == 2 ? r.u16(at + 2 * i) : (esz == 4 ? r.u32(at + 4 * i) : r.u8(at + i));
        return true;
    };
    for (int i = 0; i < n_entries; ++i) {
        const size_t e = ifd + 2 + (size_t)12 * i;
        const int tag = r.u16(e);
        std::vector<uint32_t> v;
        switch (tag) {
            case 256: case 257: case 259: case 262: case 277: case 278: case 284: case 317:
                if (!read_array(e, v) || v.empty()) {
                    err = "tiff: bad tag value";
                    return false;
                }
                (tag == 256 ? w : tag == 257 ? h : tag == 259 ? compression : tag == 262 ? photometric : tag == 277 ? spp
                 : tag == 278 ? rows_per_strip : tag == 284 ? planar : predictor) = v[0];
                break;
            case 258: if (!read_array(e, bits)) { err = "tiff: bad BitsPerSample"; return false; } break;
            case 273: if (!read_array(e, offsets)) { err = "tiff: bad StripOffsets"; return false; } break;
            case 279: if (!read_array(e, counts)) { err = "tiff: bad StripByteCounts"; return false; } break;
            case 320: if (!read_array(e, colormap)) { err = "tiff: bad ColorMap"; return false; } break;
            case 339: if (read_array(e, v) && !v.empty()) sample_format = v[0]; break;
            case 322: case 323:
                if (!read_array(e, v) || v.empty()) { err = "tiff: bad tile size"; return false; }
                (tag == 322 ? tile_w : tile_l) = v[0];
                tiled = true;
                break;
            case 324: if (!read_array(e, offsets)) { err = "tiff: bad TileOffsets"; return false; } tiled = true; break;
            case 325: if (!read_array(e, counts)) { err = "tiff: bad TileByteCounts"; return false; } tiled = true; break;
            default: break;
        }
    }
    if (tiled && (tile_w == 0 || tile_l == 0 || tile_w > (1u << 20) || tile_l > (1u << 20))) {
        err = "tiff: tiled file without a usable tile size";
        return false;
    }
    if (bits.empty()) bits.push_back(1);
    const uint32_t depth = bits[0];
    for (uint32_t b : bits)
        if (b != depth) {
            err = "tiff: mixed sample depths are not supported";
            return false;
        }
    const bool is_float = sample_format == 3;
    if (w == 0 || h == 0 || spp == 0 || spp > 8 || planar != 1 || offsets.empty() || offsets.size() != counts.size() ||
        (!((depth == 8 || depth == 16) && sample_format == 1) && !(depth == 32 && is_float))) {
        err = "tiff: unsupported layout (need chunky 8/16-bit integer or 32-bit float samples)";
        return false;
    }
    if (photometric > 3 || (photometric == 3 && (depth != 8 || colormap.size() < 768)) || (photometric == 2 && spp < 3)) {
        err = "tiff: unsupported photometric interpretation";
        return false;
    }
    if (predictor != 1 && (predictor != 2 || is_float)) {
        err = "tiff: unsupported predictor";
        return false;
    }
    rows_per_strip = std::min(rows_per_strip, h);
    if (rows_per_strip == 0) rows_per_strip = h;
    const size_t bps = depth / 8, row_bytes = (size_t)w * spp * bps;
    if (!plausible_size(w, h, (uint64_t)spp * bps, size, 4096, "tiff", err)) return false;
    std::vector<uint8_t> pixels(row_bytes * h), strip;
    auto decompress = [&](const uint8_t* src, size_t n_src, uint8_t* dst, size_t expect) -> bool {
        switch (compression) {
            case 1:
                if (n_src < expect) {
                    err = "tiff: short strip";
                    return false;
                }
                std::memcpy(dst, src, expect);
                return true;
            case 5:
                if (!tiff_lzw(src, n_src, strip, expect, err)) return false;
                if (strip.size() < expect) {
                    err = "tiff: short LZW strip";
                    return false;
                }
                std::memcpy(dst, strip.data(), expect);
                return true;
            case 8:
            case 32946: {
                size_t produced = 0;
                if (!zlib_inflate(src, n_src, dst, expect, &produced, err)) return false;
                if (produced < expect) {
                    err = "tiff: short Deflate strip";
                    return false;
                }
                return true;
            }
            case 32773:
                if (!tiff_packbits(src, n_src, strip, expect) || strip.size() < expect) {
                    err = "tiff: corrupt PackBits strip";
                    return false;
                }
                std::memcpy(dst, strip.data(), expect);
                return true;
            default: err = "tiff: unsupported compression scheme " + std::to_string(compression); return false;
        }
    };
    // horizontal differencing (predictor 2) runs along each row of a strip or tile, on samples in file byte order
    auto unpredict = [&](uint8_t* buf, size_t rows, size_t samples_per_row) {
        if (predictor != 2) return;
        for (size_t y = 0; y < rows; ++y) {
            uint8_t* row = buf + y * samples_per_row * bps;
            if (depth == 8) {
                for (size_t i = spp; i < samples_per_row; ++i) row[i] = (uint8_t)(row[i] + row[i - spp]);
            } else {
                const int hi_at = r.big ? 0 : 1, lo_at = 1 - hi_at;
                for (size_t i = spp; i < samples_per_row; ++i) {
                    const uint32_t prev = (uint32_t)row[2 * (i - spp) + hi_at] << 8 | row[2 * (i - spp) + lo_at];
                    const uint32_t cur = ((uint32_t)row[2 * i + hi_at] << 8 | row[2 * i + lo_at]) + prev;
                    row[2 * i + hi_at] = (uint8_t)(cur >> 8);
                    row[2 * i + lo_at] = (uint8_t)cur;
                }
            }
        }
    };
    if (!tiled) {
        for (size_t s = 0; s < offsets.size(); ++s) {
            const size_t y0 = s * rows_per_strip;
            if (y0 >= h) break;
            const size_t rows = std::min<size_t>(rows_per_strip, h - y0), expect = rows * row_bytes;
            if (!r.has(offsets[s], counts[s])) {
                err = "tiff: strip outside the file";
                return false;
            }
            uint8_t* dst = pixels.data() + y0 * row_bytes;
            if (!decompress(data + offsets[s], counts[s], dst, expect)) return false;
            unpredict(dst, rows, (size_t)w * spp);
        }
    } else {
        const size_t tiles_x = ((size_t)w + tile_w - 1) / tile_w, tiles_y = ((size_t)h + tile_l - 1) / tile_l;
        const size_t tile_row_bytes = (size_t)tile_w * spp * bps, tile_bytes = tile_row_bytes * tile_l;
        if (offsets.size() < tiles_x * tiles_y || tile_bytes > ((size_t)1 << 31)) {
            err = "tiff: tile table does not cover the image";
            return false;
        }
        std::vector<uint8_t> tile(tile_bytes);
        for (size_t t = 0; t < tiles_x * tiles_y; ++t) {
            if (!r.has(offsets[t], counts[t])) {
                err = "tiff: tile outside the file";
                return false;
            }
            if (!decompress(data + offsets[t], counts[t], tile.data(), tile_bytes)) return false;  // edge tiles are padded to full size
            unpredict(tile.data(), tile_l, (size_t)tile_w * spp);
            const size_t x0 = (t % tiles_x) * tile_w, y0 = (t / tiles_x) * tile_l;
            const size_t cw = std::min<size_t>(tile_w, w - x0), chh = std::min<size_t>(tile_l, h - y0);
            for (size_t y = 0; y < chh; ++y)
                std::memcpy(pixels.data() + (y0 + y) * row_bytes + x0 * spp * bps, tile.data() + y * tile_row_bytes, cw * spp * bps);
        }
    }
    // samples to host order
    out.w = w;
    out.h = h;
    out.format = "tiff";
    out.bits = (int)depth;
    const size_t n_px = (size_t)w * h;
    auto channel_of = [&](int c) -> int { return photometric == 2 ? c : 0; };
    if (depth == 8) {
        out.u8.resize(3 * n_px);
        for (size_t i = 0; i < n_px; ++i)
            for (int c = 0; c < 3; ++c) {
                uint8_t v = pixels[i * spp + channel_of(c)];
                if (photometric == 0) v = (uint8_t)(255 - v);
                // ColorMap holds 16-bit entries; the 8-bit image keeps the high byte
                if (photometric == 3) v = (uint8_t)(colormap[(size_t)c * 256 + pixels[i * spp]] >> 8);
                out.u8[3 * i + c] = v;
            }
    } else if (depth == 16) {
        std::vector<uint16_t> s16((size_t)w * h * spp);
        for (size_t i = 0; i < s16.size(); ++i)
            s16[i] = r.big ? (uint16_t)(pixels[2 * i] << 8 | pixels[2 * i + 1]) : (uint16_t)(pixels[2 * i + 1] << 8 | pixels[2 * i]);
        out.u16.resize(3 * n_px);
        for (size_t i = 0; i < n_px; ++i)
            for (int c = 0; c < 3; ++c) {
                uint16_t v = s16[i * spp + channel_of(c)];
                if (photometric == 0) v = (uint16_t)(65535 - v);
                out.u16[3 * i + c] = v;
            }
    } else {
        out.f32.resize(3 * n_px);
        for (size_t i = 0; i < n_px; ++i)
            for (int c = 0; c < 3; ++c) {
                const uint8_t* s = &pixels[(i * spp + channel_of(c)) * 4];
                const uint32_t u = r.big ? ((uint32_t)s[0] << 24 | (uint32_t)s[1] << 16 | (uint32_t)s[2] << 8 | s[3])
                                         : ((uint32_t)s[3] << 24 | (uint32_t)s[2] << 16 | (uint32_t)s[1] << 8 | s[0]);
                float f;
                std::memcpy(&f, &u, 4);
                out.f32[3 * i + c] = f;
            }
    }
    return true;
}

// ===================================================================================== Radiance HDR (RGBE)
bool decode_hdr(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    size_t pos = 0;
    auto read_line = [&](std::string& line) -> bool {
        line.clear();
        while (pos < size && data[pos] != '\n') line.push_back((char)data[pos++]);
        if (pos >= size) return false;
        ++pos;
        return true;
    };
    std::string line;
    if (!read_line(line)) {
        err = "hdr: truncated header";
        return false;
    }
    for (;;) {
        if (!read_line(line)) {
            err = "hdr: truncated header";
            return false;
        }
        if (line.empty()) break;
        if (line.rfind("FORMAT=", 0) == 0 && line.find("32-bit_rle_rgbe") == std::string::npos) {
            err = "hdr: only 32-bit_rle_rgbe is supported";
            return false;
        }
    }
    if (!read_line(line)) {
        err = "hdr: missing resolution line";
        return false;
    }
    int h = 0, w = 0;
    if (std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) {
        err = "hdr: unsupported orientation (need -Y h +X w)";
        return false;
    }
    if (!plausible_size((uint64_t)w, (uint64_t)h, 4, size, 64, "hdr", err)) return false;
    std::vector<uint8_t> rgbe((size_t)4 * w);
    out.w = (uint32_t)w;
    out.h = (uint32_t)h;
    out.bits = 32;
    out.format = "hdr";
    out.f32.resize((size_t)3 * w * h);
    for (int y = 0; y < h; ++y) {
        bool rle = false;
        if (w >= 8 && w < 32768 && pos + 4 <= size && data[pos] == 2 && data[pos + 1] == 2 && (data[pos + 2] << 8 | data[pos + 3]) == w) {
            rle = true;
            pos += 4;
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < w) {
                    if (pos >= size) {
                        err = "hdr: truncated scan line";
                        return false;
                    }
                    int count = data[pos++];
                    if (count > 128) {
                        count -= 128;
                        if (pos >= size || x + count > w) {
                            err = "hdr: corrupt run";
                            return false;
                        }
                        const uint8_t v = data[pos++];
                        for (int i = 0; i < count; ++i) rgbe[(size_t)4 * (x++) + c] = v;
                    } else {
                        if (count == 0 || pos + count > size || x + count > w) {
                            err = "hdr: corrupt literal";
                            return false;
                        }
                        for (int i = 0; i < count; ++i) rgbe[(size_t)4 * (x++) + c] = data[pos++];
                    }
                }
            }
        }
        if (!rle) {  // flat pixels, with the old-style (1,1,1,n) repeat of the previous pixel
            int x = 0, shift = 0;
            while (x < w) {
                if (pos + 4 > size) {
                    err = "hdr: truncated scan line";
                    return false;
                }
                const uint8_t* px = data + pos;
                pos += 4;
                if (px[0] == 1 && px[1] == 1 && px[2] == 1 && x > 0) {
                    const int count = (int)px[3] << shift;
                    for (int i = 0; i < count && x < w; ++i, ++x) std::memcpy(&rgbe[(size_t)4 * x], &rgbe[(size_t)4 * (x - 1)], 4);
                    shift += 8;
                } else {
                    std::memcpy(&rgbe[(size_t)4 * x], px, 4);
                    ++x;
                    shift = 0;
                }
            }
        }
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = &rgbe[(size_t)4 * x];
            float* o = &out.f32[3 * ((size_t)y * w + x)];
            if (p[3] == 0) {
                o[0] = o[1] = o[2] = 0.0f;
            } else {
                const float scale = std::ldexp(1.0f, (int)p[3] - (128 + 8));
                o[0] = scale * (float)p[0];
                o[1] = scale * (float)p[1];
                o[2] = scale * (float)p[2];
            }
        }
    }
    return true;
}

// ===================================================================================== OpenEXR (scan lines)
float half_to_float(uint16_t hbits) {
    const uint32_t sign = (uint32_t)(hbits & 0x8000) << 16;
    uint32_t exp = (hbits >> 10) & 0x1F, man = hbits & 0x3FF, u;
    if (exp == 0) {
        if (man == 0) {
            u = sign;
        } else {  // subnormal: normalise
            int e = -1;
            do {
                ++e;
                man <<= 1;
            } while ((man & 0x400) == 0);
            u = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 0x3FF) << 13;
        }
    } else if (exp == 31) {
        u = sign | 0x7F800000u | man << 13;
    } else {
        u = sign | (exp + 127 - 15) << 23 | man << 13;
    }
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// ---- PIZ: 16-bit Huffman + Haar-like wavelet + value LUT (OpenEXR file-format documentation)
struct PizHuffman {
    static constexpr int ENC_SIZE = (1 << 16) + 1;
    std::vector<uint8_t> length;    // per symbol
    std::vector<uint64_t> first_code;  // per length 1..58
    std::vector<uint32_t> first_index, count;
    std::vector<uint32_t> sorted;  // symbols ordered by (length, symbol)
    static constexpr int FAST_BITS = 12;
    std::vector<uint32_t> fast;  // (symbol << 6) | len, 0 = slow path

    bool unpack(const uint8_t* p, size_t n, size_t* used, uint32_t im, uint32_t iM, std::string& err) {
        length.assign(ENC_SIZE, 0);
        uint64_t acc = 0;
        int cnt = 0;
        size_t pos = 0;
        auto get = [&](int nb) -> int {
            while (cnt < nb) {
                acc = (acc << 8) | (pos < n ? p[pos] : 0);
                ++pos;
                cnt += 8;
            }
            cnt -= nb;
            return (int)((acc >> cnt) & ((1u << nb) - 1));
        };
        for (uint32_t i = im; i <= iM;) {
            if (pos > n) {
                err = "exr: truncated PIZ code table";
                return false;
            }
            const int l = get(6);
            if (l == 63) {
                const uint32_t run = (uint32_t)get(8) + 6;
                if (i + run > iM + 1) {
                    err = "exr: corrupt PIZ code table";
                    return false;
                }
                i += run;
            } else if (l >= 59) {
                const uint32_t run = (uint32_t)l - 59 + 2;
                if (i + run > iM + 1) {
                    err = "exr: corrupt PIZ code table";
                    return false;
                }
                i += run;
            } else {
                length[i++] = (uint8_t)l;
            }
        }
        *used = std::min(pos, n);
        // canonical codes: the longest codes get the numerically smallest values
        uint64_t n_len[59] = {0};
        for (uint32_t i = im; i <= iM; ++i) n_len[length[i]]++;
        first_code.assign(59, 0);
        count.assign(59, 0);
        first_index.assign(59, 0);
        uint64_t c = 0;
        for (int l = 58; l > 0; --l) {
            const uint64_t nc = (c + n_len[l]) >> 1;
            first_code[l] = c;
            count[l] = (uint32_t)n_len[l];
            c = nc;
        }
        uint32_t at = 0;
        for (int l = 1; l <= 58; ++l) {
            first_index[l] = at;
            at += count[l];
        }
        sorted.assign(at, 0);
        std::vector<uint32_t> fill(first_index);
        for (uint32_t i = im; i <= iM; ++i)
            if (length[i]) sorted[fill[length[i]]++] = i;
        fast.assign((size_t)1 << FAST_BITS, 0);
        for (int l = 1; l <= FAST_BITS; ++l)
            for (uint32_t k = 0; k < count[l]; ++k) {
                const uint64_t code = first_code[l] + k;
                const size_t base = (size_t)(code << (FAST_BITS - l)), span = (size_t)1 << (FAST_BITS - l);
                if (base + span > fast.size()) {
                    err = "exr: corrupt PIZ code lengths";
                    return false;
                }
                for (size_t j = 0; j < span; ++j) fast[base + j] = sorted[first_index[l] + k] << 6 | (uint32_t)l;
            }
        return true;
    }

    bool decode(const uint8_t* p, size_t n, uint64_t n_bits, uint32_t rlc, uint16_t* out, size_t n_out, std::string& err) {
        uint64_t bitpos = 0;
        auto peek = [&](int nb) -> uint64_t {  // nb <= 58; bits past the end read as zero
            const size_t byte = (size_t)(bitpos >> 3);
            unsigned __int128 w = 0;
            for (int i = 0; i < 9; ++i) w = (w << 8) | (byte + i < n ? p[byte + i] : 0);
            const int drop = 72 - (int)(bitpos & 7) - nb;
            return (uint64_t)(w >> drop) & ((1ull << nb) - 1);
        };
        size_t o = 0;
        while (bitpos < n_bits) {
            uint32_t sym;
            int len = 0;
            const uint32_t f = fast[(size_t)peek(FAST_BITS)];
            if (f && (f & 63) <= n_bits - bitpos) {
                sym = f >> 6;
                len = (int)(f & 63);
            } else {
                sym = 0;
                for (int l = FAST_BITS + 1; l <= 58; ++l) {
                    if (!count[l]) continue;
                    const uint64_t code = peek(l);
                    if (code >= first_code[l] && code - first_code[l] < count[l]) {
                        sym = sorted[first_index[l] + (uint32_t)(code - first_code[l])];
                        len = l;
                        break;
                    }
                }
                if (!len) {
                    err = "exr: corrupt PIZ bit stream";
                    return false;
                }
            }
            bitpos += (uint64_t)len;
            if (sym == rlc) {
                const uint32_t run = (uint32_t)peek(8);
                bitpos += 8;
                if (o == 0 || o + run > n_out) {
                    err = "exr: corrupt PIZ run";
                    return false;
                }
                const uint16_t v = out[o - 1];
                for (uint32_t i = 0; i < run; ++i) out[o++] = v;
            } else {
                if (o >= n_out) {
                    err = "exr: PIZ data overflow";
                    return false;
                }
                out[o++] = (uint16_t)sym;
            }
        }
        if (o != n_out) {
            err = "exr: PIZ data underflow";
            return false;
        }
        return true;
    }
};

inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int ls = (int16_t)l, hs = (int16_t)h;
    const int ai = ls + (hs & 1) + (hs >> 1);
    a = (uint16_t)(int16_t)ai;
    b = (uint16_t)(int16_t)(ai - hs);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xFFFF;
    const int aa = (d + bb - 0x8000) & 0xFFFF;
    b = (uint16_t)bb;
    a = (uint16_t)aa;
}

void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = std::min(nx, ny);
    int p = 1;
    while (p <= n) p <<= 1;
    p >>= 1;
    int p2 = p;
    p >>= 1;
    auto dec = [&](uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) { w14 ? wdec14(l, h, a, b) : wdec16(l, h, a, b); };
    while (p >= 1) {
        uint16_t* py = in;
        uint16_t* const ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py;
            uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                dec(*px, *p10, i00, i10);
                dec(*p01, *p11, i01, i11);
                dec(i00, i01, *px, *p01);
                dec(i10, i11, *p10, *p11);
            }
            if (nx & p) {
                uint16_t* p10 = px + oy1;
                dec(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t* px = py;
            uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                dec(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

struct ExrChannel {
    std::string name;
    int type = 0;  // 0 uint, 1 half, 2 float
    int size() const { return type == 1 ? 2 : 4; }
};

bool exr_piz_block(const uint8_t* src, size_t n, uint8_t* dst, size_t n_dst, const std::vector<ExrChannel>& ch, int w,
                   int rows, std::string& err) {
    if (n < 4) {
        err = "exr: truncated PIZ block";
        return false;
    }
    std::vector<uint8_t> bitmap(8192, 0);
    const uint16_t min_nz = (uint16_t)(src[0] | src[1] << 8), max_nz = (uint16_t)(src[2] | src[3] << 8);
    size_t pos = 4;
    if (min_nz <= max_nz) {
        if (max_nz >= 8192 || pos + (max_nz - min_nz + 1) > n) {
            err = "exr: corrupt PIZ bitmap";
            return false;
        }
        std::memcpy(&bitmap[min_nz], src + pos, (size_t)max_nz - min_nz + 1);
        pos += (size_t)max_nz - min_nz + 1;
    }
    std::vector<uint16_t> lut(65536, 0);
    uint32_t k = 0;
    for (uint32_t i = 0; i < 65536; ++i)
        if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t max_value = (uint16_t)(k - 1);
    if (pos + 4 > n) {
        err = "exr: truncated PIZ block";
        return false;
    }
    const uint32_t hlen = src[pos] | src[pos + 1] << 8 | src[pos + 2] << 16 | (uint32_t)src[pos + 3] << 24;
    pos += 4;
    if (pos + hlen > n) {
        err = "exr: truncated PIZ Huffman data";
        return false;
    }
    const size_t n_words = n_dst / 2;
    std::vector<uint16_t> tmp(n_words);
    if (hlen == 0) {
        if (n_words) {
            err = "exr: empty PIZ Huffman data";
            return false;
        }
    } else {
        if (hlen < 20) {
            err = "exr: truncated PIZ Huffman header";
            return false;
        }
        const uint8_t* hp = src + pos;
        auto rd = [&](int o) { return (uint32_t)(hp[o] | hp[o + 1] << 8 | hp[o + 2] << 16 | (uint32_t)hp[o + 3] << 24); };
        const uint32_t im = rd(0), iM = rd(4), n_bits = rd(12);
        if (im >= (uint32_t)PizHuffman::ENC_SIZE || iM >= (uint32_t)PizHuffman::ENC_SIZE || im > iM) {
            err = "exr: corrupt PIZ Huffman header";
            return false;
        }
        PizHuffman hf;
        size_t used = 0;
        if (!hf.unpack(hp + 20, hlen - 20, &used, im, iM, err)) return false;
        if ((uint64_t)n_bits > (uint64_t)8 * (hlen - 20 - used)) {
            err = "exr: PIZ bit count exceeds the data";
            return false;
        }
        if (!hf.decode(hp + 20 + used, hlen - 20 - used, n_bits, iM, tmp.data(), n_words, err)) return false;
    }
    // wavelet per channel and 16-bit component, then LUT
    size_t start = 0;
    std::vector<size_t> ch_start(ch.size());
    for (size_t c = 0; c < ch.size(); ++c) {
        ch_start[c] = start;
        const int sz = ch[c].size() / 2;
        for (int j = 0; j < sz; ++j) wav2_decode(tmp.data() + start + j, w, sz, rows, w * sz, max_value);
        start += (size_t)w * rows * sz;
    }
    for (uint16_t& v : tmp) v = lut[v];
    // interleave back to scan-line order
    uint8_t* d = dst;
    std::vector<size_t> cur(ch_start);
    for (int y = 0; y < rows; ++y)
        for (size_t c = 0; c < ch.size(); ++c) {
            const size_t words = (size_t)w * (ch[c].size() / 2);
            std::memcpy(d, tmp.data() + cur[c], words * 2);
            cur[c] += words;
            d += words * 2;
        }
    return true;
}

void exr_unpredict(std::vector<uint8_t>& buf, std::vector<uint8_t>& scratch) {
    const size_t n = buf.size();
    for (size_t i = 1; i < n; ++i) buf[i] = (uint8_t)(buf[i - 1] + buf[i] - 128);
    scratch.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0; i < n; ++i) scratch[i] = (i & 1) ? buf[half + (i >> 1)] : buf[i >> 1];
    buf.swap(scratch);
}

bool decode_exr(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, false};
    if (size < 8) {
        err = "exr: truncated file";
        return false;
    }
    const uint32_t version = r.u32le(4);
    if ((version & 0xFF) != 2 || (version & 0x800) || (version & 0x1000)) {
        err = "exr: only single-part flat images are supported (no deep data or multi-part files)";
        return false;
    }
    const bool tiled = (version & 0x200) != 0;
    uint32_t tile_w = 0, tile_h = 0;
    size_t pos = 8;
    std::vector<ExrChannel> ch;
    int compression = -1, x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    auto read_cstr = [&](std::string& s) -> bool {
        s.clear();
        while (pos < size && data[pos]) s.push_back((char)data[pos++]);
        if (pos >= size) return false;
        ++pos;
        return true;
    };
    for (;;) {
        std::string name, type;
        if (!read_cstr(name)) {
            err = "exr: truncated header";
            return false;
        }
        if (name.empty()) break;
        if (!read_cstr(type) || !r.has(pos, 4)) {
            err = "exr: truncated header";
            return false;
        }
        const uint32_t alen = r.u32le(pos);
        pos += 4;
        if (!r.has(pos, alen)) {
            err = "exr: truncated attribute";
            return false;
        }
        if (name == "channels") {
            size_t p = pos;
            const size_t end = pos + alen;
            while (p < end && data[p]) {
                ExrChannel c;
                while (p < end && data[p]) c.name.push_back((char)data[p++]);
                ++p;
                if (p + 16 > end) {
                    err = "exr: bad channel list";
                    return false;
                }
                c.type = (int)r.u32le(p);
                const uint32_t xs = r.u32le(p + 8), ys = r.u32le(p + 12);
                p += 16;
                if (c.type < 0 || c.type > 2 || xs != 1 || ys != 1) {
                    err = "exr: subsampled or unknown-typed channels are not supported";
                    return false;
                }
                ch.push_back(c);
            }
        } else if (name == "tiles" && alen >= 9) {
            tile_w = r.u32le(pos);
            tile_h = r.u32le(pos + 4);  // byte 8: level mode | rounding mode << 4; level 0 is read whatever the mode
        } else if (name == "compression" && alen >= 1) {
            compression = data[pos];
        } else if (name == "dataWindow" && alen >= 16) {
            x0 = (int)r.u32le(pos);
            y0 = (int)r.u32le(pos + 4);
            x1 = (int)r.u32le(pos + 8);
            y1 = (int)r.u32le(pos + 12);
        }
        pos += alen;
    }
    if (ch.empty() || compression < 0 || x1 < x0 || y1 < y0) {
        err = "exr: missing channels / compression / dataWindow";
        return false;
    }
    int lines_per_block;
    switch (compression) {
        case 0: case 1: case 2: lines_per_block = 1; break;
        case 3: lines_per_block = 16; break;
        case 4: lines_per_block = 32; break;
        case 5: lines_per_block = 16; break;
        case 6: case 7: lines_per_block = 32; break;
        default: err = "exr: unsupported compression " + std::to_string(compression) + " (none, RLE, ZIPS, ZIP, PIZ, PXR24, B44 and B44A are)"; return false;
    }
    if ((int64_t)x1 - x0 >= (1 << 30) || (int64_t)y1 - y0 >= (1 << 30)) {
        err = "exr: image dimensions out of range";
        return false;
    }
    const int w = x1 - x0 + 1, h = y1 - y0 + 1;
    {
        uint64_t px_bytes = 0;
        for (const ExrChannel& c : ch) px_bytes += (uint64_t)c.size();
        if (!plausible_size((uint64_t)w, (uint64_t)h, px_bytes, size, 4096, "exr", err)) return false;
    }
    // channel lookup: R, G, B (any layer-less name), else Y replicated
    int idx[3] = {-1, -1, -1};
    std::vector<size_t> ch_off(ch.size());
    for (size_t c = 0; c < ch.size(); ++c) {
        if (ch[c].name == "R") idx[0] = (int)c;
        if (ch[c].name == "G") idx[1] = (int)c;
        if (ch[c].name == "B") idx[2] = (int)c;
    }
    if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0) {
        int y = -1;
        for (size_t c = 0; c < ch.size(); ++c)
            if (ch[c].name == "Y") y = (int)c;
        if (y < 0) {
            err = "exr: need R, G, B or Y channels";
            return false;
        }
        idx[0] = idx[1] = idx[2] = y;
    }
    if (tiled && (tile_w == 0 || tile_h == 0 || tile_w > (1u << 20) || tile_h > (1u << 20))) {
        err = "exr: tiled file without a usable tile description";
        return false;
    }
    // scan-line files: blocks of whole lines; tiled files: the tiles of level 0 (they lead the offset table)
    const int tiles_x = tiled ? (int)((w + (int64_t)tile_w - 1) / tile_w) : 1;
    const int n_blocks = tiled ? tiles_x * (int)((h + (int64_t)tile_h - 1) / tile_h) : (h + lines_per_block - 1) / lines_per_block;
    if (!r.has(pos, (size_t)8 * n_blocks)) {
        err = "exr: truncated offset table";
        return false;
    }
    out.w = (uint32_t)w;
    out.h = (uint32_t)h;
    out.bits = 32;
    out.format = "exr";
    out.f32.assign((size_t)3 * w * h, 0.0f);
    std::vector<uint8_t> raw, scratch;
    for (int b = 0; b < n_blocks; ++b) {
        const uint64_t off = r.u64le(pos + (size_t)8 * b);
        const size_t head = tiled ? 20 : 8;
        if (off > size || !r.has((size_t)off, head)) {
            err = "exr: block outside the file";
            return false;
        }
        int bx = 0, by, bw = w, rows;
        if (tiled) {
            const uint32_t tx = r.u32le((size_t)off), ty = r.u32le((size_t)off + 4);
            if (r.u32le((size_t)off + 8) != 0 || r.u32le((size_t)off + 12) != 0 || tx >= (uint32_t)tiles_x ||
                (uint64_t)ty * tile_h >= (uint64_t)h) {
                err = "exr: corrupt tile header";
                return false;
            }
            bx = (int)(tx * tile_w);
            by = (int)(ty * tile_h);
            bw = std::min((int)tile_w, w - bx);
            rows = std::min((int)tile_h, h - by);
        } else {
            by = (int)r.u32le((size_t)off) - y0;
            if (by < 0 || by >= h) {
                err = "exr: corrupt block header";
                return false;
            }
            rows = std::min(lines_per_block, h - by);
        }
        const uint32_t csize = r.u32le((size_t)off + head - 4);
        if (!r.has((size_t)off + head, csize)) {
            err = "exr: corrupt block header";
            return false;
        }
        // layout of this block's lines: one channel after the other, bw samples each
        size_t line_bytes = 0;
        for (size_t c = 0; c < ch.size(); ++c) {
            ch_off[c] = line_bytes;
            line_bytes += (size_t)bw * ch[c].size();
        }
        const size_t expect = line_bytes * rows;
        const uint8_t* src = data + off + head;
        raw.resize(expect);
        if (compression == 6 || compression == 7) {
            // B44 / B44A: half channels in 4x4 blocks of 14 bytes (3 bytes for a flat block), other channels raw;
            // the block holds one channel after the other
            if (csize == expect) {
                std::memcpy(raw.data(), src, expect);  // stored uncompressed
            } else {
                const uint8_t* in = src;
                const uint8_t* const in_end = src + csize;
                for (size_t ci = 0; ci < ch.size(); ++ci) {
                    const ExrChannel& c = ch[ci];
                    if (c.type != 1) {
                        const size_t nb = (size_t)bw * 4;
                        for (int y = 0; y < rows; ++y) {
                            if (in + nb > in_end) {
                                err = "exr: short B44 block";
                                return false;
                            }
                            std::memcpy(raw.data() + (size_t)y * line_bytes + ch_off[ci], in, nb);
                            in += nb;
                        }
                        continue;
                    }
                    for (int y = 0; y < rows; y += 4)
                        for (int x = 0; x < bw; x += 4) {
                            uint16_t v[16];
                            if (in + 3 > in_end) {
                                err = "exr: short B44 block";
                                return false;
                            }
                            if (in[2] == 0xfc) {
                                uint16_t t = (uint16_t)(in[0] << 8 | in[1]);
                                t = (t & 0x8000) ? (uint16_t)(t & 0x7fff) : (uint16_t)~t;
                                for (int i = 0; i < 16; ++i) v[i] = t;
                                in += 3;
                            } else {
                                if (in + 14 > in_end) {
                                    err = "exr: short B44 block";
                                    return false;
                                }
                                const uint8_t* b = in;
                                const uint32_t shift = b[2] >> 2;
                                if (shift > 15) {  // 16-bit values: a larger shift only occurs in corrupt data
                                    err = "exr: corrupt B44 block";
                                    return false;
                                }
                                const uint32_t bias = 0x20u << shift;
                                uint32_t t[16];
                                t[0] = (uint32_t)b[0] << 8 | b[1];
                                t[4] = t[0] + (((((uint32_t)b[2] << 4) | (b[3] >> 4)) & 0x3f) << shift) - bias;
                                t[8] = t[4] + (((((uint32_t)b[3] << 2) | (b[4] >> 6)) & 0x3f) << shift) - bias;
                                t[12] = t[8] + (((uint32_t)b[4] & 0x3f) << shift) - bias;
                                t[1] = t[0] + (((uint32_t)b[5] >> 2) << shift) - bias;
                                t[5] = t[4] + (((((uint32_t)b[5] << 4) | (b[6] >> 4)) & 0x3f) << shift) - bias;
                                t[9] = t[8] + (((((uint32_t)b[6] << 2) | (b[7] >> 6)) & 0x3f) << shift) - bias;
                                t[13] = t[12] + (((uint32_t)b[7] & 0x3f) << shift) - bias;
                                t[2] = t[1] + (((uint32_t)b[8] >> 2) << shift) - bias;
                                t[6] = t[5] + (((((uint32_t)b[8] << 4) | (b[9] >> 4)) & 0x3f) << shift) - bias;
                                t[10] = t[9] + (((((uint32_t)b[9] << 2) | (b[10] >> 6)) & 0x3f) << shift) - bias;
                                t[14] = t[13] + (((uint32_t)b[10] & 0x3f) << shift) - bias;
                                t[3] = t[2] + (((uint32_t)b[11] >> 2) << shift) - bias;
                                t[7] = t[6] + (((((uint32_t)b[11] << 4) | (b[12] >> 4)) & 0x3f) << shift) - bias;
                                t[11] = t[10] + (((((uint32_t)b[12] << 2) | (b[13] >> 6)) & 0x3f) << shift) - bias;
                                t[15] = t[14] + (((uint32_t)b[13] & 0x3f) << shift) - bias;
                                for (int i = 0; i < 16; ++i) {
                                    const uint16_t u = (uint16_t)t[i];
                                    v[i] = (u & 0x8000) ? (uint16_t)(u & 0x7fff) : (uint16_t)~u;
                                }
                                in += 14;
                            }
                            for (int dy = 0; dy < 4 && y + dy < rows; ++dy)
                                for (int dx = 0; dx < 4 && x + dx < bw; ++dx) {
                                    uint8_t* o = raw.data() + (size_t)(y + dy) * line_bytes + ch_off[ci] + 2 * (size_t)(x + dx);
                                    o[0] = (uint8_t)v[4 * dy + dx];
                                    o[1] = (uint8_t)(v[4 * dy + dx] >> 8);
                                }
                        }
                }
            }
        } else if (compression == 5) {
            // PXR24: deflate over per-channel byte planes of differences; floats keep their top 24 bits
            size_t planes = 0;
            for (const ExrChannel& c : ch) planes += (size_t)bw * (c.type == 1 ? 2 : (c.type == 2 ? 3 : 4));
            scratch.resize(planes * rows);
            size_t produced = 0;
            if (!zlib_inflate(src, csize, scratch.data(), scratch.size(), &produced, err)) return false;
            if (produced != scratch.size()) {
                err = "exr: short PXR24 block";
                return false;
            }
            const uint8_t* in = scratch.data();
            uint8_t* o = raw.data();
            for (int y = 0; y < rows; ++y)
                for (const ExrChannel& c : ch) {
                    const int nb = c.type == 1 ? 2 : (c.type == 2 ? 3 : 4);
                    const uint8_t* p[4] = {in, in + bw, in + 2 * (size_t)bw, in + 3 * (size_t)bw};
                    in += (size_t)nb * bw;
                    uint32_t pixel = 0;
                    for (int x = 0; x < bw; ++x) {
                        uint32_t diff;
                        if (c.type == 1) diff = (uint32_t)p[0][x] << 8 | p[1][x];
                        else if (c.type == 2) diff = (uint32_t)p[0][x] << 24 | (uint32_t)p[1][x] << 16 | (uint32_t)p[2][x] << 8;
                        else diff = (uint32_t)p[0][x] << 24 | (uint32_t)p[1][x] << 16 | (uint32_t)p[2][x] << 8 | p[3][x];
                        pixel += diff;
                        if (c.type == 1) {
                            o[0] = (uint8_t)pixel;
                            o[1] = (uint8_t)(pixel >> 8);
                            o += 2;
                        } else {
                            o[0] = (uint8_t)pixel;
                            o[1] = (uint8_t)(pixel >> 8);
                            o[2] = (uint8_t)(pixel >> 16);
                            o[3] = (uint8_t)(pixel >> 24);
                            o += 4;
                        }
                    }
                }
        } else if (csize == expect || compression == 0) {  // stored
            if (csize < expect) {
                err = "exr: short block";
                return false;
            }
            std::memcpy(raw.data(), src, expect);
        } else if (compression == 2 || compression == 3) {
            size_t produced = 0;
            if (!zlib_inflate(src, csize, raw.data(), expect, &produced, err)) return false;
            if (produced != expect) {
                err = "exr: short ZIP block";
                return false;
            }
            exr_unpredict(raw, scratch);
        } else if (compression == 1) {
            size_t o = 0, p = 0;
            while (p < csize && o < expect) {
                const int8_t c = (int8_t)src[p++];
                if (c < 0) {
                    const size_t k = (size_t)(-c);
                    if (p + k > csize || o + k > expect) {
                        err = "exr: corrupt RLE block";
                        return false;
                    }
                    std::memcpy(&raw[o], src + p, k);
                    p += k;
                    o += k;
                } else {
                    const size_t k = (size_t)c + 1;
                    if (p >= csize || o + k > expect) {
                        err = "exr: corrupt RLE block";
                        return false;
                    }
                    std::memset(&raw[o], src[p++], k);
                    o += k;
                }
            }
            if (o != expect) {
                err = "exr: short RLE block";
                return false;
            }
            exr_unpredict(raw, scratch);
        } else {
            if (!exr_piz_block(src, csize, raw.data(), expect, ch, bw, rows, err)) return false;
        }
        for (int y = 0; y < rows; ++y)
            for (int c = 0; c < 3; ++c) {
                const ExrChannel& cc = ch[idx[c]];
                const uint8_t* s = raw.data() + (size_t)y * line_bytes + ch_off[idx[c]];
                float* o = &out.f32[3 * ((size_t)(by + y) * w + bx) + c];
                for (int x = 0; x < bw; ++x, o += 3) {
                    if (cc.type == 1) {
                        *o = half_to_float((uint16_t)(s[2 * x] | s[2 * x + 1] << 8));
                    } else {
                        const uint32_t u = s[4 * x] | s[4 * x + 1] << 8 | s[4 * x + 2] << 16 | (uint32_t)s[4 * x + 3] << 24;
                        if (cc.type == 2) std::memcpy(o, &u, 4);
                        else *o = (float)u;
                    }
                }
            }
    }
    return true;
}

}  // namespace

namespace {
bool decode_dispatch(const uint8_t* data, size_t size, const char* ext, DecodedImage& out, std::string& err);
}

bool decode_image_memory(const uint8_t* data, size_t size, DecodedImage& out, std::string& err, const char* ext) {
    try {
        return decode_dispatch(data, size, ext, out, err);
    } catch (const std::bad_alloc&) {
        err = "out of memory while decoding";
    } catch (const std::length_error&) {
        err = "image too large";
    }
    out = DecodedImage();
    return false;
}

namespace {
// ===================================================================================== GIF 87a / 89a
// First frame only, composed onto the logical screen the way image 0.24's GifDecoder does it: pixels the frame does
// not cover are (0, 0, 0, 0), the transparent index keeps its palette colour with alpha 0, and to_rgb32f drops the
// alpha. LZW: LSB-first codes, min_code_size + 1 .. 12 bits, clear = 1 << min, end = clear + 1, no early change.
bool gif_lzw(const uint8_t* src, size_t n, int min_code, std::vector<uint8_t>& dst, size_t expect) {
    struct Entry {
        int32_t prev;
        uint8_t first, last;
        uint16_t len;
    };
    std::vector<Entry> table(4096);
    const int clear = 1 << min_code, eoi = clear + 1;
    for (int i = 0; i < clear; ++i) table[i] = {-1, (uint8_t)i, (uint8_t)i, 1};
    dst.clear();
    dst.reserve(expect);
    int next = eoi + 1, width = min_code + 1, prev = -1;
    uint32_t acc = 0;
    int cnt = 0;
    size_t pos = 0;
    while (dst.size() < expect) {
        while (cnt < width && pos < n) {
            acc |= (uint32_t)src[pos++] << cnt;
            cnt += 8;
        }
        if (cnt < width) break;
        const int code = (int)(acc & ((1u << width) - 1));
        acc >>= width;
        cnt -= width;
        if (code == eoi) break;
        if (code == clear) {
            next = eoi + 1;
            width = min_code + 1;
            prev = -1;
            continue;
        }
        if (prev < 0) {
            if (code >= clear) return false;
            dst.push_back((uint8_t)code);
            prev = code;
            continue;
        }
        if (code > next || (code == next && next >= 4096)) return false;
        if (next < 4096) {
            const uint8_t tail = code < next ? table[code].first : table[prev].first;
            table[next] = {prev, table[prev].first, tail, (uint16_t)(table[prev].len + 1)};
        }
        const size_t len = table[code].len, at = dst.size();
        dst.resize(at + len);
        int c = code;
        for (size_t i = len; i-- > 0;) {
            dst[at + i] = table[c].last;
            c = table[c].prev;
        }
        if (next < 4096) {
            ++next;
            if (next == (1 << width) && width < 12) ++width;
        }
        prev = code;
    }
    if (dst.size() > expect) dst.resize(expect);
    return true;
}

bool decode_gif(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, false};
    if (size < 13) {
        err = "gif: truncated header";
        return false;
    }
    const uint32_t sw = data[6] | data[7] << 8, sh = data[8] | data[9] << 8;
    const uint8_t flags = data[10];
    size_t pos = 13;
    const uint8_t* global_pal = nullptr;
    size_t global_n = 0;
    if (flags & 0x80) {
        global_n = (size_t)1 << ((flags & 7) + 1);
        if (!r.has(pos, 3 * global_n)) {
            err = "gif: truncated colour table";
            return false;
        }
        global_pal = data + pos;
        pos += 3 * global_n;
    }
    if (!plausible_size(sw, sh, 1, size, 4096, "gif", err)) return false;
    for (;;) {
        if (pos >= size) {
            err = "gif: no image";
            return false;
        }
        const uint8_t tag = data[pos++];
        if (tag == 0x3B) {
            err = "gif: no image";
            return false;
        }
        if (tag == 0x21) {  // extension: label, then sub-blocks (the graphic control extension only matters for alpha)
            if (pos >= size) break;
            ++pos;
            for (;;) {
                if (pos >= size) break;
                const uint8_t len = data[pos++];
                if (!len) break;
                pos += len;
            }
            continue;
        }
        if (tag != 0x2C) {
            err = "gif: unknown block";
            return false;
        }
        if (!r.has(pos, 9)) break;
        const uint32_t fx = data[pos] | data[pos + 1] << 8, fy = data[pos + 2] | data[pos + 3] << 8;
        const uint32_t fw = data[pos + 4] | data[pos + 5] << 8, fh = data[pos + 6] | data[pos + 7] << 8;
        const uint8_t lf = data[pos + 8];
        pos += 9;
        const uint8_t* pal = global_pal;
        size_t pal_n = global_n;
        if (lf & 0x80) {
            pal_n = (size_t)1 << ((lf & 7) + 1);
            if (!r.has(pos, 3 * pal_n)) break;
            pal = data + pos;
            pos += 3 * pal_n;
        }
        if (!pal) {
            err = "gif: frame without a colour table";
            return false;
        }
        if (fw == 0 || fh == 0 || (uint64_t)fx + fw > sw || (uint64_t)fy + fh > sh) {
            err = "gif: frame outside the logical screen";
            return false;
        }
        if (pos >= size) break;
        const int min_code = data[pos++];
        if (min_code < 1 || min_code > 11) {
            err = "gif: bad LZW code size";
            return false;
        }
        std::vector<uint8_t> packed;
        for (;;) {
            if (pos >= size) break;
            const uint8_t len = data[pos++];
            if (!len) break;
            if (!r.has(pos, len)) {
                err = "gif: truncated image data";
                return false;
            }
            packed.insert(packed.end(), data + pos, data + pos + len);
            pos += len;
        }
        std::vector<uint8_t> idx;
        if (!gif_lzw(packed.data(), packed.size(), min_code, idx, (size_t)fw * fh) || idx.size() != (size_t)fw * fh) {
            err = "gif: corrupt LZW stream";
            return false;
        }
        out.w = sw;
        out.h = sh;
        out.bits = 8;
        out.format = "gif";
        out.u8.assign((size_t)3 * sw * sh, 0);
        // interlaced frames store rows 0, 8, 16, ..., then 4, 12, ..., then 2, 6, ..., then 1, 3, ...
        std::vector<uint32_t> row_of(fh);
        if (lf & 0x40) {
            uint32_t k = 0;
            static const int start[4] = {0, 4, 2, 1}, step[4] = {8, 8, 4, 2};
            for (int pass = 0; pass < 4; ++pass)
                for (uint32_t y = start[pass]; y < fh; y += step[pass]) row_of[k++] = y;
        } else {
            for (uint32_t y = 0; y < fh; ++y) row_of[y] = y;
        }
        for (uint32_t k = 0; k < fh; ++k) {
            const uint8_t* src = &idx[(size_t)k * fw];
            uint8_t* dst = &out.u8[3 * ((size_t)(fy + row_of[k]) * sw + fx)];
            for (uint32_t x = 0; x < fw; ++x, dst += 3) {
                const size_t i = src[x];
                if (i < pal_n) {
                    dst[0] = pal[3 * i];
                    dst[1] = pal[3 * i + 1];
                    dst[2] = pal[3 * i + 2];
                }
            }
        }
        return true;
    }
    err = "gif: truncated file";
    return false;
}

// ===================================================================================== DDS (DXT1 / DXT3 / DXT5)
// image 0.24's DdsDecoder takes exactly these three FourCCs. The colour arithmetic restates its dxt.rs (absent from
// /root/reference: a dependency of a dependency; written down from the published source, not verifiable here): 5:6:5
// endpoints widened as v * 255 / max (truncated), the interpolated colours as (2a + b + 1) / 3 and (a + b + 1) / 2,
// DXT1's "colour0 <= colour1" mode = average + black. Other DXT decoders widen and round differently: against Pillow
// the result is within 2 levels (tests/test_image_io.py). Alpha (all that separates DXT3 / DXT5 from DXT1's colour
// block) is dropped by to_rgb32f.
bool decode_dds(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, false};
    if (size < 128 || r.u32le(4) != 124) {
        err = "dds: truncated or malformed header";
        return false;
    }
    const uint32_t h = r.u32le(12), w = r.u32le(16);
    const uint32_t pf_flags = r.u32le(80);
    const uint8_t* cc = data + 84;
    int block = 0;
    if ((pf_flags & 4) && !std::memcmp(cc, "DXT1", 4)) block = 8;
    else if ((pf_flags & 4) && (!std::memcmp(cc, "DXT3", 4) || !std::memcmp(cc, "DXT5", 4))) block = 16;
    else {
        err = "dds: only DXT1, DXT3 and DXT5 are supported (as in the image crate)";
        return false;
    }
    if (!plausible_size(w, h, 1, size, 8, "dds", err)) return false;
    const size_t bw = ((size_t)w + 3) / 4, bh = ((size_t)h + 3) / 4;
    if (!r.has(128, bw * bh * (size_t)block)) {
        err = "dds: pixel data outside the file";
        return false;
    }
    out.w = w;
    out.h = h;
    out.bits = 8;
    out.format = "dds";
    out.u8.resize((size_t)3 * w * h);
    for (size_t by = 0; by < bh; ++by)
        for (size_t bx = 0; bx < bw; ++bx) {
            const uint8_t* b = data + 128 + (by * bw + bx) * (size_t)block + (block == 16 ? 8 : 0);
            const uint32_t c0 = b[0] | b[1] << 8, c1 = b[2] | b[3] << 8;
            const uint32_t table = (uint32_t)b[4] | (uint32_t)b[5] << 8 | (uint32_t)b[6] << 16 | (uint32_t)b[7] << 24;
            uint32_t col[4][3];
            const uint32_t e[2] = {c0, c1};
            for (int k = 0; k < 2; ++k) {
                col[k][0] = ((e[k] >> 11) & 0x1F) * 0xFF / 0x1F;  // enc565_decode: v * 255 / max, truncated
                col[k][1] = ((e[k] >> 5) & 0x3F) * 0xFF / 0x3F;
                col[k][2] = (e[k] & 0x1F) * 0xFF / 0x1F;
            }
            for (int c = 0; c < 3; ++c) {
                if (c0 > c1 || block == 16) {
                    col[2][c] = (col[0][c] * 2 + col[1][c] + 1) / 3;
                    col[3][c] = (col[0][c] + col[1][c] * 2 + 1) / 3;
                } else {
                    col[2][c] = (col[0][c] + col[1][c] + 1) / 2;
                    col[3][c] = 0;
                }
            }
            for (uint32_t py = 0; py < 4; ++py)
                for (uint32_t px = 0; px < 4; ++px) {
                    const size_t x = bx * 4 + px, y = by * 4 + py;
                    if (x >= w || y >= h) continue;
                    const uint32_t sel = (table >> (2 * (py * 4 + px))) & 3;
                    for (int c = 0; c < 3; ++c) out.u8[3 * (y * w + x) + c] = (uint8_t)col[sel][c];
                }
        }
    return true;
}

// ===================================================================================== BMP
// bits of a channel mask scaled to 8 bits with rounding (v * 255 / (2^n - 1))
uint8_t scale_bits(uint32_t v, int bits) {
    if (bits <= 0) return 0;
    if (bits >= 8) return (uint8_t)(v >> (bits - 8));
    const uint32_t maxv = (1u << bits) - 1;
    return (uint8_t)((v * 255u + maxv / 2) / maxv);
}

bool decode_bmp(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, false};
    if (size < 26) {
        err = "bmp: truncated header";
        return false;
    }
    const uint32_t data_off = r.u32le(10), hdr = r.u32le(14);
    int64_t w, h;
    uint32_t bpp, compression = 0, colours = 0;
    if (hdr == 12) {
        w = (int16_t)(data[18] | data[19] << 8);
        h = (int16_t)(data[20] | data[21] << 8);
        bpp = data[24] | data[25] << 8;
    } else if (hdr >= 40 && r.has(14, hdr)) {
        w = (int32_t)r.u32le(18);
        h = (int32_t)r.u32le(22);
        bpp = data[28] | data[29] << 8;
        compression = r.u32le(30);
        colours = r.u32le(46);
    } else {
        err = "bmp: unsupported header";
        return false;
    }
    const bool top_down = h < 0;
    if (top_down) h = -h;
    if (w <= 0 || h <= 0 || !(bpp == 1 || bpp == 4 || bpp == 8 || bpp == 16 || bpp == 24 || bpp == 32)) {
        err = "bmp: bad dimensions or bit depth";
        return false;
    }
    const bool rle = (compression == 1 && bpp == 8) || (compression == 2 && bpp == 4);
    if (compression != 0 && compression != 3 && compression != 6 && !rle) {
        err = "bmp: embedded JPEG / PNG compression is not supported";
        return false;
    }
    uint32_t mask[3] = {0, 0, 0};
    if (compression == 3 || compression == 6) {
        if (bpp != 16 && bpp != 32) {
            err = "bmp: bit fields need 16 or 32 bits per pixel";
            return false;
        }
        if (!r.has(54, 12)) {
            err = "bmp: truncated bit masks";
            return false;
        }
        for (int c = 0; c < 3; ++c) mask[c] = r.u32le(54 + 4 * (size_t)c);  // right after a 40-byte header, inside a larger one
    } else if (bpp == 16) {
        mask[0] = 0x7C00;
        mask[1] = 0x03E0;
        mask[2] = 0x001F;
    } else if (bpp == 32) {
        mask[0] = 0x00FF0000;
        mask[1] = 0x0000FF00;
        mask[2] = 0x000000FF;
    }
    int shift[3] = {0, 0, 0}, bits[3] = {0, 0, 0};
    for (int c = 0; c < 3; ++c)
        if (mask[c]) {
            while (!((mask[c] >> shift[c]) & 1)) ++shift[c];
            while (shift[c] + bits[c] < 32 && ((mask[c] >> (shift[c] + bits[c])) & 1)) ++bits[c];
        }
    const size_t row_bytes = (((size_t)w * bpp + 31) / 32) * 4;
    if (!plausible_size((uint64_t)w, (uint64_t)h, 1, size, rle ? 128 : 8, "bmp", err)) return false;
    if (rle ? data_off > size : !r.has(data_off, row_bytes * (size_t)h)) {
        err = "bmp: pixel data outside the file";
        return false;
    }
    std::vector<uint8_t> palette;
    if (bpp <= 8) {
        const size_t entry = hdr == 12 ? 3 : 4, n = colours ? colours : (1u << bpp);
        const size_t at = 14 + (size_t)hdr;
        if (n > 256 || !r.has(at, n * entry)) {
            err = "bmp: bad palette";
            return false;
        }
        palette.resize(3 * 256, 0);
        for (size_t i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) palette[3 * i + c] = data[at + i * entry + (2 - c)];
    }
    out.w = (uint32_t)w;
    out.h = (uint32_t)h;
    out.bits = 8;
    out.format = "bmp";
    out.u8.resize((size_t)3 * w * h);
    if (rle) {
        // RLE8 / RLE4: (count, value) runs, 0 0 = end of line, 0 1 = end of bitmap, 0 2 dx dy = skip, 0 n = n literal
        // indices padded to 16 bits. Pixels no run reaches stay black (as in image 0.24, not palette entry 0).
        std::fill(out.u8.begin(), out.u8.end(), (uint8_t)0);
        size_t pos = data_off;
        int64_t x = 0, y = 0;  // y counts stored rows (bottom-up unless top_down)
        auto put = [&](uint32_t idx) {
            if (x < w && y < h) {
                uint8_t* o = &out.u8[(size_t)3 * ((size_t)w * (size_t)(top_down ? y : h - 1 - y) + (size_t)x)];
                for (int c = 0; c < 3; ++c) o[c] = palette[3 * idx + c];
            }
            ++x;
        };
        while (pos + 1 < size && y < h) {
            const uint8_t n = data[pos], v = data[pos + 1];
            pos += 2;
            if (n) {
                for (uint32_t k = 0; k < n; ++k) put(bpp == 8 ? v : ((k & 1) ? (v & 15) : (v >> 4)));
            } else if (v == 0) {
                x = 0;
                ++y;
            } else if (v == 1) {
                break;
            } else if (v == 2) {
                if (pos + 1 >= size) break;
                x += data[pos];
                y += data[pos + 1];
                pos += 2;
            } else {
                const size_t bytes = bpp == 8 ? v : ((size_t)v + 1) / 2;
                if (!r.has(pos, bytes)) break;
                for (uint32_t k = 0; k < v; ++k) put(bpp == 8 ? data[pos + k] : ((k & 1) ? (data[pos + k / 2] & 15) : (data[pos + k / 2] >> 4)));
                pos += (bytes + 1) & ~(size_t)1;
            }
        }
        return true;
    }
    for (int64_t y = 0; y < h; ++y) {
        const uint8_t* row = data + data_off + row_bytes * (size_t)(top_down ? y : h - 1 - y);
        uint8_t* o = &out.u8[(size_t)3 * w * y];
        for (int64_t x = 0; x < w; ++x, o += 3) {
            if (bpp <= 8) {
                const size_t bit = (size_t)x * bpp;
                const uint32_t idx = (row[bit >> 3] >> (8 - bpp - (bit & 7))) & ((1u << bpp) - 1);
                for (int c = 0; c < 3; ++c) o[c] = palette[3 * idx + c];
            } else if (bpp == 24) {
                o[0] = row[3 * x + 2];
                o[1] = row[3 * x + 1];
                o[2] = row[3 * x];
            } else {
                const uint32_t px = bpp == 16 ? (uint32_t)(row[2 * x] | row[2 * x + 1] << 8) : r.u32le((size_t)(row - data) + 4 * (size_t)x);
                for (int c = 0; c < 3; ++c) o[c] = scale_bits((px & mask[c]) >> shift[c], bits[c]);
            }
        }
    }
    return true;
}

// ===================================================================================== ICO / CUR
// image 0.24's IcoDecoder: the directory entry with the largest (bits per pixel, width x height) wins, the last one
// among equals; its payload is a PNG or a BMP without the 14-byte file header whose height field counts the XOR and
// the AND bitmap. Only the colours are needed here (to_rgb32f drops the alpha the AND mask would give).
bool decode_ico(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, false};
    const uint32_t count = size >= 6 ? (uint32_t)(data[4] | data[5] << 8) : 0;
    if (count == 0 || !r.has(6, 16 * (size_t)count)) {
        err = "ico: truncated directory";
        return false;
    }
    size_t best = count - 1;
    auto score = [&](size_t i, uint32_t& bpp, uint32_t& area) {
        const uint8_t* e = data + 6 + 16 * i;
        bpp = e[6] | e[7] << 8;
        area = (uint32_t)(e[0] ? e[0] : 256) * (uint32_t)(e[1] ? e[1] : 256);
    };
    uint32_t best_bpp, best_area;
    score(best, best_bpp, best_area);
    for (size_t i = count - 1; i-- > 0;) {
        uint32_t bpp, area;
        score(i, bpp, area);
        if (bpp > best_bpp || (bpp == best_bpp && area > best_area)) {
            best = i;
            best_bpp = bpp;
            best_area = area;
        }
    }
    const uint8_t* e = data + 6 + 16 * best;
    const uint32_t bytes = r.u32le(6 + 16 * best + 8), off = r.u32le(6 + 16 * best + 12);
    (void)e;
    if (!r.has(off, bytes) || bytes < 40) {
        err = "ico: image data outside the file";
        return false;
    }
    const uint8_t* img = data + off;
    if (!std::memcmp(img, "\x89PNG\r\n\x1a\n", 8)) {
        if (!decode_png(img, bytes, out, err)) return false;
        out.format = "ico";
        return true;
    }
    // DIB: rebuild a BMP file around it (file header, halved height)
    Reader d{img, bytes, false};
    const uint32_t hdr = d.u32le(0);
    if (hdr < 40 || hdr > bytes) {
        err = "ico: unsupported bitmap header";
        return false;
    }
    const uint32_t bpp = img[14] | img[15] << 8, compression = d.u32le(16), colours = d.u32le(32);
    size_t palette_bytes = 0;
    if (bpp <= 8) palette_bytes = 4 * (size_t)(colours ? colours : (1u << bpp));
    if (compression == 3 && hdr == 40) palette_bytes += 12;
    std::vector<uint8_t> bmp(14 + (size_t)bytes);
    bmp[0] = 'B';
    bmp[1] = 'M';
    const uint32_t data_off = 14 + hdr + (uint32_t)palette_bytes;
    for (int k = 0; k < 4; ++k) bmp[10 + k] = (uint8_t)(data_off >> (8 * k));
    std::memcpy(bmp.data() + 14, img, bytes);
    const int32_t full_h = (int32_t)d.u32le(8);
    const int32_t half = full_h / 2;
    for (int k = 0; k < 4; ++k) bmp[14 + 8 + k] = (uint8_t)((uint32_t)half >> (8 * k));
    if (!decode_bmp(bmp.data(), bmp.size(), out, err)) return false;
    out.format = "ico";
    return true;
}

// ===================================================================================== TGA
bool tga_header_plausible(const uint8_t* d, size_t size) {
    if (size < 18) return false;
    const int cmap = d[1], type = d[2], bpp = d[16];
    const bool type_ok = type == 1 || type == 2 || type == 3 || type == 9 || type == 10 || type == 11;
    return cmap <= 1 && type_ok && (bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32) && (d[12] | d[13] << 8) > 0 &&
           (d[14] | d[15] << 8) > 0;
}

bool decode_tga(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    if (!tga_header_plausible(data, size)) {
        err = "tga: bad header";
        return false;
    }
    const int id_len = data[0], cmap_type = data[1], type = data[2] & 7, rle = data[2] & 8;
    const uint32_t cmap_first = data[3] | data[4] << 8, cmap_len = data[5] | data[6] << 8, cmap_bits = data[7];
    const uint32_t w = data[12] | data[13] << 8, h = data[14] | data[15] << 8, bpp = data[16], desc = data[17];
    const size_t pb = (bpp + 7) / 8;
    size_t pos = 18 + (size_t)id_len;
    auto to_rgb = [](const uint8_t* p, uint32_t bits, uint8_t* o) {
        if (bits == 8) {
            o[0] = o[1] = o[2] = p[0];
        } else if (bits == 15 || bits == 16) {
            const uint32_t v = p[0] | p[1] << 8;
            o[0] = scale_bits((v >> 10) & 31, 5);
            o[1] = scale_bits((v >> 5) & 31, 5);
            o[2] = scale_bits(v & 31, 5);
        } else {
            o[0] = p[2];
            o[1] = p[1];
            o[2] = p[0];
        }
    };
    std::vector<uint8_t> cmap;
    if (cmap_type == 1) {
        const size_t eb = (cmap_bits + 7) / 8;
        if (!(cmap_bits == 15 || cmap_bits == 16 || cmap_bits == 24 || cmap_bits == 32) || pos + cmap_len * eb > size) {
            err = "tga: bad colour map";
            return false;
        }
        cmap.resize((size_t)3 * cmap_len);
        for (uint32_t i = 0; i < cmap_len; ++i) to_rgb(data + pos + i * eb, cmap_bits, &cmap[3 * i]);
        pos += cmap_len * eb;
    }
    if (type == 1 && (cmap_type != 1 || bpp != 8)) {
        err = "tga: colour-mapped image without an 8-bit index / colour map";
        return false;
    }
    if (type == 3 && bpp != 8 && bpp != 16) {
        err = "tga: unsupported grey depth";
        return false;
    }
    if (!plausible_size(w, h, pb, size, 128, "tga", err)) return false;
    std::vector<uint8_t> px((size_t)w * h * pb);
    if (!rle) {
        if (pos + px.size() > size) {
            err = "tga: truncated pixel data";
            return false;
        }
        std::memcpy(px.data(), data + pos, px.size());
    } else {
        size_t o = 0;
        while (o < px.size()) {
            if (pos >= size) {
                err = "tga: truncated RLE data";
                return false;
            }
            const uint8_t hd = data[pos++];
            const size_t count = (size_t)(hd & 127) + 1;
            if (o + count * pb > px.size()) {
                err = "tga: RLE packet past the image";
                return false;
            }
            if (hd & 128) {
                if (pos + pb > size) {
                    err = "tga: truncated RLE data";
                    return false;
                }
                for (size_t i = 0; i < count; ++i, o += pb) std::memcpy(&px[o], data + pos, pb);
                pos += pb;
            } else {
                if (pos + count * pb > size) {
                    err = "tga: truncated RLE data";
                    return false;
                }
                std::memcpy(&px[o], data + pos, count * pb);
                pos += count * pb;
                o += count * pb;
            }
        }
    }
    out.w = w;
    out.h = h;
    out.bits = 8;
    out.format = "tga";
    out.u8.resize((size_t)3 * w * h);
    const bool top_down = desc & 0x20, right_left = desc & 0x10;
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const uint32_t sy = top_down ? y : h - 1 - y, sx = right_left ? w - 1 - x : x;
            const uint8_t* p = &px[((size_t)sy * w + sx) * pb];
            uint8_t* o = &out.u8[3 * ((size_t)y * w + x)];
            if (type == 1) {
                const uint32_t idx = p[0];
                if (idx < cmap_first || idx - cmap_first >= cmap_len) {
                    err = "tga: colour index out of range";
                    return false;
                }
                std::memcpy(o, &cmap[3 * (size_t)(idx - cmap_first)], 3);
            } else if (type == 3) {
                o[0] = o[1] = o[2] = p[0];  // 16-bit grey = grey + alpha
            } else {
                to_rgb(p, bpp, o);
            }
        }
    return true;
}

// ===================================================================================== PNM (P1..P6), farbfeld
bool decode_pnm(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    const int kind = data[1] - '0';
    size_t pos = 2;
    auto token = [&](uint32_t& v) -> bool {
        for (;;) {  // whitespace and comments
            while (pos < size && (data[pos] == ' ' || data[pos] == '\t' || data[pos] == '\r' || data[pos] == '\n')) ++pos;
            if (pos < size && data[pos] == '#') {
                while (pos < size && data[pos] != '\n') ++pos;
                continue;
            }
            break;
        }
        if (pos >= size || data[pos] < '0' || data[pos] > '9') return false;
        uint64_t x = 0;
        while (pos < size && data[pos] >= '0' && data[pos] <= '9') {
            x = x * 10 + (uint64_t)(data[pos++] - '0');
            if (x > 0xFFFFFFFFull) return false;
        }
        v = (uint32_t)x;
        return true;
    };
    uint32_t w = 0, h = 0, maxv = 1;
    if (!token(w) || !token(h) || ((kind != 1 && kind != 4) && !token(maxv)) || maxv == 0 || maxv > 65535) {
        err = "pnm: bad header";
        return false;
    }
    const bool ascii = kind <= 3, bitmap = kind == 1 || kind == 4;
    const int ch = (kind == 3 || kind == 6) ? 3 : 1;
    const bool wide = maxv > 255;
    if (!plausible_size(w, h, 1, size, bitmap ? 8 : 1, "pnm", err)) return false;
    if (!ascii) ++pos;  // the single whitespace byte after the header
    out.w = w;
    out.h = h;
    out.format = "pnm";
    out.bits = wide ? 16 : 8;
    if (wide) out.u16.resize((size_t)3 * w * h);
    else out.u8.resize((size_t)3 * w * h);
    // samples are stored as they are, like image 0.24's decoder (maxval only chooses between 8 and 16 bits)
    auto put = [&](size_t px, int c, uint32_t v) {
        if (wide) out.u16[3 * px + c] = (uint16_t)v;
        else out.u8[3 * px + c] = (uint8_t)v;
        if (ch == 1) {
            if (wide) out.u16[3 * px + 1] = out.u16[3 * px + 2] = (uint16_t)v;
            else out.u8[3 * px + 1] = out.u8[3 * px + 2] = (uint8_t)v;
        }
    };
    const size_t n_px = (size_t)w * h;
    if (ascii) {
        for (size_t px = 0; px < n_px; ++px)
            for (int c = 0; c < ch; ++c) {
                uint32_t v;
                if (bitmap) {  // digits need no separators in P1
                    while (pos < size && data[pos] != '0' && data[pos] != '1') {
                        if (data[pos] == '#') while (pos < size && data[pos] != '\n') ++pos;
                        else ++pos;
                    }
                    if (pos >= size) {
                        err = "pnm: truncated data";
                        return false;
                    }
                    v = data[pos++] == '1' ? 0u : 255u;
                } else if (!token(v) || v > maxv) {
                    err = "pnm: bad sample";
                    return false;
                }
                put(px, c, v);
            }
        return true;
    }
    if (bitmap) {
        const size_t row_bytes = ((size_t)w + 7) / 8;
        if (pos + row_bytes * h > size) {
            err = "pnm: truncated data";
            return false;
        }
        for (uint32_t y = 0; y < h; ++y)
            for (uint32_t x = 0; x < w; ++x) put((size_t)y * w + x, 0, (data[pos + row_bytes * y + (x >> 3)] >> (7 - (x & 7))) & 1 ? 0u : 255u);
        return true;
    }
    const size_t sb = wide ? 2 : 1;
    if (pos + n_px * ch * sb > size) {
        err = "pnm: truncated data";
        return false;
    }
    for (size_t px = 0; px < n_px; ++px)
        for (int c = 0; c < ch; ++c) {
            const uint8_t* p = data + pos + (px * ch + c) * sb;
            put(px, c, wide ? (uint32_t)(p[0] << 8 | p[1]) : p[0]);
        }
    return true;
}

bool decode_farbfeld(const uint8_t* data, size_t size, DecodedImage& out, std::string& err) {
    Reader r{data, size, true};
    if (size < 16) {
        err = "farbfeld: truncated header";
        return false;
    }
    const uint32_t w = r.u32(8), h = r.u32(12);
    if (!plausible_size(w, h, 8, size, 1, "farbfeld", err)) return false;
    if ((uint64_t)w * h * 8 + 16 > size) {
        err = "farbfeld: truncated data";
        return false;
    }
    out.w = w;
    out.h = h;
    out.bits = 16;
    out.format = "farbfeld";
    out.u16.resize((size_t)3 * w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i)
        for (int c = 0; c < 3; ++c) out.u16[3 * i + c] = r.u16(16 + 8 * i + 2 * (size_t)c);
    return true;
}

bool decode_dispatch(const uint8_t* data, size_t size, const char* ext, DecodedImage& out, std::string& err) {
    out = DecodedImage();
    if (size >= 8 && !std::memcmp(data, "\x89PNG\r\n\x1a\n", 8)) return decode_png(data, size, out, err);
    if (size >= 4 && data[0] == 0xFF && data[1] == 0xD8) return decode_jpeg(data, size, out, err);
    if (size >= 8 && ((data[0] == 'I' && data[1] == 'I') || (data[0] == 'M' && data[1] == 'M'))) return decode_tiff(data, size, out, err);
    if (size >= 10 && (!std::memcmp(data, "#?RADIANCE", 10) || !std::memcmp(data, "#?RGBE", 6))) return decode_hdr(data, size, out, err);
    if (size >= 4 && data[0] == 0x76 && data[1] == 0x2F && data[2] == 0x31 && data[3] == 0x01) return decode_exr(data, size, out, err);
    if (size >= 2 && data[0] == 'B' && data[1] == 'M') return decode_bmp(data, size, out, err);
    if (size >= 3 && data[0] == 'P' && data[1] >= '1' && data[1] <= '6' && (data[2] == ' ' || data[2] == '\n' || data[2] == '\r' || data[2] == '\t' || data[2] == '#'))
        return decode_pnm(data, size, out, err);
    if (size >= 8 && !std::memcmp(data, "farbfeld", 8)) return decode_farbfeld(data, size, out, err);
    if (size >= 6 && (!std::memcmp(data, "GIF87a", 6) || !std::memcmp(data, "GIF89a", 6))) return decode_gif(data, size, out, err);
    if (size >= 4 && !std::memcmp(data, "DDS ", 4)) return decode_dds(data, size, out, err);
    // ICO / CUR: reserved 0, type 1 or 2, then the directory; no magic beyond that, so the extension or a sane count decides
    if (size >= 6 && data[0] == 0 && data[1] == 0 && (data[2] == 1 || data[2] == 2) && data[3] == 0 &&
        ((ext && (!std::strcmp(ext, "ico") || !std::strcmp(ext, "cur"))) || (!ext && data[4] != 0 && data[5] == 0 && !tga_header_plausible(data, size))))
        return decode_ico(data, size, out, err);
    // TGA has no signature: the file extension decides (as in image::open), else a plausible header
    if ((ext && !std::strcmp(ext, "tga")) || (!ext && tga_header_plausible(data, size))) return decode_tga(data, size, out, err);
    err = "unrecognised image format (PNG, JPEG, TIFF, BMP, GIF, ICO, DDS, TGA, PNM, farbfeld, Radiance HDR and OpenEXR are supported)";
    return false;
}
}  // namespace

bool decode_image_file(const char* path, DecodedImage& out, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        err = "cannot open file";
        return false;
    }
    std::vector<uint8_t> buf;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0 || n > (1L << 40)) {  // a directory reports LONG_MAX
        std::fclose(f);
        err = "not a regular file";
        return false;
    }
    buf.resize((size_t)n);
    const size_t got = n ? std::fread(buf.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    if (got != (size_t)n) {
        err = "short read";
        return false;
    }
    std::string ext;
    if (const char* dot = std::strrchr(path, '.'))
        for (const char* c = dot + 1; *c; ++c) ext.push_back((char)std::tolower((unsigned char)*c));
    return decode_image_memory(buf.data(), buf.size(), out, err, ext.c_str());
}

void image_to_rgb32f(const DecodedImage& img, std::vector<float>& rgb) {
    const size_t n = (size_t)3 * img.w * img.h;
    rgb.resize(n);
    if (img.bits == 8)
        for (size_t i = 0; i < n; ++i) rgb[i] = (float)img.u8[i] / 255.0f;
    else if (img.bits == 16)
        for (size_t i = 0; i < n; ++i) rgb[i] = (float)img.u16[i] / 65535.0f;
    else
        rgb = img.f32;
}

}  // namespace vr
