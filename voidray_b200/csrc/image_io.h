// image_io.h — host-side image file decoders behind vr_image_load_rgb32f / vr_scene_add_image_texture_file /
// vr_scene_set_environment_hdri_file. Restates `image::open(path).unwrap().to_rgb32f()` as the reference uses
// it (core/texture.rs:37, voidray_common/src/environments.rs:43; image 0.24.3): 8-bit samples / 255,
// 16-bit samples / 65535, float samples passed through, grey replicated to RGB, alpha dropped, no sRGB decode.
// Pure host C++ (zlib is the only dependency); no CUDA types.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace vr {

struct DecodedImage {
    uint32_t w = 0, h = 0;
    int bits = 0;                // 8, 16 (integer samples) or 32 (float samples)
    std::vector<uint8_t> u8;     // 3*w*h when bits == 8
    std::vector<uint16_t> u16;   // 3*w*h when bits == 16
    std::vector<float> f32;      // 3*w*h when bits == 32
    const char* format = "";     // "png", "jpeg", "tiff", "bmp", "gif", "ico", "dds", "tga", "pnm", "farbfeld", "hdr", "exr"
};

// Decodes PNG, baseline/progressive JPEG, TIFF (strips and tiles; none / LZW / Deflate / PackBits), BMP (uncompressed, bit
// fields, RLE4 / RLE8), GIF (first frame), ICO / CUR, DDS (DXT1 / DXT3 / DXT5), TGA (colour-mapped / true-colour / grey, RLE),
// PNM (P1..P6), farbfeld, Radiance HDR and scan-line / tiled OpenEXR (none / RLE / ZIPS / ZIP / PIZ / PXR24 / B44 / B44A),
// recognised by their magic bytes (TGA, ICO: by extension or a plausible header).
bool decode_image_file(const char* path, DecodedImage& out, std::string& err);
// ext: lower-case file extension when known (TGA has no signature); nullptr = guess from the bytes alone.
bool decode_image_memory(const uint8_t* data, size_t size, DecodedImage& out, std::string& err, const char* ext = nullptr);

// to_rgb32f: row-major, top row first, 3 floats per pixel.
void image_to_rgb32f(const DecodedImage& img, std::vector<float>& rgb);

}  // namespace vr
