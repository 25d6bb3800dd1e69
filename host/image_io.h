// Image files either side of the hot path: textures and env maps in, result images out.
// The reference goes through stb_image / stb_image_write and its OpenEXR fork (src/core/texture.cpp:16-75,
// 306-373, third-party code).  These are independent readers / writers for the same formats, built on zlib:
//   in : .png (8/16-bit, all colour types, non-interlaced), .hdr (Radiance RGBE), .exr (scanline, half/float,
//        uncompressed or ZIP), .pfm, .npy (float32)
//   out: .exr (half RGB, display window = data window, like RgbaOutputFile(..., WRITE_RGB)), .png (RGBA8),
//        .hdr, .pfm, .npy
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace asuna_host {

struct ImageF {  // RGBA32F, row 0 = top
  int w = 0, h = 0;
  std::vector<float> px;  // w*h*4
};

// readImage of the reference (src/core/texture.cpp:306-339): LDR files become pow(byte/255, gamma) (alpha
// linear), HDR files are taken as they are; always four channels.
ImageF read_image(const std::string& path, float gamma);

// writeImage of the reference (src/core/texture.cpp:341-373) by extension.  LDR conversion is stb's
// hdr_to_ldr with gamma 1: byte = clamp(int(x * 255 + 0.5), 0, 255).
void write_image(const std::string& path, int w, int h, const float* rgba);

void write_npy_f32(const std::string& path, const std::vector<size_t>& shape, const float* data);
void write_pfm(const std::string& path, int w, int h, const float* rgba);

uint16_t float_to_half(float f);
float half_to_float(uint16_t h);

}  // namespace asuna_host
