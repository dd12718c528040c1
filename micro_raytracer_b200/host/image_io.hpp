// image_io.hpp — image files of the front-ends (see image_io.cpp).
#pragma once
#include <string>

#include "render.hpp"

namespace mrt_host {

std::string encode_png(const Image& im);
std::string encode_ppm(const Image& im);
std::string encode_jpeg(const Image& im, int quality);
Image decode_png_rgb8(const std::string& bytes);
Image load_image_rgb8(const std::string& path);           // PNG or binary PPM, RGB8 only (parser.rs:660-672)
void save_image(const Image& im, const std::string& path); // format by extension (cli.rs:168,174)

}  // namespace mrt_host
