// Minimal image file I/O for the sequence driver (host only): PNG decode on top of zlib (libpng is not required) and
// binary PGM read / write.  Replaces cv::imread (src/parallel_includes/main/stereo_vision.cu:661-662) and
// loadPGM / savePGM (src/common_includes/image.h:134-171) for the formats the reference's data uses.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace svb {

struct ImageU8 {
    int width = 0, height = 0, channels = 0;  // channels: 1 (gray) or 4 (BGRA)
    std::vector<uint8_t> data;
};

// 8- or 16-bit, non-interlaced PNG of colour type gray / gray+alpha / RGB / RGBA / palette.  Colour images are returned
// as BGRA (what cv::imread + cvtColor(BGR2BGRA) gives, sv.py:185-188), gray images as 1 channel.  Returns false and
// sets *err on failure.
bool read_png(const std::string &path, ImageU8 *out, std::string *err);

// P5 PGM with optional '#' comment lines (the reference's urban*.pgm carry a GIMP comment), maxval <= 255
bool read_pgm(const std::string &path, ImageU8 *out, std::string *err);
bool write_pgm(const std::string &path, const uint8_t *data, int width, int height, std::string *err);

}  // namespace svb
