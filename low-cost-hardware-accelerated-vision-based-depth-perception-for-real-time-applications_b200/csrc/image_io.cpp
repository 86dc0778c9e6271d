#include "image_io.h"

#include <exception>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

namespace svb {

namespace {

bool slurp(const std::string &path, std::vector<uint8_t> *buf, std::string *err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open " + path;
        return false;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf->resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(buf->data(), 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != buf->size()) {
        *err = "short read on " + path;
        return false;
    }
    return true;
}

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline int paeth(int a, int b, int c) {
    const int p = a + b - c;
    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

}  // namespace

bool read_png(const std::string &path, ImageU8 *out, std::string *err) {
    std::vector<uint8_t> file;
    if (!slurp(path, &file, err)) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) {
        *err = path + ": not a PNG file";
        return false;
    }
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    size_t pos = 8;
    bool seen_end = false;
    while (pos + 12 <= file.size() && !seen_end) {
        const uint32_t len = be32(&file[pos]);
        const char *type = (const char *)&file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) {
            *err = path + ": truncated chunk";
            return false;
        }
        const uint8_t *body = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            if (W != 0) {
                *err = path + ": duplicate IHDR chunk";
                return false;
            }
            W = be32(body);
            H = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            seen_end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (W == 0 || H == 0 || W > 16384 || H > 16384) {
        *err = path + ": bad IHDR (sides of 1 .. 16384 pixels are accepted)";
        return false;
    }
    if (interlace != 0 || !(depth == 8 || depth == 16) || !(ctype == 0 || ctype == 2 || ctype == 3 || ctype == 4 || ctype == 6) ||
        (ctype == 3 && depth != 8)) {
        *err = path + ": unsupported PNG flavour (interlaced, or bit depth / colour type not handled)";
        return false;
    }
    const int samples = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
    const int bps = depth / 8;
    const size_t bpp = (size_t)samples * bps;   // bytes per pixel in the stream
    const size_t stride = (size_t)W * bpp;       // bytes per scanline without the filter byte
    // deflate expands by at most ~1032x: a header that promises more than the IDAT payload can hold is rejected BEFORE anything
    // of that size is allocated (a crafted 100-byte file must not ask for gigabytes)
    if ((stride + 1) * (size_t)H > idat.size() * 1032 + 1024) {
        *err = path + ": IDAT payload too small for the image size in IHDR";
        return false;
    }
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf raw_len = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || raw_len != raw.size()) {
        *err = path + ": zlib inflate failed";
        return false;
    }
    // undo the scanline filters in place (rows are stored with a leading filter-type byte)
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    const bool gray = (ctype == 0 || ctype == 4);
    out->width = (int)W;
    out->height = (int)H;
    out->channels = gray ? 1 : 4;
    out->data.assign((size_t)W * H * out->channels, 255);
    for (uint32_t y = 0; y < H; y++) {
        const uint8_t *src = &raw[(size_t)y * (stride + 1)];
        const int ft = src[0];
        src++;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = src[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default:
                    *err = path + ": bad filter type";
                    return false;
            }
            cur[i] = (uint8_t)v;
        }
        uint8_t *dst = &out->data[(size_t)y * W * out->channels];
        for (uint32_t x = 0; x < W; x++) {
            const uint8_t *px = &cur[(size_t)x * bpp];  // 16-bit samples: the high byte comes first
            if (gray) {
                dst[x] = px[0];
            } else if (ctype == 3) {
                const size_t k = (size_t)px[0] * 3;
                if (k + 2 < plte.size()) {
                    dst[4 * x + 0] = plte[k + 2];
                    dst[4 * x + 1] = plte[k + 1];
                    dst[4 * x + 2] = plte[k + 0];
                }
            } else {
                dst[4 * x + 0] = px[2 * bps];  // B
                dst[4 * x + 1] = px[1 * bps];  // G
                dst[4 * x + 2] = px[0];        // R
                dst[4 * x + 3] = ctype == 6 ? px[3 * bps] : 255;
            }
        }
        prev.swap(cur);
    }
    return true;
}

bool read_pgm(const std::string &path, ImageU8 *out, std::string *err) {
    std::vector<uint8_t> file;
    if (!slurp(path, &file, err)) return false;
    size_t pos = 0;
    auto token = [&](std::string *tok) {
        tok->clear();
        while (pos < file.size()) {  // skip white space and comment lines
            if (file[pos] == '#') {
                while (pos < file.size() && file[pos] != '\n') pos++;
            } else if (file[pos] == ' ' || file[pos] == '\t' || file[pos] == '\n' || file[pos] == '\r') {
                pos++;
            } else {
                break;
            }
        }
        while (pos < file.size() && !(file[pos] == ' ' || file[pos] == '\t' || file[pos] == '\n' || file[pos] == '\r')) tok->push_back((char)file[pos++]);
        return !tok->empty();
    };
    std::string t;
    if (!token(&t) || t != "P5") {
        *err = path + ": not a binary PGM (P5)";
        return false;
    }
    int vals[3];
    for (int i = 0; i < 3; i++) {
        if (!token(&t)) {
            *err = path + ": truncated PGM header";
            return false;
        }
        vals[i] = atoi(t.c_str());
    }
    pos++;  // the single white-space byte after maxval
    const int W = vals[0], H = vals[1];
    if (W <= 0 || H <= 0 || vals[2] <= 0 || vals[2] > 255 || pos + (size_t)W * H > file.size()) {
        *err = path + ": unsupported or truncated PGM";
        return false;
    }
    out->width = W;
    out->height = H;
    out->channels = 1;
    out->data.assign(file.begin() + pos, file.begin() + pos + (size_t)W * H);
    return true;
}

bool write_pgm(const std::string &path, const uint8_t *data, int width, int height, std::string *err) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) {
        *err = "cannot write " + path;
        return false;
    }
    fprintf(f, "P5\n%d %d\n255\n", width, height);
    const size_t n = (size_t)width * height;
    const bool ok = fwrite(data, 1, n, f) == n;
    fclose(f);
    if (!ok) *err = "short write on " + path;
    return ok;
}

}  // namespace svb

// C-ABI wrapper (include/elas_b200.h): PNG or PGM by signature
#include "../../include/elas_b200.h"
namespace svb {
void set_error(const char *fmt, ...);
}
extern "C" int svb_image_read(const char *path, uint8_t *out, int64_t capacity, int *width, int *height, int *channels) {
    if (!path || !width || !height || !channels) return SVB_ERR_ARG;
    svb::ImageU8 im;
    std::string err;
    const size_t n = strlen(path);
    const bool pgm = n > 4 && (!strcmp(path + n - 4, ".pgm") || !strcmp(path + n - 4, ".PGM"));
    bool ok = false;
    try {  // no exception may cross the C boundary (std::bad_alloc on a huge image, ...)
        ok = pgm ? svb::read_pgm(path, &im, &err) : svb::read_png(path, &im, &err);
    } catch (const std::exception &e) {
        err = std::string(path) + ": " + e.what();
        ok = false;
    }
    if (!ok) {
        svb::set_error("%s", err.c_str());
        return SVB_ERR_ARG;
    }
    *width = im.width;
    *height = im.height;
    *channels = im.channels;
    if (out) {
        if ((int64_t)im.data.size() > capacity) {
            svb::set_error("image needs %zu bytes, capacity %lld", im.data.size(), (long long)capacity);
            return SVB_ERR_ARG;
        }
        memcpy(out, im.data.data(), im.data.size());
    }
    return SVB_OK;
}
