// TEST INFRASTRUCTURE (CPU oracle; never linked or called by the product).
// The image-loading step in front of the path: Examples/RGB-L/rgbl_kitti.cc:87  `cv::imread(file, cv::IMREAD_UNCHANGED)` followed by
// Tracking::GrabImageRGBL's conversion to gray (src/Tracking.cc:1567-1580: cvtColor RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY by
// mbRGB and the channel count).  The PNG reader lives in a third-party dependency that is absent from /root/reference: OpenCV's imgcodecs
// (find_package(OpenCV 4.4), unpinned) over libpng + zlib.  Restated here from the published format (PNG specification, ISO/IEC 15948 /
// RFC 2083: chunk layout 5.3, IHDR 11.2.2, IDAT = one zlib stream over all IDAT chunks 10, scanline filters None/Sub/Up/Average/Paeth
// 9.2-9.4) and from cv::cvtColor's 8-bit gray formula (15-bit fixed point: R 9798, G 19235, B 3735, +2^14, >> 15).
// Pinned by tests/test_oracle_png.py against python-cv2 4.13 (cv2.imdecode(IMREAD_UNCHANGED) + cv2.cvtColor) live when cv2 is
// importable, and by committed fixtures (tests/golden/png_*.npz) otherwise.
// Scope (what imread returns as CV_8U without a conversion of its own): bit depth 8, colour types 0 (gray), 2 (RGB), 6 (RGBA),
// no interlace.  imread hands colour data over in B, G, R(, A) order; the flag `camera_rgb` is Camera.RGB of the settings file (mbRGB).
#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

extern "C" {

// -> 0, or <0: -1 not a PNG / truncated, -2 unsupported format, -3 corrupt stream, -4 output too small
// channels_out: 1, 3, 4 (of the file).  gray: h x w bytes (stride w).  pixels (nullable): the decoded samples in FILE order (R,G,B[,A]).
int orc_png_decode_gray(const uint8_t* png, size_t n, int camera_rgb, int* w_out, int* h_out, int* channels_out, uint8_t* gray,
                        size_t gray_cap, uint8_t* pixels, size_t pixels_cap) {
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (n < 8 + 25 || std::memcmp(png, sig, 8)) return -1;
    size_t pos = 8;
    int w = 0, h = 0, ch = 0;
    std::vector<uint8_t> idat;
    bool have_ihdr = false, end = false;
    while (!end) {
        if (pos + 12 > n) return -1;
        const uint32_t len = be32(png + pos);
        const uint8_t* type = png + pos + 4;
        if (pos + 12 + (size_t)len > n) return -1;
        const uint8_t* data = png + pos + 8;
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4) != be32(data + len)) return -3;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return -3;
            w = (int)be32(data); h = (int)be32(data + 4);
            const int depth = data[8], ctype = data[9], interlace = data[12];
            if (depth != 8 || interlace != 0 || data[10] != 0 || data[11] != 0) return -2;
            ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 6 ? 4 : 0;
            if (!ch || w <= 0 || h <= 0) return -2;
            have_ihdr = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            if (!have_ihdr) return -3;
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr) return -1;
    const size_t row = (size_t)w * ch, raw_bytes = (row + 1) * h;
    std::vector<uint8_t> raw(raw_bytes);
    uLongf got = (uLongf)raw_bytes;
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw_bytes) return -3;
    std::vector<uint8_t> img(row * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t ft = raw[(row + 1) * y];
        const uint8_t* f = &raw[(row + 1) * y + 1];
        uint8_t* cur = &img[row * y];
        const uint8_t* up = y ? &img[row * (y - 1)] : nullptr;
        if (ft > 4) return -3;
        for (size_t i = 0; i < row; ++i) {
            const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
            int pred = 0;
            switch (ft) { case 1: pred = a; break; case 2: pred = b; break; case 3: pred = (a + b) >> 1; break; case 4: pred = paeth(a, b, c); break; default: break; }
            cur[i] = (uint8_t)(f[i] + pred);
        }
    }
    *w_out = w; *h_out = h; *channels_out = ch;
    if (pixels) { if (pixels_cap < img.size()) return -4; std::memcpy(pixels, img.data(), img.size()); }
    if (gray) {
        if (gray_cap < (size_t)w * h) return -4;
        for (size_t p = 0; p < (size_t)w * h; ++p) {
            if (ch == 1) { gray[p] = img[p]; continue; }
            const int R = img[p * ch], G = img[p * ch + 1], B = img[p * ch + 2];
            // imread: Mat channels (B, G, R); cvtColor(RGB2GRAY) reads channel 0 as red, cvtColor(BGR2GRAY) as blue
            const int c0 = B, c1 = G, c2 = R;
            const int v = camera_rgb ? (c0 * 9798 + c1 * 19235 + c2 * 3735) : (c0 * 3735 + c1 * 19235 + c2 * 9798);
            gray[p] = (uint8_t)((v + (1 << 14)) >> 15);
        }
    }
    return 0;
}

}  // extern "C"
