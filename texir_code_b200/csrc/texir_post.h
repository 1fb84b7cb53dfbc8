// texir_post.h -- per-texel arithmetic of the IrT post-processing row, host + device so that
// tests/cpu_emul can run the very same code against scipy / cv2 / the reference on a box without a GPU.
//
// Reference semantics followed (file:line into the reference repo):
//   tools/padding_texture.py:49-80   padding_texture: empty texels (b+g+r == 0) take the value of the nearest
//                                    non-empty texel.  "Nearest" is scipy.ndimage.distance_transform_edt(
//                                    mask, return_indices=True); the value is then read with
//                                    F.grid_sample(mode='nearest') at uv = idx / (w, h) * 2 - 1.
//   trainer/generate_ir_texture.py:80-82, models/mat_nvdiffrast.py:73,89   cv2.imwrite / cv2.imread of
//                                    Radiance .hdr files (RGBE): float2rgbe / rgbe2float of OpenCV's rgbe.cpp.
//
// Nearest-texel rule (what scipy's feature transform computes; checked against scipy 1.18 on random masks with
// many ties, tests/test_oracle.py):
//   phase 1, per column j:  r_j(i) = the non-empty row of column j nearest to row i, ties -> LOWER row
//   phase 2, per row i:     texel (i, x) takes column j* = argmin_j (r_j(i) - i)^2 + (j - x)^2 over the columns that
//                           have a non-empty texel, ties -> LOWER column, and row r_{j*}(i).
// Phase 2 is the lower envelope of parabolas; the envelope is built with scipy's own dominance test (all integer
// arithmetic here; scipy's doubles are exact at these magnitudes), so equal-distance cases resolve identically.
#pragma once
#include "texir_math.h"

namespace texir {

// (b + g) + r == 0 in float32, the channel order of the stored image (numpy adds left to right)
TX_HD bool pad_texel_empty(const float* __restrict__ px, int C) {
    float s = px[0];
    for (int c = 1; c < C; ++c) s = tx_add(s, px[c]);
    return s == 0.0f;
}

// phase 1 combine: nearest of `up` (<= i, or -1) and `down` (>= i, or -1); ties -> the lower row (`up`)
TX_HD int pad_pick_row(int i, int up, int down) {
    if (up < 0) return down;
    if (down < 0) return up;
    return (i - up) <= (down - i) ? up : down;
}

// squared distance of texel (i, x) to the site of column j, whose nearest non-empty row is rj (< 2^31 for sides < 32768)
TX_HD int pad_dist2(int i, int x, int rj, int j) {
    const int dy = rj - i, dx = j - x;
    return dy * dy + dx * dx;
}

// phase 2a: lower envelope of the row's sites, scipy's _VoronoiFT construction loop.  nr[j] = nearest non-empty
// row of column j (kPadNone: the column has none), g = stack of kept columns; returns the index of the last kept site
// (-1: none).  The dominance test is evaluated in float64 exactly as scipy evaluates it (every term is an integer
// below 2^53, so the arithmetic is exact).
constexpr unsigned short kPadNone = 0xFFFFu;
TX_HD int pad_build_envelope(const unsigned short* nr, int W, int i, unsigned short* g) {
    int l = -1;
    int i1 = 0, i2 = 0;        // the two topmost kept columns (g[l], g[l-1]) and their squared row distances,
    double v1 = 0.0, v2 = 0.0;  // cached so that the common no-pop step touches no memory but nr[ii]
    for (int ii = 0; ii < W; ++ii) {
        if (nr[ii] == kPadNone) continue;
        const double dw = (double)((int)nr[ii] - i), wR = dw * dw;
        while (l >= 1) {
            const double a = (double)(i1 - i2), b = (double)(ii - i1), c = a + b;
            if (c * v1 - b * v2 - a * wR - a * b * c <= 0.0) break;  // g[l] still owns part of the row
            --l;
            i1 = i2;
            v1 = v2;
            if (l >= 1) {
                i2 = g[l - 1];
                const double du = (double)((int)nr[i2] - i);
                v2 = du * du;
            }
        }
        g[++l] = (unsigned short)ii;
        i2 = i1;
        v2 = v1;
        i1 = ii;
        v1 = wR;
    }
    return l;
}

// phase 2b: scipy's query loop advances while the NEXT site is strictly nearer; along the envelope that
// predicate holds for a prefix of the sites, so a binary search finds the same site for every x independently.
TX_HD int pad_query_envelope(const unsigned short* nr, const unsigned short* g, int maxl, int i, int x) {
    int lo = 0, hi = maxl;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pad_dist2(i, x, nr[g[mid]], g[mid]) > pad_dist2(i, x, nr[g[mid + 1]], g[mid + 1]))
            lo = mid + 1;
        else
            hi = mid;
    }
    return g[lo];
}

// where F.grid_sample(mode='nearest', align_corners=False) reads for uv = idx / size * 2 - 1
// (tools/padding_texture.py:73-77): torch's CPU kernel un-normalises with (u + 1) * (size / 2) - 0.5 and rounds
// half to even, so an ODD index lands one texel lower at power-of-two sizes -- kept, it is what the reference does.
TX_HD int pad_source_index(int idx, int size) {
    const float s = (float)size;
    const float u = tx_sub(tx_mul(tx_div((float)idx, s), 2.0f), 1.0f);
    const float x = tx_sub(tx_mul(tx_add(u, 1.0f), tx_div(s, 2.0f)), 0.5f);
    const float r = nearbyintf(x);
    return (r >= 0.0f && r <= (float)(size - 1)) ? (int)r : -1;  // -1: outside, padding_mode='zeros'
}

// ---------------------------------------------------------------------------------------------
// Radiance RGBE (OpenCV imgcodecs rgbe.cpp, after Bruce Walter's rgbe.c)
// ---------------------------------------------------------------------------------------------
// float2rgbe: v = max(r,g,b); v < 1e-32 -> 0; else scale = frexp(v, &e) * 256 / v  (= 2^(8-e) exactly),
// mantissas truncate, exponent byte = e + 128.
TX_HD void rgbe_from_float(float red, float green, float blue, unsigned char out[4]) {
    float v = red;
    if (green > v) v = green;
    if (blue > v) v = blue;
    if ((double)v < 1e-32) {  // the reference compares the float with the double constant 1e-32
        out[0] = out[1] = out[2] = out[3] = 0;
        return;
    }
    int e;
    const double m = frexp((double)v, &e);
    const float scale = (float)(m * 256.0 / (double)v);
    out[0] = (unsigned char)(int)tx_mul(red, scale);
    out[1] = (unsigned char)(int)tx_mul(green, scale);
    out[2] = (unsigned char)(int)tx_mul(blue, scale);
    out[3] = (unsigned char)(e + 128);
}
// rgbe2float: f = ldexp(1, e - 136); channel = mantissa * f; exponent byte 0 -> black
TX_HD void rgbe_to_float(const unsigned char in[4], float* red, float* green, float* blue) {
    if (in[3]) {
        const float f = (float)ldexp(1.0, (int)in[3] - (128 + 8));
        *red = tx_mul((float)in[0], f);
        *green = tx_mul((float)in[1], f);
        *blue = tx_mul((float)in[2], f);
    } else {
        *red = *green = *blue = 0.0f;
    }
}

}  // namespace texir
