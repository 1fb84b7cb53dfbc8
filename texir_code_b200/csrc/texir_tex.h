// texir_tex.h -- material-texture addressing of the Mat stage, host + device.
//
// Replaces nvdiffrast's dr.texture at the reference call sites models/mat_nvdiffrast.py:131-139:
//     albedo    = dr.texture(materials_a[None], texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=13)
//     roughness = dr.texture(materials_r[None], texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=13)
//     roughness_womipmap = dr.texture(materials_r[None], texc, texd, filter_mode='linear')
//     irr       = dr.texture(irrt[None], texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=13)
// nvdiffrast is a third-party CUDA extension that is neither vendored in the reference nor installable in the build
// container, so the rules below restate its published texture op (Laine et al. 2020, "Modular Primitives for
// High-Performance Differentiable Rendering", section 3.4, and the behaviour of its texture.cu): PARITY UNPINNED.
//   * texel centres at (i + 0.5) / size, 'wrap' boundary mode (dr.texture's default);
//   * mip pyramid: level l+1 is the 2x2 box average of level l (2x1 when one side is already 1); levels stop at 1x1
//     or at max_mip_level, and a side that is odd and > 1 cannot be halved;
//   * level of detail from the pixel footprint: with uv_da = (du/dX, du/dY, dv/dX, dv/dY) scaled to texels, the major
//     axis of the footprint ellipse has squared length  l2b + sqrt(l2n),  l2b = (A+B)/2,  l2n = (A-B)^2/4 + C^2,
//     A = |d(s,t)/dX|^2, B = |d(s,t)/dY|^2, C = d(s,t)/dX . d(s,t)/dY;  level = log2(major axis length), clamped to
//     [0, max level]; level0 = floor, level1 = min(level0 + 1, max), blend by the fractional part (trilinear).
#pragma once
#include "texir_math.h"

namespace texir {

struct TexTap {
    int i00, i10, i01, i11;
    float w00, w10, w01, w11;
};
TX_HD int wrap_texel(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
TX_HD TexTap tex_tap(float u, float v, int H, int W) {
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    int x0 = (int)fx, y0 = (int)fy;
    int x1 = wrap_texel(x0 + 1, W), y1 = wrap_texel(y0 + 1, H);
    x0 = wrap_texel(x0, W);
    y0 = wrap_texel(y0, H);
    TexTap t;
    t.i00 = y0 * W + x0;
    t.i10 = y0 * W + x1;
    t.i01 = y1 * W + x0;
    t.i11 = y1 * W + x1;
    t.w00 = (1.f - ax) * (1.f - ay);
    t.w10 = ax * (1.f - ay);
    t.w01 = (1.f - ax) * ay;
    t.w11 = ax * ay;
    return t;
}

TX_HD int mip_dim(int n, int level) {
    const int d = n >> level;
    return d > 0 ? d : 1;
}
// number of mip levels above level 0 a [H,W] texture supports, capped at max_mip_level (< 0: no cap);
// -1 if a side that still has to be halved is odd (nvdiffrast raises for that)
TX_HD int mip_level_count(int H, int W, int max_mip_level) {
    int level = 0, h = H, w = W;
    while ((w | h) > 1 && !(max_mip_level >= 0 && level >= max_mip_level)) {
        if ((w > 1 && (w & 1)) || (h > 1 && (h & 1))) return -1;
        ++level;
        if (w > 1) w >>= 1;
        if (h > 1) h >>= 1;
    }
    return level;
}
// element offset of level l >= 1 inside the packed pyramid buffer (levels 1..L back to back, [h,w,C] each)
TX_HD int64_t mip_offset(int H, int W, int C, int level) {
    int64_t off = 0;
    for (int l = 1; l < level; ++l) off += (int64_t)mip_dim(H, l) * mip_dim(W, l) * C;
    return off;
}

struct MipLod {
    int level0, level1;
    float frac;  // weight of level1; 0 when magnifying or clamped on the last level
};
TX_HD MipLod mip_lod(float dudx, float dudy, float dvdx, float dvdy, int H, int W, int max_level) {
    const float dsdx = dudx * (float)W, dsdy = dudy * (float)W, dtdx = dvdx * (float)H, dtdy = dvdy * (float)H;
    const float A = dsdx * dsdx + dtdx * dtdx, B = dsdy * dsdy + dtdy * dtdy, C = dsdx * dsdy + dtdx * dtdy;
    const float l2b = 0.5f * (A + B);
    const float l2n = 0.25f * (A - B) * (A - B) + C * C;
    const float major2 = l2b + sqrtf(l2n);
#ifdef __CUDA_ARCH__
    float flevel = 0.5f * __log2f(major2);  // may be -inf / NaN: the clamp below fixes both
#else
    float flevel = 0.5f * log2f(major2);
#endif
    flevel = fminf(fmaxf(flevel, 0.f), (float)max_level);
    MipLod m;
    m.level0 = (int)floorf(flevel);
    m.level1 = m.level0;
    m.frac = 0.f;
    if (flevel > 0.f) {
        m.level1 = m.level0 + 1 < max_level ? m.level0 + 1 : max_level;
        m.frac = flevel - (float)m.level0;
    }
    return m;
}

// a texture-space gradient target: level 0 [H,W,C] and the packed pyramid (levels 1..L), both zero-initialised by the
// caller; tex == nullptr means "no texture-space gradient wanted"
struct TexGrad {
    float* tex;
    float* mips;
    int H, W, C, L;
};

#ifdef __CUDACC__
// transpose of the (tri)linear fetch for ONE pixel: g[0..C) scattered with atomics into both levels of detail
__device__ __forceinline__ void tex_scatter_point(const TexGrad& t, float u, float v, const MipLod& m, const float* g) {
    for (int k = 0; k < 2; ++k) {
        const int l = k ? m.level1 : m.level0;
        const float wl = k ? m.frac : 1.0f - m.frac;
        if (k && m.frac == 0.f) break;
        const int h = mip_dim(t.H, l), w = mip_dim(t.W, l);
        float* dst = l ? t.mips + mip_offset(t.H, t.W, t.C, l) : t.tex;
        const TexTap tp = tex_tap(u, v, h, w);
        for (int c = 0; c < t.C; ++c) {
            const float gc = g[c] * wl;
            atomicAdd(dst + (int64_t)tp.i00 * t.C + c, gc * tp.w00);
            atomicAdd(dst + (int64_t)tp.i10 * t.C + c, gc * tp.w10);
            atomicAdd(dst + (int64_t)tp.i01 * t.C + c, gc * tp.w01);
            atomicAdd(dst + (int64_t)tp.i11 * t.C + c, gc * tp.w11);
        }
    }
}
#endif

}  // namespace texir
