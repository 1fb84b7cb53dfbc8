// texir_texture.cu -- mip-mapped material-texture fetch and its texture-space gradient: the
// 'linear-mipmap-linear' dr.texture calls of models/mat_nvdiffrast.py:131-139.  Addressing rules: texir_tex.h.
//
//   k_mip_build_level      level l -> level l+1, 2x2 box average (a streaming pass: reads 4, writes 1)
//   k_texture_fetch_mip    per pixel: level of detail from uv_da, two bilinear taps, blend      (HBM gather)
//   k_texture_scatter_mip  its transpose: red.global.add.f32 into the gradient pyramid         (HBM atomics)
//   k_mip_fold_level       gradient of level l+1 folded into level l (transpose of the box average), coarsest first
#include <cuda_runtime.h>

#include <string>

#include "../../include/texir_b200.h"
#include "texir_host.h"
#include "texir_tex.h"

using namespace texir;
using texir_host::fail;
using texir_host::finish_launch;

__global__ void __launch_bounds__(256) k_mip_build_level(const float* __restrict__ src, int Hs, int Ws, int C,
                                                         float* __restrict__ dst, int Hd, int Wd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Hd * Wd * C) return;
    const int c = (int)(i % C);
    const int x = (int)((i / C) % Wd), y = (int)(i / ((int64_t)C * Wd));
    const int sx = Ws > 1 ? 2 : 1, sy = Hs > 1 ? 2 : 1;
    float acc = 0.f;
    for (int dy = 0; dy < sy; ++dy)
        for (int dx = 0; dx < sx; ++dx) acc += __ldg(src + ((int64_t)(y * sy + dy) * Ws + (x * sx + dx)) * C + c);
    dst[i] = acc * (1.0f / (float)(sx * sy));
}

__global__ void __launch_bounds__(256) k_mip_fold_level(const float* __restrict__ g_coarse, int Hc, int Wc, int C,
                                                        float* __restrict__ g_fine, int Hf, int Wf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Hf * Wf * C) return;
    const int c = (int)(i % C);
    const int x = (int)((i / C) % Wf), y = (int)(i / ((int64_t)C * Wf));
    const int sx = Wf > 1 ? 2 : 1, sy = Hf > 1 ? 2 : 1;
    g_fine[i] += __ldg(g_coarse + ((int64_t)(y / sy) * Wc + (x / sx)) * C + c) * (1.0f / (float)(sx * sy));
}

struct MipPtr {
    const float* base;   // level 0 [H,W,C]
    const float* mips;   // levels 1..L packed
};
__device__ __forceinline__ MipLod lod_of(const float* __restrict__ uv_da, int64_t p, int H, int W, int L) {
    const float4 d = __ldg(reinterpret_cast<const float4*>(uv_da) + p);
    return mip_lod(d.x, d.y, d.z, d.w, H, W, L);
}

__global__ void __launch_bounds__(256) k_texture_fetch_mip(const float* __restrict__ tex,
                                                           const float* __restrict__ mips, int H, int W, int C,
                                                           int L, const float* __restrict__ uv,
                                                           const float* __restrict__ uv_da, int64_t P,
                                                           float* __restrict__ out, float* __restrict__ lod_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float u = __ldg(uv + 2 * p), v = __ldg(uv + 2 * p + 1);
    const MipLod m = lod_of(uv_da, p, H, W, L);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 2; ++k) {
        const int l = k ? m.level1 : m.level0;
        const float wl = k ? m.frac : 1.0f - m.frac;
        if (k && m.frac == 0.f) break;
        const int h = mip_dim(H, l), w = mip_dim(W, l);
        const float* t = l ? mips + mip_offset(H, W, C, l) : tex;
        const TexTap tp = tex_tap(u, v, h, w);
        for (int c = 0; c < C; ++c)
            acc[c] += wl * (__ldg(t + (int64_t)tp.i00 * C + c) * tp.w00 + __ldg(t + (int64_t)tp.i10 * C + c) * tp.w10 +
                            __ldg(t + (int64_t)tp.i01 * C + c) * tp.w01 + __ldg(t + (int64_t)tp.i11 * C + c) * tp.w11);
    }
    for (int c = 0; c < C; ++c) out[p * C + c] = acc[c];
    if (lod_out) lod_out[p] = (float)m.level0 + m.frac;
}

__global__ void __launch_bounds__(256) k_texture_scatter_mip(const float* __restrict__ g_out, int H, int W, int C,
                                                             int L, const float* __restrict__ uv,
                                                             const float* __restrict__ uv_da, int64_t P,
                                                             float* __restrict__ g_tex, float* __restrict__ g_mips) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const TexGrad t{g_tex, g_mips, H, W, C, L};
    float g[4];
    for (int c = 0; c < C; ++c) g[c] = __ldg(g_out + p * C + c);
    tex_scatter_point(t, __ldg(uv + 2 * p), __ldg(uv + 2 * p + 1), lod_of(uv_da, p, H, W, L), g);
}

static inline unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

static int check_tex(const char* fn, int H, int W, int C, int L) {
    if (H <= 0 || W <= 0 || C <= 0 || C > 4) return fail(std::string(fn) + ": bad texture shape (C <= 4)");
    if (L < 0 || L != mip_level_count(H, W, L))
        return fail(std::string(fn) + ": the texture cannot be halved that many times (odd side, or below 1x1)");
    return 0;
}

extern "C" {

int texir_texture_mip_levels(int H, int W, int max_mip_level) {
    return (H > 0 && W > 0) ? mip_level_count(H, W, max_mip_level) : -1;
}

int64_t texir_texture_mip_elems(int H, int W, int C, int levels) {
    return (H > 0 && W > 0 && C > 0 && levels >= 0) ? mip_offset(H, W, C, levels + 1) : -1;
}

int texir_texture_mip_build(const float* tex, int H, int W, int C, int levels, float* mips, void* stream) {
    if (int rc = check_tex("texir_texture_mip_build", H, W, C, levels)) return rc;
    if (levels == 0) return 0;
    if (!tex || !mips) return fail("texir_texture_mip_build: NULL argument");
    texir_host::ScopedDevice guard(texir_host::device_of(mips));
    const float* src = tex;
    for (int l = 1; l <= levels; ++l) {
        float* dst = mips + mip_offset(H, W, C, l);
        const int hs = mip_dim(H, l - 1), ws = mip_dim(W, l - 1), hd = mip_dim(H, l), wd = mip_dim(W, l);
        k_mip_build_level<<<blocks_for((int64_t)hd * wd * C), 256, 0, (cudaStream_t)stream>>>(src, hs, ws, C, dst, hd, wd);
        if (int rc = finish_launch("k_mip_build_level")) return rc;
        src = dst;
    }
    return 0;
}

int texir_texture_mip_fold(float* g_tex, float* g_mips, int H, int W, int C, int levels, void* stream) {
    if (int rc = check_tex("texir_texture_mip_fold", H, W, C, levels)) return rc;
    if (levels == 0) return 0;
    if (!g_tex || !g_mips) return fail("texir_texture_mip_fold: NULL argument");
    texir_host::ScopedDevice guard(texir_host::device_of(g_tex));
    for (int l = levels; l >= 1; --l) {
        const float* coarse = g_mips + mip_offset(H, W, C, l);
        float* fine = l > 1 ? g_mips + mip_offset(H, W, C, l - 1) : g_tex;
        const int hc = mip_dim(H, l), wc = mip_dim(W, l), hf = mip_dim(H, l - 1), wf = mip_dim(W, l - 1);
        k_mip_fold_level<<<blocks_for((int64_t)hf * wf * C), 256, 0, (cudaStream_t)stream>>>(coarse, hc, wc, C, fine, hf, wf);
        if (int rc = finish_launch("k_mip_fold_level")) return rc;
    }
    return 0;
}

int texir_texture_fetch_mip(const float* tex, const float* mips, int H, int W, int C, int levels, const float* uv,
                            const float* uv_da, int64_t P, float* out, float* lod_out, void* stream) {
    if (int rc = check_tex("texir_texture_fetch_mip", H, W, C, levels)) return rc;
    if (P < 0) return fail("texir_texture_fetch_mip: bad size");
    if (P == 0) return 0;
    if (!tex || !uv || !uv_da || !out || (levels > 0 && !mips)) return fail("texir_texture_fetch_mip: NULL argument");
    texir_host::ScopedDevice guard(texir_host::device_of(out));
    k_texture_fetch_mip<<<blocks_for(P), 256, 0, (cudaStream_t)stream>>>(tex, mips, H, W, C, levels, uv, uv_da, P, out,
                                                                         lod_out);
    return finish_launch("k_texture_fetch_mip");
}

int texir_texture_scatter_mip(const float* g_out, int H, int W, int C, int levels, const float* uv,
                              const float* uv_da, int64_t P, float* g_tex, float* g_mips, void* stream) {
    if (int rc = check_tex("texir_texture_scatter_mip", H, W, C, levels)) return rc;
    if (P < 0) return fail("texir_texture_scatter_mip: bad size");
    if (P == 0) return 0;
    if (!g_out || !uv || !uv_da || !g_tex || (levels > 0 && !g_mips))
        return fail("texir_texture_scatter_mip: NULL argument");
    texir_host::ScopedDevice guard(texir_host::device_of(g_tex));
    k_texture_scatter_mip<<<blocks_for(P), 256, 0, (cudaStream_t)stream>>>(g_out, H, W, C, levels, uv, uv_da, P, g_tex,
                                                                           g_mips);
    return finish_launch("k_texture_scatter_mip");
}

}  // extern "C"
