// texir_matstage.cu -- the per-step texture work of the Mat stage in a handful of launches (the reference
// issues four dr.texture calls per step, models/mat_nvdiffrast.py:131-139, each of which rebuilds its pyramid, and its
// backward walks the pyramids level by level).  Addressing rules: texir_tex.h.  Everything here computes exactly what
// the level-by-level kernels of texir_texture.cu compute -- same operands, same order per texel -- in fewer passes:
//
//   k_mip_build_tile   levels 1..5 of a pyramid from one read of level 0 (a CTA owns a 32x32 tile, shared memory)
//   k_mip_build_tail   levels 6..L by one CTA (at most 128x128 texels are left)
//   k_mip_fold_tile    transpose of the build as a GATHER: every level-0 texel evaluates its own chain
//                      g0 + w (g1 + w (g2 + ...)) coarsest first, the part above level 5 once per tile; optionally
//                      writes to a separate output and leaves the scatter buffers zeroed for the next step
//   k_mat_fetch        albedo (mip), roughness (mip), roughness (bilinear) and irradiance (mip) of a pixel in one pass
#include <cuda_runtime.h>

#include <string>

#include "../../include/texir_b200.h"
#include "texir_host.h"
#include "texir_tex.h"

using namespace texir;
using texir_host::fail;
using texir_host::finish_launch;

namespace {

constexpr int kTile = 32;       // level-0 texels per tile side
constexpr int kTileLevels = 5;  // levels produced / consumed inside a tile (32 -> 1)

// 2x2 box average in the operand order of k_mip_build_level: ((a + b) + c) + d, times 0.25
__device__ __forceinline__ float box4(float a, float b, float c, float d) {
    float acc = 0.f;
    acc += a;
    acc += b;
    acc += c;
    acc += d;
    return acc * 0.25f;
}

template <int C>
__global__ void __launch_bounds__(256) k_mip_build_tile(const float* __restrict__ tex, int H, int W, int nl,
                                                        float* __restrict__ mips) {
    __shared__ float bufA[kTile * kTile * C];
    __shared__ float bufB[(kTile / 2) * (kTile / 2) * C];
    const int tx = blockIdx.x, ty = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < kTile * kTile * C; i += 256) {
        const int row = i / (kTile * C), within = i - row * (kTile * C);
        bufA[i] = __ldg(tex + ((int64_t)(ty * kTile + row) * W + (int64_t)tx * kTile) * C + within);
    }
    __syncthreads();
    float* src = bufA;
    float* dst = bufB;
    int side = kTile;
    for (int l = 1; l <= nl; ++l) {
        const int half = side >> 1;
        const int wl = W >> l;
        float* out = mips + mip_offset(H, W, C, l);
        for (int i = tid; i < half * half * C; i += 256) {
            const int c = i % C, x = (i / C) % half, y = i / (C * half);
            const float v = box4(src[((2 * y) * side + 2 * x) * C + c], src[((2 * y) * side + 2 * x + 1) * C + c],
                                 src[((2 * y + 1) * side + 2 * x) * C + c], src[((2 * y + 1) * side + 2 * x + 1) * C + c]);
            dst[i] = v;
            out[((int64_t)(ty * half + y) * wl + (tx * half + x)) * C + c] = v;
        }
        __syncthreads();
        float* t = src;
        src = dst;
        dst = t;
        side = half;
    }
}

// levels first..L, each from the level below it, by ONE CTA (plain loads: the source was written by this kernel)
__global__ void __launch_bounds__(1024) k_mip_build_tail(float* __restrict__ mips, const float* __restrict__ tex, int H,
                                                         int W, int C, int first, int L) {
    for (int l = first; l <= L; ++l) {
        const float* src = l > 1 ? mips + mip_offset(H, W, C, l - 1) : tex;
        float* dst = mips + mip_offset(H, W, C, l);
        const int Hs = mip_dim(H, l - 1), Ws = mip_dim(W, l - 1), Hd = mip_dim(H, l), Wd = mip_dim(W, l);
        const int sx = Ws > 1 ? 2 : 1, sy = Hs > 1 ? 2 : 1;
        for (int i = threadIdx.x; i < Hd * Wd * C; i += blockDim.x) {
            const int c = i % C, x = (i / C) % Wd, y = i / (C * Wd);
            float acc = 0.f;
            for (int dy = 0; dy < sy; ++dy)
                for (int dx = 0; dx < sx; ++dx) acc += src[((int64_t)(y * sy + dy) * Ws + (x * sx + dx)) * C + c];
            dst[i] = acc * (1.0f / (float)(sx * sy));
        }
        __threadfence_block();
        __syncthreads();
    }
}

// The fold as a gather.  The level-by-level fold computes, coarsest level first,  g_{l-1}[i] += g_l[parent(i)] * w_l ;
// for one level-0 texel that is the chain  t_L = g_L,  t_{l-1} = g_{l-1} + t_l * w_l,  result t_0 -- evaluated here
// per texel in that very order (bit-identical), the chain above level 5 once per 32x32 tile.
template <int C>
__global__ void __launch_bounds__(256) k_mip_fold_tile(float* __restrict__ g_tex, float* __restrict__ g_mips, int H, int W,
                                                       int L, float* __restrict__ out, int clear) {
    __shared__ float top[C];
    __shared__ float bufA[kTile * kTile / 4 * C];
    __shared__ float bufB[kTile * kTile / 16 * C];
    const int tx = blockIdx.x, ty = blockIdx.y, tid = threadIdx.x;
    const int nl = L < kTileLevels ? L : kTileLevels;  // levels handled inside the tile
    if (tid < C) {
        float t = 0.f;
        if (L > kTileLevels) {
            // ancestors of this tile's level-5 texel (ty, tx), walking up with the build's halving rule
            int ys[32], xs[32];
            ys[kTileLevels] = ty;
            xs[kTileLevels] = tx;
            for (int l = kTileLevels + 1; l <= L; ++l) {
                ys[l] = mip_dim(H, l - 1) > 1 ? ys[l - 1] >> 1 : ys[l - 1];
                xs[l] = mip_dim(W, l - 1) > 1 ? xs[l - 1] >> 1 : xs[l - 1];
            }
            t = g_mips[mip_offset(H, W, C, L) + ((int64_t)ys[L] * mip_dim(W, L) + xs[L]) * C + tid];
            for (int l = L - 1; l >= kTileLevels; --l) {
                const int sx = mip_dim(W, l) > 1 ? 2 : 1, sy = mip_dim(H, l) > 1 ? 2 : 1;
                t = g_mips[mip_offset(H, W, C, l) + ((int64_t)ys[l] * mip_dim(W, l) + xs[l]) * C + tid] +
                    t * (1.0f / (float)(sx * sy));
            }
        } else if (L == kTileLevels) {
            t = g_mips[mip_offset(H, W, C, L) + ((int64_t)ty * mip_dim(W, L) + tx) * C + tid];
        }
        top[tid] = t;
        if (clear && L >= kTileLevels) g_mips[mip_offset(H, W, C, kTileLevels) + ((int64_t)ty * mip_dim(W, kTileLevels) + tx) * C + tid] = 0.f;
    }
    __syncthreads();
    // levels nl-1 .. 1 of the tile in shared memory, then level 0 straight to `out`
    const float* coarse = top;  // side 1 when nl == 5; for nl < 5 there is nothing above the tile's level nl
    int cside = 1;
    for (int l = nl - 1; l >= 0; --l) {
        const int side = kTile >> l;
        const int wl = mip_dim(W, l);
        float* g = l ? g_mips + mip_offset(H, W, C, l) : g_tex;
        float* dst = l ? ((l & 1) ? bufA : bufB) : nullptr;  // level 1 (16x16) and 3 in bufA, 2 and 4 in bufB
        for (int i = tid; i < side * side * C; i += 256) {
            const int c = i % C, x = (i / C) % side, y = i / (C * side);
            const int64_t gi = ((int64_t)(ty * side + y) * wl + (tx * side + x)) * C + c;
            float v = g[gi];
            if (l < L) {
                const float up = (l + 1 == nl && nl < kTileLevels)
                                     ? g_mips[mip_offset(H, W, C, nl) + ((int64_t)(ty * (side >> 1) + (y >> 1)) * mip_dim(W, nl) + (tx * (side >> 1) + (x >> 1))) * C + c]
                                     : coarse[((y >> 1) * cside + (x >> 1)) * C + c];
                v = v + up * 0.25f;
            }
            if (l) {
                dst[i] = v;
                if (clear) g[gi] = 0.f;
            } else {
                out[gi] = v;
                if (clear && out != g_tex) g[gi] = 0.f;
            }
        }
        __syncthreads();
        coarse = dst;
        cside = side;
    }
}

struct FetchTex {
    const float* tex;
    const float* mips;
    int H, W, L;
};

template <int C>
__device__ __forceinline__ void fetch_trilinear(const FetchTex& t, float u, float v, const float4& da, float* out) {
    const MipLod m = mip_lod(da.x, da.y, da.z, da.w, t.H, t.W, t.L);
    float acc[C];
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int k = 0; k < 2; ++k) {
        const int l = k ? m.level1 : m.level0;
        const float wl = k ? m.frac : 1.0f - m.frac;
        if (k && m.frac == 0.f) break;
        const int h = mip_dim(t.H, l), w = mip_dim(t.W, l);
        const float* src = l ? t.mips + mip_offset(t.H, t.W, C, l) : t.tex;
        const TexTap tp = tex_tap(u, v, h, w);
        for (int c = 0; c < C; ++c)
            acc[c] += wl * (__ldg(src + (int64_t)tp.i00 * C + c) * tp.w00 + __ldg(src + (int64_t)tp.i10 * C + c) * tp.w10 +
                            __ldg(src + (int64_t)tp.i01 * C + c) * tp.w01 + __ldg(src + (int64_t)tp.i11 * C + c) * tp.w11);
    }
    for (int c = 0; c < C; ++c) out[c] = acc[c];
}

__global__ void __launch_bounds__(256) k_mat_fetch(FetchTex ta, FetchTex tr, FetchTex ti, const float* __restrict__ uv,
                                                   const float* __restrict__ uv_da, int64_t P, float* __restrict__ albedo,
                                                   float* __restrict__ rough, float* __restrict__ rough_womip,
                                                   float* __restrict__ irr) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float u = __ldg(uv + 2 * p), v = __ldg(uv + 2 * p + 1);
    const float4 da = __ldg(reinterpret_cast<const float4*>(uv_da) + p);
    float o[3];
    if (albedo) {
        fetch_trilinear<3>(ta, u, v, da, o);
        albedo[3 * p + 0] = o[0]; albedo[3 * p + 1] = o[1]; albedo[3 * p + 2] = o[2];
    }
    if (rough) {
        fetch_trilinear<1>(tr, u, v, da, o);
        rough[p] = o[0];
    }
    if (rough_womip) {  // filter_mode='linear': the four taps of k_texture_fetch
        const TexTap tp = tex_tap(u, v, tr.H, tr.W);
        rough_womip[p] = __ldg(tr.tex + tp.i00) * tp.w00 + __ldg(tr.tex + tp.i10) * tp.w10 + __ldg(tr.tex + tp.i01) * tp.w01 +
                         __ldg(tr.tex + tp.i11) * tp.w11;
    }
    if (irr) {
        fetch_trilinear<3>(ti, u, v, da, o);
        irr[3 * p + 0] = o[0]; irr[3 * p + 1] = o[1]; irr[3 * p + 2] = o[2];
    }
}

bool tileable(int H, int W) { return H >= kTile && W >= kTile && H % kTile == 0 && W % kTile == 0; }

int check_tex(const char* fn, int H, int W, int C, int L) {
    if (H <= 0 || W <= 0 || C <= 0 || C > 4) return fail(std::string(fn) + ": bad texture shape (C <= 4)");
    if (L < 0 || L != mip_level_count(H, W, L))
        return fail(std::string(fn) + ": the texture cannot be halved that many times (odd side, or below 1x1)");
    return 0;
}

}  // namespace

extern "C" {

// the whole pyramid in two launches; same values as texir_texture_mip_build
int texir_texture_mip_build_fast(const float* tex, int H, int W, int C, int levels, float* mips, void* stream) {
    if (int rc = check_tex("texir_texture_mip_build_fast", H, W, C, levels)) return rc;
    if (levels == 0) return 0;
    if (!tex || !mips) return fail("texir_texture_mip_build_fast: NULL argument");
    if (!tileable(H, W) || (C != 1 && C != 3)) return texir_texture_mip_build(tex, H, W, C, levels, mips, stream);
    texir_host::ScopedDevice guard(texir_host::device_of(mips));
    cudaStream_t st = (cudaStream_t)stream;
    const int nl = levels < kTileLevels ? levels : kTileLevels;
    dim3 grid((unsigned)(W / kTile), (unsigned)(H / kTile));
    if (C == 1)
        k_mip_build_tile<1><<<grid, 256, 0, st>>>(tex, H, W, nl, mips);
    else
        k_mip_build_tile<3><<<grid, 256, 0, st>>>(tex, H, W, nl, mips);
    if (int rc = finish_launch("k_mip_build_tile")) return rc;
    if (levels > kTileLevels) {
        k_mip_build_tail<<<1, 1024, 0, st>>>(mips, tex, H, W, C, kTileLevels + 1, levels);
        if (int rc = finish_launch("k_mip_build_tail")) return rc;
    }
    return 0;
}

// g_tex + folded pyramid -> out (out == g_tex: in place).  clear != 0: g_tex (unless it is the output) and g_mips are
// left zeroed, ready to be scattered into again.  Same values as texir_texture_mip_fold.
int texir_texture_mip_fold_into(float* g_tex, float* g_mips, int H, int W, int C, int levels, float* out, int clear,
                                void* stream) {
    if (int rc = check_tex("texir_texture_mip_fold_into", H, W, C, levels)) return rc;
    if (!g_tex || !out || (levels > 0 && !g_mips)) return fail("texir_texture_mip_fold_into: NULL argument");
    texir_host::ScopedDevice guard(texir_host::device_of(g_tex));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes0 = (size_t)H * W * C * sizeof(float);
    const size_t bytes_m = (size_t)mip_offset(H, W, C, levels + 1) * sizeof(float);
    if (!tileable(H, W) || (C != 1 && C != 3) || levels == 0) {  // level by level, then copy / clear
        if (int rc = texir_texture_mip_fold(g_tex, g_mips, H, W, C, levels, stream)) return rc;
        if (out != g_tex) {
            if (cudaMemcpyAsync(out, g_tex, bytes0, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                return fail("texir_texture_mip_fold_into: copy failed", 2);
            if (clear && cudaMemsetAsync(g_tex, 0, bytes0, st) != cudaSuccess) return fail("texir_texture_mip_fold_into: memset failed", 2);
        }
        if (clear && bytes_m && cudaMemsetAsync(g_mips, 0, bytes_m, st) != cudaSuccess)
            return fail("texir_texture_mip_fold_into: memset failed", 2);
        return 0;
    }
    dim3 grid((unsigned)(W / kTile), (unsigned)(H / kTile));
    if (C == 1)
        k_mip_fold_tile<1><<<grid, 256, 0, st>>>(g_tex, g_mips, H, W, levels, out, clear);
    else
        k_mip_fold_tile<3><<<grid, 256, 0, st>>>(g_tex, g_mips, H, W, levels, out, clear);
    if (int rc = finish_launch("k_mip_fold_tile")) return rc;
    if (clear && levels != kTileLevels) {  // a level that several CTAs read is cleared afterwards
        const size_t off = (size_t)mip_offset(H, W, C, levels > kTileLevels ? kTileLevels + 1 : levels) * sizeof(float);
        if (cudaMemsetAsync((char*)g_mips + off, 0, bytes_m - off, st) != cudaSuccess)
            return fail("texir_texture_mip_fold_into: memset failed", 2);
    }
    return 0;
}

int texir_mat_fetch(const float* tex_a, const float* mips_a, int Ha, int Wa, int La, const float* tex_r,
                    const float* mips_r, int Hr, int Wr, int Lr, const float* tex_i, const float* mips_i, int Hi, int Wi,
                    int Li, const float* uv, const float* uv_da, int64_t P, float* albedo, float* rough,
                    float* rough_womip, float* irr, void* stream) {
    const char* fn = "texir_mat_fetch";
    if (P < 0) return fail(std::string(fn) + ": bad size");
    if (P == 0) return 0;
    if (!uv || !uv_da) return fail(std::string(fn) + ": NULL uv / uv_da");
    if (albedo) {
        if (int rc = check_tex(fn, Ha, Wa, 3, La)) return rc;
        if (!tex_a || (La > 0 && !mips_a)) return fail(std::string(fn) + ": NULL albedo texture");
    }
    if (rough || rough_womip) {
        if (int rc = check_tex(fn, Hr, Wr, 1, Lr)) return rc;
        if (!tex_r || (rough && Lr > 0 && !mips_r)) return fail(std::string(fn) + ": NULL roughness texture");
    }
    if (irr) {
        if (int rc = check_tex(fn, Hi, Wi, 3, Li)) return rc;
        if (!tex_i || (Li > 0 && !mips_i)) return fail(std::string(fn) + ": NULL irradiance texture");
    }
    if (!albedo && !rough && !rough_womip && !irr) return fail(std::string(fn) + ": no output");
    texir_host::ScopedDevice guard(texir_host::device_of(uv));
    FetchTex ta{tex_a, mips_a, Ha, Wa, La}, tr{tex_r, mips_r, Hr, Wr, Lr}, ti{tex_i, mips_i, Hi, Wi, Li};
    k_mat_fetch<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ta, tr, ti, uv, uv_da, P, albedo, rough,
                                                                              rough_womip, irr);
    return finish_launch("k_mat_fetch");
}

}  // extern "C"
