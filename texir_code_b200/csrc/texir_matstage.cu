// texir_matstage.cu -- the per-step texture work of the Mat stage in a handful of launches (the reference
// issues four dr.texture calls per step, models/mat_nvdiffrast.py:131-139, each of which rebuilds its pyramid, and its
// backward walks the pyramids level by level).  Addressing rules: texir_tex.h.  Everything here computes exactly what
// the level-by-level kernels of texir_texture.cu compute -- same operands, same order per texel -- in fewer passes:
//
//   k_mip_build_tile   levels 1..5 of a pyramid from one read of level 0 (a CTA owns a 32x32 tile, shared memory)
//   k_mip_build_tail   levels 6..L by one CTA (at most 128x128 texels are left)
//   k_mip_fold_gather  transpose of the build as a GATHER: every level-0 texel evaluates its own chain
//                      g0 + w (g1 + w (g2 + ...)) coarsest first, one streaming pass; optionally writes to a separate
//                      output and leaves the scatter buffers zeroed for the next step
//   k_mat_fetch        albedo (mip), roughness (mip), roughness (bilinear) and irradiance (mip) of a pixel in one pass
#include <cuda_runtime.h>

#include <cstdlib>
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

// A CTA of 64 threads owns a 32x32 tile of level 0; thread (tx4, ty4) of the 8x8 thread grid owns a 4x4 block of it.
// Levels 1 and 2 of a block stay in the owner's registers; levels 3 and 4 are 2x2 averages across threads, fetched
// with warp shuffles in the operand order of box4 (lane = ty4 * 8 + tx4, so the partners sit at lane ^ 1 and lane ^ 8,
// then ^ 2 and ^ 16); level 5 joins the two warps through shared memory.  No other synchronisation, every load and
// store of level 0 is a 128-bit access.
template <int C>
__device__ __forceinline__ void load_row4(const float* __restrict__ p, float* row) {  // 4 texels x C channels
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const float4 v = q[k];
        row[4 * k + 0] = v.x; row[4 * k + 1] = v.y; row[4 * k + 2] = v.z; row[4 * k + 3] = v.w;
    }
}
template <int C>
__device__ __forceinline__ void store_row4(float* __restrict__ p, const float* row) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int k = 0; k < C; ++k) q[k] = make_float4(row[4 * k + 0], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]);
}

template <int C>
__global__ void __launch_bounds__(64) k_mip_build_tile(const float* __restrict__ tex, int H, int W, int nl,
                                                       float* __restrict__ mips) {
    __shared__ float s4[4][C];  // the tile's four level-4 values
    const int tx = blockIdx.x, ty = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx4 = lane & 7, ty4 = (lane >> 3) + 4 * warp;
    float t0[4][4 * C];
#pragma unroll
    for (int y = 0; y < 4; ++y)
        load_row4<C>(tex + ((int64_t)(ty * kTile + 4 * ty4 + y) * W + (tx * kTile + 4 * tx4)) * C, t0[y]);
    float t1[2][2][C], t2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i)
                t1[j][i][c] = box4(t0[2 * j][(2 * i) * C + c], t0[2 * j][(2 * i + 1) * C + c], t0[2 * j + 1][(2 * i) * C + c],
                                   t0[2 * j + 1][(2 * i + 1) * C + c]);
        t2[c] = box4(t1[0][0][c], t1[0][1][c], t1[1][0][c], t1[1][1][c]);
    }
    if (nl >= 1) {
        float* out = mips + mip_offset(H, W, C, 1);
        const int w1 = W >> 1;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int c = 0; c < C; ++c)
                    out[((int64_t)(ty * 16 + 2 * ty4 + j) * w1 + (tx * 16 + 2 * tx4 + i)) * C + c] = t1[j][i][c];
    }
    if (nl >= 2) {
        float* out = mips + mip_offset(H, W, C, 2);
#pragma unroll
        for (int c = 0; c < C; ++c) out[((int64_t)(ty * 8 + ty4) * (W >> 2) + (tx * 8 + tx4)) * C + c] = t2[c];
    }
    if (nl < 3) return;  // uniform over the CTA
    float t3[C], t4[C];
    {
        const int b3 = lane & ~9;  // the 2x2 thread group's first lane
        float* out = mips + mip_offset(H, W, C, 3);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            t3[c] = box4(__shfl_sync(0xFFFFFFFFu, t2[c], b3), __shfl_sync(0xFFFFFFFFu, t2[c], b3 + 1),
                         __shfl_sync(0xFFFFFFFFu, t2[c], b3 + 8), __shfl_sync(0xFFFFFFFFu, t2[c], b3 + 9));
            if (lane == b3) out[((int64_t)(ty * 4 + (ty4 >> 1)) * (W >> 3) + (tx * 4 + (tx4 >> 1))) * C + c] = t3[c];
        }
    }
    if (nl < 4) return;
    {
        const int b4 = lane & 4;  // the 4x4 thread group's first lane (this warp holds two such groups side by side)
        float* out = mips + mip_offset(H, W, C, 4);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            t4[c] = box4(__shfl_sync(0xFFFFFFFFu, t3[c], b4), __shfl_sync(0xFFFFFFFFu, t3[c], b4 + 2),
                         __shfl_sync(0xFFFFFFFFu, t3[c], b4 + 16), __shfl_sync(0xFFFFFFFFu, t3[c], b4 + 18));
            if (lane == b4) {
                out[((int64_t)(ty * 2 + warp) * (W >> 4) + (tx * 2 + (tx4 >> 2))) * C + c] = t4[c];
                s4[2 * warp + (tx4 >> 2)][c] = t4[c];
            }
        }
    }
    if (nl < 5) return;
    __syncthreads();
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        mips[mip_offset(H, W, C, 5) + ((int64_t)ty * (W >> 5) + tx) * C + c] = box4(s4[0][c], s4[1][c], s4[2][c], s4[3][c]);
    }
}

// levels first..L, each from the level below it, by ONE CTA.  Level `first` is read from global memory; every later
// level from the copy of its source that the CTA keeps in shared memory (two buffers, ping-pong), so the chain of
// levels costs one global round trip instead of one per level.  smem_floats = elements of level `first`.
__global__ void __launch_bounds__(1024) k_mip_build_tail(float* __restrict__ mips, const float* __restrict__ tex, int H,
                                                         int W, int C, int first, int L, int smem_floats) {
    extern __shared__ float sm[];
    float* buf[2] = {sm, sm + smem_floats};
    for (int l = first; l <= L; ++l) {
        const float* src = l == first ? (l > 1 ? mips + mip_offset(H, W, C, l - 1) : tex) : buf[(l - first + 1) & 1];
        float* keep = buf[(l - first) & 1];
        float* dst = mips + mip_offset(H, W, C, l);
        const int Hs = mip_dim(H, l - 1), Ws = mip_dim(W, l - 1), Hd = mip_dim(H, l), Wd = mip_dim(W, l);
        const int sx = Ws > 1 ? 2 : 1, sy = Hs > 1 ? 2 : 1;
        for (int i = threadIdx.x; i < Hd * Wd * C; i += blockDim.x) {
            const int c = i % C, x = (i / C) % Wd, y = i / (C * Wd);
            float acc = 0.f;
            for (int dy = 0; dy < sy; ++dy)
                for (int dx = 0; dx < sx; ++dx) acc += src[((int64_t)(y * sy + dy) * Ws + (x * sx + dx)) * C + c];
            const float v = acc * (1.0f / (float)(sx * sy));
            dst[i] = v;
            keep[i] = v;
        }
        __syncthreads();
    }
}

// The fold as a gather.  The level-by-level fold computes, coarsest level first,  g_{l-1}[i] += g_l[parent(i)] * w_l ;
// for one level-0 texel that is the chain  t_L = g_L,  t_{l-1} = g_{l-1} + t_l * w_l,  result t_0 -- evaluated here
// per texel in that very order (bit-identical), in ONE streaming pass: a thread owns four consecutive texels of a row
// (128-bit loads and stores), their two level-1 parents and the single chain above; the coarse values are shared
// between neighbouring threads through L1 / L2.  Level l of a [H,W] texture has the texel (y >> min(l, log2 H),
// x >> min(l, log2 W)) above (y, x), and folding into level l weighs by 1 / ((l < log2 W ? 2 : 1) (l < log2 H ? 2 : 1)).
struct FoldLevels {
    int L, ly, lx;        // levels above 0; floor(log2 H), floor(log2 W)
    int64_t off[24];      // element offset of level l >= 1 in the packed pyramid
    int w[24];            // width of level l
};

template <int C>
__global__ void __launch_bounds__(256) k_mip_fold_gather(float* __restrict__ g_tex, const float* __restrict__ g_mips,
                                                         FoldLevels lv, int H, int W, float* __restrict__ out, int clear) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int wq = W >> 2;
    if (idx >= (int64_t)H * wq) return;
    const int y = (int)(idx / wq), x = (int)(idx - (int64_t)y * wq) << 2;
    float t2[C];  // the chain down to level 2 (one texel covers this thread's four)
    const int L = lv.L;
#pragma unroll
    for (int c = 0; c < C; ++c) t2[c] = 0.f;
    if (L >= 2) {
        {
            const int sy = L < lv.ly ? L : lv.ly, sx = L < lv.lx ? L : lv.lx;
            const float* p = g_mips + lv.off[L] + ((int64_t)(y >> sy) * lv.w[L] + (x >> sx)) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) t2[c] = p[c];
        }
        for (int l = L - 1; l >= 2; --l) {
            const int sy = l < lv.ly ? l : lv.ly, sx = l < lv.lx ? l : lv.lx;
            const float wgt = 1.0f / (float)((l < lv.lx ? 2 : 1) * (l < lv.ly ? 2 : 1));
            const float* p = g_mips + lv.off[l] + ((int64_t)(y >> sy) * lv.w[l] + (x >> sx)) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) t2[c] = p[c] + t2[c] * wgt;
        }
    }
    // level 1: two texels above the four (W is a multiple of 4, so level 1 is at least 2 wide and halves in x)
    float t1[2][C];
    if (L >= 1) {
        const int y1 = lv.ly >= 1 ? y >> 1 : y;
        const float w1 = 1.0f / (float)(2 * (1 < lv.ly ? 2 : 1));  // folding level 2 into level 1
        const float* p = g_mips + lv.off[1] + ((int64_t)y1 * lv.w[1] + (x >> 1)) * C;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) t1[i][c] = L == 1 ? p[i * C + c] : p[i * C + c] + t2[c] * w1;
    }
    const float w0 = 1.0f / (float)(2 * (0 < lv.ly ? 2 : 1));  // folding level 1 into level 0
    const int64_t off = ((int64_t)y * W + x) * C;
    float row[4 * C];
    load_row4<C>(g_tex + off, row);
    if (L >= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) row[k * C + c] = row[k * C + c] + t1[k >> 1][c] * w0;
    }
    store_row4<C>(out + off, row);
    if (clear && out != g_tex) {
        const float zero[4 * C] = {};
        store_row4<C>(g_tex + off, zero);
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

// per-pixel gradient rows scattered into the two textures' gradient pyramids: the texture-space half of the Mat-stage
// backward on its own, for rows that were gathered from all ranks (parallel: "gather" gradient synchronisation).
// row = (u, v, du/dX, du/dY, dv/dX, dv/dY, g_albedo[3], g_rough through the mip-mapped fetch, g_rough through the bilinear one)
__global__ void __launch_bounds__(256) k_mat_scatter_rows(const float* __restrict__ rows, int64_t P, TexGrad ga, TexGrad gr) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float* r = rows + 11 * p;
    const float u = __ldg(r + 0), v = __ldg(r + 1);
    const float dudx = __ldg(r + 2), dudy = __ldg(r + 3), dvdx = __ldg(r + 4), dvdy = __ldg(r + 5);
    if (ga.tex) {
        const float t[3] = {__ldg(r + 6), __ldg(r + 7), __ldg(r + 8)};
        if (t[0] != 0.f || t[1] != 0.f || t[2] != 0.f) tex_scatter_point(ga, u, v, mip_lod(dudx, dudy, dvdx, dvdy, ga.H, ga.W, ga.L), t);
    }
    if (gr.tex) {
        const float r_mip = __ldg(r + 9), r_bil = __ldg(r + 10);
        if (r_mip != 0.f) tex_scatter_point(gr, u, v, mip_lod(dudx, dudy, dvdx, dvdy, gr.H, gr.W, gr.L), &r_mip);
        if (r_bil != 0.f) {
            MipLod m0;
            m0.level0 = m0.level1 = 0;
            m0.frac = 0.f;
            tex_scatter_point(gr, u, v, m0, &r_bil);
        }
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
        k_mip_build_tile<1><<<grid, 64, 0, st>>>(tex, H, W, nl, mips);
    else
        k_mip_build_tile<3><<<grid, 64, 0, st>>>(tex, H, W, nl, mips);
    if (int rc = finish_launch("k_mip_build_tile")) return rc;
    if (levels > kTileLevels) {
        const int first = kTileLevels + 1;
        const int n_first = mip_dim(H, first) * mip_dim(W, first) * C;   // <= 16 K floats at 4096^2
        const size_t smem = 2 * (size_t)n_first * sizeof(float);
        if (smem > 96 * 1024) return texir_texture_mip_build(tex, H, W, C, levels, mips, stream);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_mip_build_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_mip_build_tail<<<1, 1024, smem, st>>>(mips, tex, H, W, C, first, levels, n_first);
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
    if (W % 4 != 0 || (C != 1 && C != 3) || levels == 0 || levels >= 24) {  // level by level, then copy / clear
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
    FoldLevels lv{};
    lv.L = levels;
    for (lv.ly = 0; (H >> (lv.ly + 1)) >= 1; ++lv.ly) {}
    for (lv.lx = 0; (W >> (lv.lx + 1)) >= 1; ++lv.lx) {}
    for (int l = 1; l <= levels; ++l) {
        lv.off[l] = mip_offset(H, W, C, l);
        lv.w[l] = mip_dim(W, l);
    }
    const int64_t threads = (int64_t)H * (W / 4);
    const unsigned grid = (unsigned)((threads + 255) / 256);
    // (sixteen texels of a row per thread -- the chain above level 4 shared by 16 texels, 19 loads instead of 56 -- was
    // measured at 129 us against 64 us for the 4096^2 x 1 texture: lanes 64 bytes apart halve the sector use of every
    // 128-bit access)
    if (C == 1)
        k_mip_fold_gather<1><<<grid, 256, 0, st>>>(g_tex, g_mips, lv, H, W, out, clear);
    else
        k_mip_fold_gather<3><<<grid, 256, 0, st>>>(g_tex, g_mips, lv, H, W, out, clear);
    if (int rc = finish_launch("k_mip_fold_gather")) return rc;
    if (clear && bytes_m && cudaMemsetAsync(g_mips, 0, bytes_m, st) != cudaSuccess)
        return fail("texir_texture_mip_fold_into: memset failed", 2);
    return 0;
}

int texir_mat_scatter_rows(const float* rows, int64_t P, float* g_tex_a, float* g_mips_a, int Ha, int Wa, int La, float* g_tex_r,
                           float* g_mips_r, int Hr, int Wr, int Lr, void* stream) {
    const char* fn = "texir_mat_scatter_rows";
    if (P < 0) return fail(std::string(fn) + ": bad size");
    if (P == 0) return 0;
    if (!rows || (!g_tex_a && !g_tex_r)) return fail(std::string(fn) + ": NULL argument");
    if (g_tex_a) {
        if (int rc = check_tex(fn, Ha, Wa, 3, La)) return rc;
        if (La > 0 && !g_mips_a) return fail(std::string(fn) + ": NULL pyramid");
    }
    if (g_tex_r) {
        if (int rc = check_tex(fn, Hr, Wr, 1, Lr)) return rc;
        if (Lr > 0 && !g_mips_r) return fail(std::string(fn) + ": NULL pyramid");
    }
    texir_host::ScopedDevice guard(texir_host::device_of(rows));
    const TexGrad ga{g_tex_a, g_mips_a, Ha, Wa, 3, La}, gr{g_tex_r, g_mips_r, Hr, Wr, 1, Lr};
    k_mat_scatter_rows<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, P, ga, gr);
    return finish_launch("k_mat_scatter_rows");
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
