// texir_math.h -- per-sample arithmetic of the TexIR hot path, shared by every kernel.
//
// Everything here is `TX_HD` (host + device) so the exact code the sm_100a kernels run can
// also be compiled with g++ by tests/cpu_emul/ and checked against the oracle on a box with
// no GPU.  The product (libtexir_b200.so) only ever runs it on the device.
//
// Reference semantics followed (file:line into the reference repo):
//   utils/sample_util.py:28-41     RadicalInverse / Hammersley
//   utils/sample_util.py:63-146    generate_dir (frame, Cranley-Patterson shift, 3 modes)
//   models/mat_nvdiffrast.py:218   view = normalize(cam - p, eps=1e-4)
//   models/mat_nvdiffrast.py:252-279  diffuse_reflectance / specular_reflectance
//   models/mat_nvdiffrast.py:292-320  query_irf (hit mask, bary clip, UV interp, grid_sample)
//   models/tracer_o3d_irt.py:165-173  IrT accumulation
//
// Rounding discipline.  Whatever feeds a RAY DIRECTION or a TEXEL COORDINATE is written with
// explicitly rounded single operations (tx_mul/tx_add/...: never contracted into FMA) in the
// order the reference's chain of separate torch kernels evaluates them, so that directions
// and texel coordinates are bit-identical to the reference wherever sin/cos agree.  A 1-ulp
// change of direction can move a hit across a triangle edge into another atlas cell, which
// is the one discontinuity on this path.  Smooth factors (BRDF weights) use ordinary
// expressions.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define TX_HD __host__ __device__ __forceinline__
#define TX_D __device__ __forceinline__
#else
#define TX_HD inline
#define TX_D inline
#endif

#ifdef __CUDA_ARCH__
#define tx_mul(a, b) __fmul_rn((a), (b))
#define tx_add(a, b) __fadd_rn((a), (b))
#define tx_sub(a, b) __fsub_rn((a), (b))
#define tx_div(a, b) __fdiv_rn((a), (b))
#define tx_sqrt(a) __fsqrt_rn((a))
#define tx_fma(a, b, c) __fmaf_rn((a), (b), (c))
#define tx_dmul(a, b) __dmul_rn((a), (b))
#define tx_dadd(a, b) __dadd_rn((a), (b))
#define tx_dsub(a, b) __dsub_rn((a), (b))
#else
// host build: the translation unit must be compiled with -ffp-contract=off
static inline float tx_mul(float a, float b) { return a * b; }
static inline float tx_add(float a, float b) { return a + b; }
static inline float tx_sub(float a, float b) { return a - b; }
static inline float tx_div(float a, float b) { return a / b; }
static inline float tx_sqrt(float a) { return sqrtf(a); }
static inline float tx_fma(float a, float b, float c) { return fmaf(a, b, c); }
static inline double tx_dmul(double a, double b) { return a * b; }
static inline double tx_dadd(double a, double b) { return a + b; }
static inline double tx_dsub(double a, double b) { return a - b; }
#endif

namespace texir {

enum SampleMode { MODE_UNIFORM = 0, MODE_COSINE = 1, MODE_IMPORTANCE = 2 };

constexpr float kTiny = 1e-6f;             // TINY_NUMBER, utils/sample_util.py:25
constexpr float kOneMinusTiny = 0.999999f; // float32(1 - 1e-6)
constexpr float kTwoPi = 6.283185307179586f;
constexpr float kPi = 3.141592653589793f;
constexpr float kHitEps = 1e-4f;           // post-hoc t > 1e-4 mask, mat_nvdiffrast.py:299
constexpr uint32_t kInvalidId = 0xFFFFFFFFu;

struct V3 {
    float x, y, z;
};
TX_HD V3 v3(float x, float y, float z) {
    V3 r;
    r.x = x; r.y = y; r.z = z;
    return r;
}

// ---------------------------------------------------------------------------------------------
// explicitly rounded helpers (mirror separate torch elementwise kernels)
// ---------------------------------------------------------------------------------------------
TX_HD V3 rn_scale(V3 a, float s) { return v3(tx_mul(a.x, s), tx_mul(a.y, s), tx_mul(a.z, s)); }
TX_HD V3 rn_add(V3 a, V3 b) { return v3(tx_add(a.x, b.x), tx_add(a.y, b.y), tx_add(a.z, b.z)); }
TX_HD V3 rn_sub(V3 a, V3 b) { return v3(tx_sub(a.x, b.x), tx_sub(a.y, b.y), tx_sub(a.z, b.z)); }
// Reductions over the last dimension (three elements).  The reference executes them with torch ON THE CUDA DEVICE, whose
// reduction kernel sums a CONTIGUOUS row as (e0 + e2) + e1 (two lanes share the row: lane 0 takes elements 0 and 2) and a
// row of a BROADCAST (expanded, stride-0) tensor as (e0 + e1) + e2, every product rounded separately -- measured on
// B200 with torch 2.11: 100 % bit-equal on 4e5 vectors for each of the two cases, every other order <= 95 %
// (profiles/discover_torch_cuda_rounding.py, profiles/r02_torch_cuda_rounding.txt).  The DEVICE code follows that.
// torch on the CPU (where tests/golden were generated) accumulates acc = fma(x_i, x_i, acc) for the norm and sums a
// dot product in order; the HOST build of this header (tests/cpu_emul) keeps those forms so that the emulation stays
// bit-equal to the CPU goldens.
#ifdef __CUDA_ARCH__
// torch.sum(a * b, dim=-1) of a contiguous product tensor
TX_HD float rn_dot(V3 a, V3 b) {
    return tx_add(tx_add(tx_mul(a.x, b.x), tx_mul(a.z, b.z)), tx_mul(a.y, b.y));
}
// torch.norm(x, dim=-1) of a contiguous tensor: norm_axis(torch.cross(..)), F.normalize(cam - p)
TX_HD float rn_norm(V3 a) { return tx_sqrt(tx_add(tx_add(tx_mul(a.x, a.x), tx_mul(a.z, a.z)), tx_mul(a.y, a.y))); }
// torch.norm(x, dim=-1) of an expanded tensor: norm_axis(normals.unsqueeze(1).expand(b, N, 3)), sample_util.py:80,89
TX_HD float rn_norm_bcast(V3 a) { return tx_sqrt(tx_add(tx_add(tx_mul(a.x, a.x), tx_mul(a.y, a.y)), tx_mul(a.z, a.z))); }
#else
TX_HD float rn_dot(V3 a, V3 b) {
    return tx_add(tx_add(tx_mul(a.x, b.x), tx_mul(a.y, b.y)), tx_mul(a.z, b.z));
}
TX_HD float rn_norm(V3 a) { return tx_sqrt(tx_fma(a.z, a.z, tx_fma(a.y, a.y, tx_mul(a.x, a.x)))); }
TX_HD float rn_norm_bcast(V3 a) { return rn_norm(a); }
#endif
// torch.cross on CPU: c.x = fma(a.y, b.z, -(a.z*b.y))  (100 % bit-equal; the unfused form is 75 %)
TX_HD V3 rn_cross(V3 a, V3 b) {
    return v3(tx_fma(a.y, b.z, -tx_mul(a.z, b.y)),
              tx_fma(a.z, b.x, -tx_mul(a.x, b.z)),
              tx_fma(a.x, b.y, -tx_mul(a.y, b.x)));
}
// norm_axis(x) = x / (norm(x) + 1e-6), utils/sample_util.py:86-87
TX_HD V3 rn_norm_axis(V3 a) {
    float d = tx_add(rn_norm(a), kTiny);
    return v3(tx_div(a.x, d), tx_div(a.y, d), tx_div(a.z, d));
}
TX_HD V3 rn_norm_axis_bcast(V3 a) {
    float d = tx_add(rn_norm_bcast(a), kTiny);
    return v3(tx_div(a.x, d), tx_div(a.y, d), tx_div(a.z, d));
}

// ordinary (contractable) helpers for smooth factors
TX_HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
TX_HD float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// ---------------------------------------------------------------------------------------------
// Hammersley point set (utils/sample_util.py:28-41)
// ---------------------------------------------------------------------------------------------
TX_HD uint32_t reverse_bits32(uint32_t b) {
#ifdef __CUDA_ARCH__
    return __brev(b);
#else
    b = (b << 16) | (b >> 16);
    b = ((b & 0x55555555u) << 1) | ((b & 0xAAAAAAAAu) >> 1);
    b = ((b & 0x33333333u) << 2) | ((b & 0xCCCCCCCCu) >> 2);
    b = ((b & 0x0F0F0F0Fu) << 4) | ((b & 0xF0F0F0F0u) >> 4);
    b = ((b & 0x00FF00FFu) << 8) | ((b & 0xFF00FF00u) >> 8);
    return b;
#endif
}
// float32(float64(bits) * 2^-32): the product is exact in float64, so one rounding to float32.
TX_HD float radical_inverse(uint32_t i) {
    return (float)((double)reverse_bits32(i) * 2.3283064365386963e-10);
}
// float32(float64(i)/float64(N)) == correctly rounded float32 quotient for i, N < 2^24
TX_HD float hammersley0(uint32_t i, uint32_t n) { return tx_div((float)i, (float)n); }

// Cranley-Patterson rotation + wrap + clamp (utils/sample_util.py:102-108)
TX_HD float shift_wrap_clamp(float s, float shift) {
    float v = tx_add(s, shift);
    if (v > 1.0f) v = tx_sub(v, 1.0f);
    if (v < 0.0f) v = tx_add(v, 1.0f);
    return fminf(fmaxf(v, kTiny), kOneMinusTiny);
}

// ---------------------------------------------------------------------------------------------
// tangent frame (utils/sample_util.py:84-91).  x-axis is chosen from the RAW normal.
// ---------------------------------------------------------------------------------------------
struct Frame {
    V3 n;  // normalised normal  N^
    V3 u;  // U
    V3 v;  // V
};
TX_HD Frame make_frame(V3 nraw) {
    Frame f;
    V3 xa = (fabsf(nraw.x) > 0.99f) ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
    f.n = rn_norm_axis_bcast(nraw);
    f.u = rn_norm_axis(rn_cross(xa, f.n));
    f.v = rn_norm_axis(rn_cross(f.n, f.u));
    return f;
}

// ---------------------------------------------------------------------------------------------
// sample -> hemisphere direction (utils/sample_util.py:115-143)
// ---------------------------------------------------------------------------------------------
struct DirSample {
    V3 d;        // direction
    float c, st; // cos(theta), sin(theta) after clamps
    float sphi, cphi;
    // importance-mode tangents d/d(roughness); zero for the other modes
    float dc, dst;
};

TX_HD void tx_sincos(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
    sincosf(x, s, c);
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}

// L = V*(sin(phi)*sinT) + N*cosT + U*(-(cos(phi)*sinT))
TX_HD V3 compose_dir(const Frame& f, float sphi, float cphi, float c, float st) {
    float a = tx_mul(sphi, st);
    float b = -tx_mul(cphi, st);
    V3 t = rn_add(rn_scale(f.v, a), rn_scale(f.n, c));
    return rn_add(t, rn_scale(f.u, b));
}

template <bool TANGENT>
TX_HD DirSample sample_dir(const Frame& f, int mode, float s0, float s1, float rough) {
    DirSample o;
    float phi = tx_sub(tx_mul(s1, kTwoPi), kPi);
    tx_sincos(phi, &o.sphi, &o.cphi);
    o.dc = 0.f;
    o.dst = 0.f;
    if (mode == MODE_UNIFORM) {
        o.c = tx_sub(1.0f, s0);
        o.st = tx_sqrt(tx_sub(1.0f, tx_mul(o.c, o.c)));
    } else if (mode == MODE_COSINE) {
        o.c = tx_sqrt(tx_sub(1.0f, s0));
        o.st = tx_sqrt(tx_sub(1.0f, tx_mul(o.c, o.c)));
    } else {
        const float lo = -0.999999f, hi = 0.999999f;
        float a = tx_mul(rough, rough);
        float a2m1 = tx_sub(tx_mul(a, a), 1.0f);
        float D = tx_add(1.0f, tx_mul(a2m1, s0));
        float q = tx_div(tx_sub(1.0f, s0), D);
        float craw = tx_sqrt(q);
        o.c = fminf(fmaxf(craw, lo), hi);
        float straw = tx_sqrt(tx_sub(1.0f, tx_mul(o.c, o.c)));
        o.st = fminf(fmaxf(straw, lo), hi);
        if (TANGENT) {
            // d/dr:  a = r^2, (a^2)' = 4 r^3, D' = s0 * 4 r^3, craw' = -craw * D' / (2 D)
            float dD = s0 * 4.0f * rough * rough * rough;
            float dcraw = -craw * dD / (2.0f * D);
            o.dc = (craw >= lo && craw <= hi) ? dcraw : 0.f;
            float dstraw = -o.c * o.dc / straw;
            o.dst = (straw >= lo && straw <= hi) ? dstraw : 0.f;
        }
    }
    o.d = compose_dir(f, o.sphi, o.cphi, o.c, o.st);
    return o;
}

// view = (cam - p) / max(||cam - p||, 1e-4)   (F.normalize, mat_nvdiffrast.py:218)
TX_HD V3 view_dir(V3 cam, V3 p) {
    V3 w = rn_sub(cam, p);
    float n = fmaxf(rn_norm(w), 1e-4f);
    return v3(tx_div(w.x, n), tx_div(w.y, n), tx_div(w.z, n));
}

// ---------------------------------------------------------------------------------------------
// closest-hit contract (Open3D RaycastingScene.cast_rays -> Embree rtcIntersect1, default flags)
// Moeller-Trumbore in Embree's formulation: e1 = v0 - v1, e2 = v2 - v0, Ng = e2 x e1,
// C = v0 - O, R = C x D, den = Ng.D, U = R.e2, V = R.e1, T = Ng.C (all times sign(den)).  tnear = 0 (exclusive), tfar = inf, no culling.  (u, v) weight v1, v2.
// Every product-sum is an explicit fmaf chain so gcc and nvcc round identically.
// ---------------------------------------------------------------------------------------------
TX_HD float fdot(V3 a, V3 b) { return tx_fma(a.z, b.z, tx_fma(a.y, b.y, tx_mul(a.x, b.x))); }
TX_HD V3 fcross(V3 a, V3 b) {
    return v3(tx_fma(a.y, b.z, -tx_mul(a.z, b.y)),
              tx_fma(a.z, b.x, -tx_mul(a.x, b.z)),
              tx_fma(a.x, b.y, -tx_mul(a.y, b.x)));
}

struct Hit {
    float t, u, v;
    uint32_t prim;  // original primitive id, kInvalidId on miss
};

// returns true and updates (t,u,v) when the triangle is hit at 0 < t and (t < best_t, or
// t == best_t with a lower primitive id)
TX_HD bool tri_test(V3 v0, V3 e1, V3 e2, V3 ng, V3 o, V3 d, float* t, float* u, float* v) {
    V3 c = rn_sub(v0, o);
    V3 r = fcross(c, d);
    float den = fdot(ng, d);
    float aden = fabsf(den);
    float sgn = (den < 0.f) ? -1.0f : 1.0f;
    float U = tx_mul(fdot(r, e2), sgn);
    float V = tx_mul(fdot(r, e1), sgn);
    float T = tx_mul(fdot(ng, c), sgn);
    // NaN-safe: every comparison must be true
    if (!(den != 0.f && U >= 0.f && V >= 0.f && tx_add(U, V) <= aden && T > 0.f)) return false;
    float rcp = tx_div(1.0f, aden);
    *t = tx_mul(T, rcp);
    *u = tx_mul(U, rcp);
    *v = tx_mul(V, rcp);
    return true;
}

// ---------------------------------------------------------------------------------------------
// radiance lookup (mat_nvdiffrast.py:303-318 + ATen grid_sampler_2d, bilinear / border /
// align_corners=False).  UV interpolation in float64 from float64 per-corner UVs like the
// reference's numpy code, then float32 from `grid = torch.from_numpy(grid).float()` on.
// ---------------------------------------------------------------------------------------------
TX_HD void interp_uv(const double* __restrict__ tuv /* 6 doubles: uv0 uv1 uv2 */, float bu, float bv,
                     float* gu, float* gv) {
    // np.clip(prim_uvs, 0, 1) and (1 - u - v) stay float32 (numpy keeps float32 for float32-array
    // with Python-scalar arithmetic); the products with the float64 triangle_uvs are float64
    float b1f = fminf(fmaxf(bu, 0.f), 1.f), b2f = fminf(fmaxf(bv, 0.f), 1.f);
    double b0 = (double)tx_sub(tx_sub(1.0f, b1f), b2f);
    double b1 = (double)b1f, b2 = (double)b2f;
    double u = tx_dadd(tx_dadd(tx_dmul(tuv[0], b0), tx_dmul(tuv[2], b1)), tx_dmul(tuv[4], b2));
    double v = tx_dadd(tx_dadd(tx_dmul(tuv[1], b0), tx_dmul(tuv[3], b1)), tx_dmul(tuv[5], b2));
    *gu = (float)u;
    *gv = (float)v;
}

struct Bilerp {
    int x0, x1, y0, y1;
    float wnw, wne, wsw, wse;
};

// grid[...,0] = u*2-1 ; grid[...,1] = -(1 - v*2) ; then ATen's vectorised CPU grid_sampler_2d
// (aten/src/ATen/native/cpu/GridSamplerKernel.cpp): unnormalize = fma(g + 1, size/2, -0.5) (the
// compiler contracts `(in + 1) * scaling_factor - 0.5`), border clip to [0, size-1], floor,
// weights, and the four taps accumulated as one fma chain.  Measured against torch 2.11 CPU
// F.grid_sample: 100 % bit-equal on 2e5 random lookups at sizes 64, 1000 and 4096.
TX_HD Bilerp bilerp_setup(float gu, float gv, int W, int H) {
    float gx = tx_sub(tx_mul(gu, 2.0f), 1.0f);
    float gy = -tx_sub(1.0f, tx_mul(gv, 2.0f));
    float x = tx_fma(tx_add(gx, 1.0f), tx_div((float)W, 2.0f), -0.5f);
    float y = tx_fma(tx_add(gy, 1.0f), tx_div((float)H, 2.0f), -0.5f);
    x = fminf((float)(W - 1), fmaxf(x, 0.f));
    y = fminf((float)(H - 1), fmaxf(y, 0.f));
    float xw = floorf(x), yn = floorf(y);
    float w = tx_sub(x, xw), e = tx_sub(1.0f, w);
    float n = tx_sub(y, yn), s = tx_sub(1.0f, n);
    Bilerp b;
    b.x0 = (int)xw;
    b.y0 = (int)yn;
    b.x1 = b.x0 + 1 < W ? b.x0 + 1 : W - 1;  // out-of-range neighbour always has weight 0
    b.y1 = b.y0 + 1 < H ? b.y0 + 1 : H - 1;
    b.wnw = tx_mul(e, s);
    b.wne = tx_mul(w, s);
    b.wsw = tx_mul(e, n);
    b.wse = tx_mul(w, n);
    return b;
}
TX_HD float bilerp_apply(const Bilerp& b, float nw, float ne, float sw, float se) {
    return tx_fma(se, b.wse, tx_fma(sw, b.wsw, tx_fma(ne, b.wne, tx_mul(nw, b.wnw))));
}

// ---------------------------------------------------------------------------------------------
// specular GGX/Schlick-Smith sample weight (mat_nvdiffrast.py:260-279) and its d/d(roughness)
// (chain through generate_dir 'importance', the clamp sub-gradients follow torch: a clamp
// passes the gradient on its closed interval, clamp(min=eps) passes where x >= eps).
// ---------------------------------------------------------------------------------------------
struct SpecPoint {
    V3 n;      // raw normal
    V3 v;      // view
    float ndv; // clamp(n.v, 0, 1)
    float k;   // (r+1)^2/8
    float dk;  // (r+1)/4
    float g1v, dg1v;
    float eps;
};
TX_HD SpecPoint spec_point(V3 nraw, V3 view, float rough, float eps) {
    SpecPoint s;
    s.n = nraw;
    s.v = view;
    s.eps = eps;
    s.ndv = clamp01(rn_dot(nraw, view));
    float r1 = rough + 1.0f;
    s.k = r1 * r1 / 8.0f;
    s.dk = r1 / 4.0f;
    float den = s.ndv * (1.0f - s.k) + s.k;
    float denc = fmaxf(den, eps);
    s.g1v = s.ndv / denc;
    float dden = (den >= eps) ? (1.0f - s.ndv) * s.dk : 0.f;
    s.dg1v = -s.ndv / (denc * denc) * dden;
    return s;
}

// light direction of a specular sample: vdh = clamp(h.v), l = 2*vdh*h - v  (mat_nvdiffrast.py:235-236)
TX_HD V3 reflect_dir(V3 h, V3 view, float* vdh_raw) {
    *vdh_raw = rn_dot(h, view);
    float vdh = clamp01(*vdh_raw);
    float s = tx_mul(2.0f, vdh);
    return rn_sub(rn_scale(h, s), view);
}

// weight w such that specular += L * w (before the 1/N), and dw = dw/d(roughness) if TANGENT
template <bool TANGENT>
TX_HD float spec_weight(const SpecPoint& sp, const Frame& f, const DirSample& ds, V3 l, float vdh_raw,
                        float* dw_out) {
    const float eps = sp.eps;
    V3 h = ds.d;
    float vdh = clamp01(vdh_raw);
    float ndl_raw = rn_dot(sp.n, l);
    float ndh_raw = rn_dot(sp.n, h);
    float ndl = clamp01(ndl_raw), ndh = clamp01(ndh_raw);
    float e = (-5.55472f * vdh - 6.98316f) * vdh;
    float p2 = exp2f(e);
    float F = 0.04f + 0.96f * p2;
    float k = sp.k;
    float denl = ndl * (1.0f - k) + k;
    float denlc = fmaxf(denl, eps);
    float g1l = ndl / denlc;
    float g = g1l * sp.g1v;
    float q4 = 4.0f * ndl * sp.ndv;
    float q4c = fmaxf(q4, eps);
    float brdf = F * g / q4c;
    float ndhc = fmaxf(ndh, eps);
    float w = brdf * ndl * 4.0f * vdh / ndhc;
    if (TANGENT) {
        // h' = (V sin(phi) - U cos(phi)) * st' + N^ * c'
        V3 A = v3(f.v.x * ds.sphi - f.u.x * ds.cphi, f.v.y * ds.sphi - f.u.y * ds.cphi,
                  f.v.z * ds.sphi - f.u.z * ds.cphi);
        V3 dh = v3(A.x * ds.dst + f.n.x * ds.dc, A.y * ds.dst + f.n.y * ds.dc, A.z * ds.dst + f.n.z * ds.dc);
        float dvdh = (vdh_raw >= 0.f && vdh_raw <= 1.f) ? dot3(dh, sp.v) : 0.f;
        V3 dl = v3(2.0f * (dvdh * h.x + vdh * dh.x), 2.0f * (dvdh * h.y + vdh * dh.y),
                   2.0f * (dvdh * h.z + vdh * dh.z));
        float dndl = (ndl_raw >= 0.f && ndl_raw <= 1.f) ? dot3(sp.n, dl) : 0.f;
        float dndh = (ndh_raw >= 0.f && ndh_raw <= 1.f) ? dot3(sp.n, dh) : 0.f;
        float dF = 0.96f * p2 * 0.6931471805599453f * (-2.0f * 5.55472f * vdh - 6.98316f) * dvdh;
        float ddenl = (denl >= eps) ? (dndl * (1.0f - k) + (1.0f - ndl) * sp.dk) : 0.f;
        float dg1l = dndl / denlc - ndl / (denlc * denlc) * ddenl;
        float dg = dg1l * sp.g1v + g1l * sp.dg1v;
        float dq4 = (q4 >= eps) ? 4.0f * dndl * sp.ndv : 0.f;
        float dbrdf = (dF * g + F * dg) / q4c - F * g / (q4c * q4c) * dq4;
        float dndhc = (ndh >= eps) ? dndh : 0.f;
        *dw_out = 4.0f * (dbrdf * ndl * vdh + brdf * dndl * vdh + brdf * ndl * dvdh) / ndhc - w / ndhc * dndhc;
    }
    return w;
}

}  // namespace texir
