// texir_point.h -- the per-lane sample loops of the fused kernels (host + device).
//
// A group of G lanes owns one surface point; lane `gl` of the group runs these loops over its share
// of the N samples and the caller reduces the partial sums over the group (warp shuffles on the
// device, a plain loop in tests/cpu_emul).  Sample order inside a group follows coherent_bits().
#pragma once
#include "texir_bvh.h"
#include "texir_stream.h"

namespace texir {

// visits the sample indices of lane gl:  for c in [0, 2^ab): for j = gl, gl+G, ...: i = (j << ab) + c
#define TX_FOR_LANE_SAMPLES(i, N, gl, G, ab)                              \
    for (int c_ = 0, M_ = ((N) + (1 << (ab)) - 1) >> (ab); c_ < (1 << (ab)); ++c_) \
        for (int j_ = (gl); j_ < M_; j_ += (G))                           \
            for (uint32_t i = ((uint32_t)j_ << (ab)) + (uint32_t)c_, once_ = 1; once_ && i < (uint32_t)(N); once_ = 0)

// IrT: partial sum of L_i * clamp(n . l_i, 0, 1)     (models/tracer_o3d_irt.py:168-171)
template <int KIND>
TX_HD void irradiance_lane(const SceneView& s, V3 n, V3 o, float sh0, float sh1, int N, int mode, int gl, int G,
                           int ab, float acc[3]) {
    const Frame f = make_frame(n);
    float ax = 0.f, ay = 0.f, az = 0.f;
    TX_FOR_LANE_SAMPLES(i, N, gl, G, ab) {
        float s0 = shift_wrap_clamp(hammersley0(i, (uint32_t)N), sh0);
        float s1 = shift_wrap_clamp(radical_inverse(i), sh1);
        DirSample ds = sample_dir<false>(f, mode, s0, s1, 0.f);
        Hit h = trace<KIND>(s, o, ds.d);
        V3 L = fetch_radiance(s, h);
        float ndl = clamp01(rn_dot(n, ds.d));
        ax += L.x * ndl;
        ay += L.y * ndl;
        az += L.z * ndl;
    }
    acc[0] = ax;
    acc[1] = ay;
    acc[2] = az;
}
// torch: sum * 2 * np.pi / N
TX_HD float irradiance_finish(float sum, int N) { return tx_div(tx_mul(tx_mul(sum, 2.0f), kPi), (float)N); }

// render(): specular importance samples (+ optional MC diffuse), partial sums per lane
struct ShadeLaneOut {
    float spec[3];   // sum L * w
    float dspec[3];  // sum L * dw/d(roughness)     (TANGENT only)
    float diff[3];   // sum L * (ndl | 1)           (N_diff > 0 only)
};

template <int KIND, bool TANGENT>
TX_HD void shade_lane(const SceneView& s, V3 n, V3 o, V3 cam, float rough, float eps, float shs0, float shs1,
                      int N_spec, int mode_spec, int ab_spec, float shd0, float shd1, int N_diff, int mode_diff,
                      int ab_diff, int gl, int G, ShadeLaneOut& out) {
    const V3 view = view_dir(cam, o);
    const Frame f = make_frame(n);
    const SpecPoint sp = spec_point(n, view, rough, eps);
    float sx = 0.f, sy = 0.f, sz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
    TX_FOR_LANE_SAMPLES(i, N_spec, gl, G, ab_spec) {
        float s0 = shift_wrap_clamp(hammersley0(i, (uint32_t)N_spec), shs0);
        float s1 = shift_wrap_clamp(radical_inverse(i), shs1);
        DirSample ds = sample_dir<TANGENT>(f, mode_spec, s0, s1, rough);
        float vdh_raw;
        V3 l = reflect_dir(ds.d, view, &vdh_raw);
        Hit h = trace<KIND>(s, o, l);
        V3 L = fetch_radiance(s, h);
        float dw = 0.f;
        float w = spec_weight<TANGENT>(sp, f, ds, l, vdh_raw, &dw);
        sx += L.x * w;
        sy += L.y * w;
        sz += L.z * w;
        if (TANGENT) {
            dx += L.x * dw;
            dy += L.y * dw;
            dz += L.z * dw;
        }
    }
    float ex = 0.f, ey = 0.f, ez = 0.f;
    if (N_diff > 0) {
        TX_FOR_LANE_SAMPLES(i, N_diff, gl, G, ab_diff) {
            float s0 = shift_wrap_clamp(hammersley0(i, (uint32_t)N_diff), shd0);
            float s1 = shift_wrap_clamp(radical_inverse(i), shd1);
            DirSample ds = sample_dir<false>(f, mode_diff, s0, s1, 0.f);
            Hit h = trace<KIND>(s, o, ds.d);
            V3 L = fetch_radiance(s, h);
            float wgt = (mode_diff == MODE_COSINE) ? 1.0f : clamp01(rn_dot(n, ds.d));
            ex += L.x * wgt;
            ey += L.y * wgt;
            ez += L.z * wgt;
        }
    }
    out.spec[0] = sx; out.spec[1] = sy; out.spec[2] = sz;
    out.dspec[0] = dx; out.dspec[1] = dy; out.dspec[2] = dz;
    out.diff[0] = ex; out.diff[1] = ey; out.diff[2] = ez;
}

// ---------------------------------------------------------------------------------------------------
// ray-stream work sources (texir_stream.h): the same per-lane sample sequences and the same accumulation
// order as the loops above, so the results are bit-identical to them
// ---------------------------------------------------------------------------------------------------
struct LaneSamples {  // iterator over TX_FOR_LANE_SAMPLES
    int c, j, gl, G, ab, M, N;
    TX_HD void init(int N_, int gl_, int G_, int ab_) {
        N = N_; gl = gl_; G = G_; ab = ab_;
        M = (N + (1 << ab) - 1) >> ab;
        c = 0;
        j = gl;
    }
    TX_HD bool next(uint32_t* i) {
        for (;;) {
            if (j >= M) {
                j = gl;
                ++c;
            }
            if (c >= (1 << ab) || j >= M) return false;
            const uint32_t idx = ((uint32_t)j << ab) + (uint32_t)c;
            j += G;
            if (idx < (uint32_t)N) {
                *i = idx;
                return true;
            }
        }
    }
};

struct IrtWork {
    const SceneView* s;
    Frame f;
    V3 n, o;
    float sh0, sh1;
    int mode;
    bool enabled;
    LaneSamples it;
    float ax, ay, az;
    TX_HD void init(const SceneView* s_, V3 n_, V3 o_, float sh0_, float sh1_, int N, int mode_, int gl, int G, int ab,
                    bool enabled_) {
        s = s_; n = n_; o = o_; sh0 = sh0_; sh1 = sh1_; mode = mode_; enabled = enabled_;
        f = make_frame(n_);
        it.init(N, gl, G, ab);
        ax = ay = az = 0.f;
    }
    TX_HD bool next(V3* org, V3* dir) {
        uint32_t i;
        if (!enabled || !it.next(&i)) return false;
        float s0 = shift_wrap_clamp(hammersley0(i, (uint32_t)it.N), sh0);
        float s1 = shift_wrap_clamp(radical_inverse(i), sh1);
        DirSample ds = sample_dir<false>(f, mode, s0, s1, 0.f);
        *org = o;
        *dir = ds.d;
        return true;
    }
    TX_HD void done(const Hit& h, V3 d) {
        V3 L = fetch_radiance(*s, h);
        float ndl = clamp01(rn_dot(n, d));
        ax += L.x * ndl;
        ay += L.y * ndl;
        az += L.z * ndl;
    }
};

// render(): specular importance samples first, then the optional Monte-Carlo diffuse samples
template <bool TANGENT>
struct ShadeWork {
    const SceneView* s;
    Frame f;
    SpecPoint sp;
    V3 n, o, view;
    float rough, shs0, shs1, shd0, shd1;
    int mode_spec, mode_diff;
    bool enabled, diff_phase;
    LaneSamples it_spec, it_diff;
    float w, dw;  // weight of the ray in flight
    ShadeLaneOut out;
    TX_HD void init(const SceneView* s_, V3 n_, V3 o_, V3 cam, float rough_, float eps, float shs0_, float shs1_,
                    int N_spec, int mode_spec_, int ab_spec, float shd0_, float shd1_, int N_diff, int mode_diff_,
                    int ab_diff, int gl, int G, bool enabled_) {
        s = s_; n = n_; o = o_; rough = rough_; shs0 = shs0_; shs1 = shs1_; shd0 = shd0_; shd1 = shd1_;
        mode_spec = mode_spec_; mode_diff = mode_diff_; enabled = enabled_;
        view = view_dir(cam, o_);
        f = make_frame(n_);
        sp = spec_point(n_, view, rough_, eps);
        it_spec.init(N_spec, gl, G, ab_spec);
        it_diff.init(N_diff > 0 ? N_diff : 0, gl, G, ab_diff);
        diff_phase = false;
        w = dw = 0.f;
        for (int k = 0; k < 3; ++k) out.spec[k] = out.dspec[k] = out.diff[k] = 0.f;
    }
    TX_HD bool next(V3* org, V3* dir) {
        if (!enabled) return false;
        uint32_t i;
        if (!diff_phase) {
            if (it_spec.next(&i)) {
                float s0 = shift_wrap_clamp(hammersley0(i, (uint32_t)it_spec.N), shs0);
                float s1 = shift_wrap_clamp(radical_inverse(i), shs1);
                DirSample ds = sample_dir<TANGENT>(f, mode_spec, s0, s1, rough);
                float vdh_raw;
                V3 l = reflect_dir(ds.d, view, &vdh_raw);
                dw = 0.f;
                w = spec_weight<TANGENT>(sp, f, ds, l, vdh_raw, &dw);
                *org = o;
                *dir = l;
                return true;
            }
            diff_phase = true;
        }
        if (it_diff.N > 0 && it_diff.next(&i)) {
            float s0 = shift_wrap_clamp(hammersley0(i, (uint32_t)it_diff.N), shd0);
            float s1 = shift_wrap_clamp(radical_inverse(i), shd1);
            DirSample ds = sample_dir<false>(f, mode_diff, s0, s1, 0.f);
            w = (mode_diff == MODE_COSINE) ? 1.0f : clamp01(rn_dot(n, ds.d));
            *org = o;
            *dir = ds.d;
            return true;
        }
        return false;
    }
    TX_HD void done(const Hit& h, V3 d) {
        (void)d;
        V3 L = fetch_radiance(*s, h);
        if (!diff_phase || it_diff.N == 0) {
            out.spec[0] += L.x * w; out.spec[1] += L.y * w; out.spec[2] += L.z * w;
            if (TANGENT) {
                out.dspec[0] += L.x * dw; out.dspec[1] += L.y * dw; out.dspec[2] += L.z * dw;
            }
        } else {
            out.diff[0] += L.x * w; out.diff[1] += L.y * w; out.diff[2] += L.z * w;
        }
    }
};

// per-point epilogue shared by forward and recompute-backward: diffuse factor such that
// diffuse = albedo * dfac  (irr/pi when IrT-lit; MC estimate, x2 for the uniform pdf, when relighting)
TX_HD void diffuse_factor(const float irr[3], const float diff_sum[3], int N_diff, int mode_diff, float dfac[3]) {
    if (N_diff > 0) {
        float k = ((mode_diff == MODE_COSINE) ? 1.0f : 2.0f) / (float)N_diff;
        dfac[0] = diff_sum[0] * k;
        dfac[1] = diff_sum[1] * k;
        dfac[2] = diff_sum[2] * k;
    } else {
        dfac[0] = irr[0] / kPi;
        dfac[1] = irr[1] / kPi;
        dfac[2] = irr[2] / kPi;
    }
}

}  // namespace texir
