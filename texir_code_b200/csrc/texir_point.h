// texir_point.h -- the per-lane sample loops of the fused kernels (host + device).
//
// A group of G lanes owns one surface point; lane `gl` of the group runs these loops over its share
// of the N samples and the caller reduces the partial sums over the group (warp shuffles on the
// device, a plain loop in tests/cpu_emul).  Sample order inside a group follows coherent_bits().
#pragma once
#include "texir_bvh.h"
#include "texir_walk.h"

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
