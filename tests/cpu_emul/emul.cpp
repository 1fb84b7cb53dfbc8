// emul.cpp -- TEST INFRASTRUCTURE.  Compiles the kernels' host/device headers (texir_math.h,
// texir_bvh.h, texir_point.h) and the host BVH builder with g++ so that the exact per-sample code
// of the sm_100a kernels can be checked against the oracle on a machine without a GPU.
// Never loaded by the product package.  Build: see tests/cpu_emul/build.py (-ffp-contract=off).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../texir_code_b200/csrc/bvh_build.h"
#include "../../texir_code_b200/csrc/texir_point.h"
#include "../../texir_code_b200/csrc/texir_post.h"
#include "../../texir_code_b200/csrc/texir_tex.h"

using namespace texir;

struct EmulScene {
    HostBVH bvh;
    std::vector<double> uvs;
    std::vector<float4> tex;
    SceneView view{};
    std::string err;
};

static int group_lanes(int N) {
    int g = 1;
    while (g < N && g < 32) g <<= 1;
    return g;
}

extern "C" {

EmulScene* emul_scene_create(const float* verts, int64_t nv, const int32_t* tris, int64_t nt, const double* uvs,
                             int threads) {
    EmulScene* s = new EmulScene();
    if (!build_bvh(verts, nv, tris, nt, threads, s->bvh, s->err)) {
        delete s;
        return nullptr;
    }
    if (uvs) s->uvs.assign(uvs, uvs + 6 * (size_t)nt);
    s->view.nodes2 = s->bvh.nodes2.data();
    s->view.tris2 = s->bvh.tris2.data();
    s->view.qnodes2 = s->bvh.qnodes2.data();
    for (int a = 0; a < 3; ++a) {
        s->view.qmin[a] = s->bvh.qmin[a];
        s->view.qstep[a] = s->bvh.qstep[a];
    }
    s->view.nodes8 = s->bvh.nodes8.data();
    s->view.tris8 = s->bvh.tris8.data();
    s->view.tri_uvs = s->uvs.empty() ? nullptr : s->uvs.data();
    s->view.num_tris = (uint32_t)nt;
    return s;
}
void emul_scene_destroy(EmulScene* s) { delete s; }

void emul_scene_stats(EmulScene* s, double* out /* nodes2, nodes8, depth2, depth8, sah2, sah8, build_s */) {
    out[0] = (double)s->bvh.nodes2.size() / 4;
    out[1] = (double)s->bvh.nodes8.size() / 5;
    out[2] = s->bvh.depth2;
    out[3] = s->bvh.depth8;
    out[4] = s->bvh.sah2;
    out[5] = s->bvh.sah8;
    out[6] = s->bvh.build_seconds;
}

void emul_set_texture(EmulScene* s, const float* chw, int H, int W) {
    s->tex.resize((size_t)H * W);
    size_t n = (size_t)H * W;
    for (size_t i = 0; i < n; ++i) s->tex[i] = float4{chw[i], chw[n + i], chw[2 * n + i], 0.f};
    s->view.tex = s->tex.data();
    s->view.tex_h = H;
    s->view.tex_w = W;
}

void emul_cast_rays(EmulScene* s, int kind, const float* rays, int64_t R, float* t, uint32_t* id, float* uv,
                    uint32_t* nvis, uint32_t* ntri) {
    for (int64_t r = 0; r < R; ++r) {
        V3 o = v3(rays[6 * r], rays[6 * r + 1], rays[6 * r + 2]);
        V3 d = v3(rays[6 * r + 3], rays[6 * r + 4], rays[6 * r + 5]);
        uint32_t a = 0, b = 0;
        Hit h = kind == 8 ? trace_bvh8(s->view, o, d, &a, &b) : trace_bvh2(s->view, o, d, &a, &b);
        t[r] = h.t;
        id[r] = h.prim;
        uv[2 * r] = h.u;
        uv[2 * r + 1] = h.v;
        if (nvis) nvis[r] = a;
        if (ntri) ntri[r] = b;
    }
}

void emul_generate_dir(const float* nrm, const float* rough, const float* shift, int64_t P, int N, int mode,
                       float* out) {
    for (int64_t p = 0; p < P; ++p) {
        Frame f = make_frame(v3(nrm[3 * p], nrm[3 * p + 1], nrm[3 * p + 2]));
        for (int i = 0; i < N; ++i) {
            float s0 = shift_wrap_clamp(hammersley0((uint32_t)i, (uint32_t)N), shift[2 * p]);
            float s1 = shift_wrap_clamp(radical_inverse((uint32_t)i), shift[2 * p + 1]);
            DirSample ds = sample_dir<false>(f, mode, s0, s1, rough ? rough[p] : 0.f);
            float* o = out + 3 * (p * N + i);
            o[0] = ds.d.x;
            o[1] = ds.d.y;
            o[2] = ds.d.z;
        }
    }
}

void emul_query_irf(EmulScene* s, int kind, const float* org, const float* dir, int64_t P, int N, float* rad) {
    for (int64_t p = 0; p < P; ++p)
        for (int i = 0; i < N; ++i) {
            int64_t q = p * N + i;
            V3 o = v3(org[3 * p], org[3 * p + 1], org[3 * p + 2]);
            V3 d = v3(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2]);
            Hit h = kind == 8 ? trace<8>(s->view, o, d) : trace<2>(s->view, o, d);
            V3 L = fetch_radiance(s->view, h);
            rad[3 * q] = L.x;
            rad[3 * q + 1] = L.y;
            rad[3 * q + 2] = L.z;
        }
}

void emul_irradiance(EmulScene* s, int kind, const float* pos, const float* nrm, const float* shift, int64_t P,
                     int N, int mode, float* E) {
    const int G = group_lanes(N), ab = coherent_bits(N);
    for (int64_t p = 0; p < P; ++p) {
        float sum[3] = {0.f, 0.f, 0.f};
        for (int gl = 0; gl < G; ++gl) {
            float acc[3];
            V3 n = v3(nrm[3 * p], nrm[3 * p + 1], nrm[3 * p + 2]), o = v3(pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]);
            if (kind == 8)
                irradiance_lane<8>(s->view, n, o, shift[2 * p], shift[2 * p + 1], N, mode, gl, G, ab, acc);
            else
                irradiance_lane<2>(s->view, n, o, shift[2 * p], shift[2 * p + 1], N, mode, gl, G, ab, acc);
            for (int c = 0; c < 3; ++c) sum[c] += acc[c];
        }
        for (int c = 0; c < 3; ++c) E[3 * p + c] = irradiance_finish(sum[c], N);
    }
}

// kind 3 = BVH2 with the cheaper node step of texir_walk.h
void emul_cast_rays_fast(EmulScene* s, int quantised, const float* rays, int64_t R, float* t, uint32_t* id, float* uv) {
    for (int64_t r = 0; r < R; ++r) {
        V3 o = v3(rays[6 * r], rays[6 * r + 1], rays[6 * r + 2]), d = v3(rays[6 * r + 3], rays[6 * r + 4], rays[6 * r + 5]);
        Hit h = quantised ? trace_qbvh2(s->view, o, d) : trace_bvh2_fast(s->view, o, d);
        t[r] = h.t;
        id[r] = h.prim;
        uv[2 * r] = h.u;
        uv[2 * r + 1] = h.v;
    }
}

// forward render with the forward-mode roughness tangent; outputs rgb [P,3], drgb [P,3], ediff [P,3]
void emul_shade(EmulScene* s, int kind, const float* nrm, const float* albedo, const float* rough, const float* pos,
                const float* cam, const float* irr, const float* shift_spec, const float* shift_diff, int64_t P,
                int N_spec, int N_diff, int mode_spec, int mode_diff, float eps, float* rgb, float* drgb,
                float* ediff) {
    int gs = group_lanes(N_spec), gd = N_diff > 0 ? group_lanes(N_diff) : gs;
    const int G = gs > gd ? gs : gd;
    const int abs_ = coherent_bits(N_spec), abd = N_diff > 0 ? coherent_bits(N_diff) : 0;
    V3 c = v3(cam[0], cam[1], cam[2]);
    for (int64_t p = 0; p < P; ++p) {
        float sp[3] = {0, 0, 0}, dsp[3] = {0, 0, 0}, df[3] = {0, 0, 0};
        V3 n = v3(nrm[3 * p], nrm[3 * p + 1], nrm[3 * p + 2]), o = v3(pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]);
        for (int gl = 0; gl < G; ++gl) {
            ShadeLaneOut lo;
            float d0 = N_diff > 0 ? shift_diff[2 * p] : 0.f, d1 = N_diff > 0 ? shift_diff[2 * p + 1] : 0.f;
            if (kind == 8)
                shade_lane<8, true>(s->view, n, o, c, rough[p], eps, shift_spec[2 * p], shift_spec[2 * p + 1], N_spec,
                                    mode_spec, abs_, d0, d1, N_diff, mode_diff, abd, gl, G, lo);
            else
                shade_lane<2, true>(s->view, n, o, c, rough[p], eps, shift_spec[2 * p], shift_spec[2 * p + 1], N_spec,
                                    mode_spec, abs_, d0, d1, N_diff, mode_diff, abd, gl, G, lo);
            for (int k = 0; k < 3; ++k) {
                sp[k] += lo.spec[k];
                dsp[k] += lo.dspec[k];
                df[k] += lo.diff[k];
            }
        }
        float ir[3] = {0, 0, 0}, dfac[3];
        if (N_diff == 0)
            for (int k = 0; k < 3; ++k) ir[k] = irr[3 * p + k];
        diffuse_factor(ir, df, N_diff, mode_diff, dfac);
        for (int k = 0; k < 3; ++k) {
            rgb[3 * p + k] = albedo[3 * p + k] * dfac[k] + sp[k] / (float)N_spec;
            drgb[3 * p + k] = dsp[k] / (float)N_spec;
            ediff[3 * p + k] = dfac[k];
        }
    }
}

// ---- IrT post-processing (texir_post.h): the two passes of k_pad_columns / k_pad_rows, texel by texel ----
void emul_pad_texture(const float* img, int H, int W, int C, float* out, int32_t* index) {
    std::vector<int> near((size_t)H * W);
    std::vector<unsigned short> g((size_t)W), nr((size_t)W);
    for (int j = 0; j < W; ++j) {
        int last = -1;
        for (int i = 0; i < H; ++i) {
            if (!pad_texel_empty(img + ((size_t)i * W + j) * C, C)) last = i;
            near[(size_t)i * W + j] = last;
        }
        int next = -1;
        for (int i = H - 1; i >= 0; --i) {
            const int up = near[(size_t)i * W + j];
            if (up == i) next = i;
            near[(size_t)i * W + j] = pad_pick_row(i, up, next);
        }
    }
    for (int i = 0; i < H; ++i) {
        for (int x = 0; x < W; ++x) nr[x] = near[(size_t)i * W + x] < 0 ? kPadNone : (unsigned short)near[(size_t)i * W + x];
        const int maxl = pad_build_envelope(nr.data(), W, i, g.data());
        for (int x = 0; x < W; ++x) {
            int sr = i, sc = x, rr = i, rc = x;
            if ((int)nr[x] != i && maxl >= 0) {
                sc = pad_query_envelope(nr.data(), g.data(), maxl, i, x);
                sr = nr[sc];
                rr = pad_source_index(sr, H);
                rc = pad_source_index(sc, W);
            }
            for (int c = 0; c < C; ++c)
                out[((size_t)i * W + x) * C + c] = (rr >= 0 && rc >= 0) ? img[((size_t)rr * W + rc) * C + c] : 0.0f;
            index[2 * ((size_t)i * W + x) + 0] = sr;
            index[2 * ((size_t)i * W + x) + 1] = sc;
        }
    }
}
void emul_rgbe_encode(const float* img, int64_t n, int bgr, unsigned char* rgbe) {
    for (int64_t i = 0; i < n; ++i)
        rgbe_from_float(img[3 * i + (bgr ? 2 : 0)], img[3 * i + 1], img[3 * i + (bgr ? 0 : 2)], rgbe + 4 * i);
}
void emul_rgbe_decode(const unsigned char* rgbe, int64_t n, int bgr, float* img) {
    for (int64_t i = 0; i < n; ++i)
        rgbe_to_float(rgbe + 4 * i, img + 3 * i + (bgr ? 2 : 0), img + 3 * i + 1, img + 3 * i + (bgr ? 0 : 2));
}

// ---- mip-mapped material-texture fetch (texir_tex.h), texel arithmetic of k_mip_build_level / k_texture_fetch_mip ----
int emul_mip_levels(int H, int W, int max_mip_level) { return mip_level_count(H, W, max_mip_level); }
int64_t emul_mip_elems(int H, int W, int C, int levels) { return mip_offset(H, W, C, levels + 1); }
void emul_mip_build(const float* tex, int H, int W, int C, int levels, float* mips) {
    const float* src = tex;
    for (int l = 1; l <= levels; ++l) {
        float* dst = mips + mip_offset(H, W, C, l);
        const int hs = mip_dim(H, l - 1), ws = mip_dim(W, l - 1), hd = mip_dim(H, l), wd = mip_dim(W, l);
        const int sx = ws > 1 ? 2 : 1, sy = hs > 1 ? 2 : 1;
        for (int y = 0; y < hd; ++y)
            for (int x = 0; x < wd; ++x)
                for (int c = 0; c < C; ++c) {
                    float acc = 0.f;
                    for (int dy = 0; dy < sy; ++dy)
                        for (int dx = 0; dx < sx; ++dx) acc += src[((size_t)(y * sy + dy) * ws + (x * sx + dx)) * C + c];
                    dst[((size_t)y * wd + x) * C + c] = acc * (1.0f / (float)(sx * sy));
                }
        src = dst;
    }
}
void emul_texture_fetch_mip(const float* tex, const float* mips, int H, int W, int C, int L, const float* uv,
                            const float* uv_da, int64_t P, float* out, float* lod_out) {
    for (int64_t p = 0; p < P; ++p) {
        const MipLod m = mip_lod(uv_da[4 * p], uv_da[4 * p + 1], uv_da[4 * p + 2], uv_da[4 * p + 3], H, W, L);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < 2; ++k) {
            const int l = k ? m.level1 : m.level0;
            const float wl = k ? m.frac : 1.0f - m.frac;
            if (k && m.frac == 0.f) break;
            const int h = mip_dim(H, l), w = mip_dim(W, l);
            const float* t = l ? mips + mip_offset(H, W, C, l) : tex;
            const TexTap tp = tex_tap(uv[2 * p], uv[2 * p + 1], h, w);
            for (int c = 0; c < C; ++c)
                acc[c] += wl * (t[(size_t)tp.i00 * C + c] * tp.w00 + t[(size_t)tp.i10 * C + c] * tp.w10 +
                                t[(size_t)tp.i01 * C + c] * tp.w01 + t[(size_t)tp.i11 * C + c] * tp.w11);
        }
        for (int c = 0; c < C; ++c) out[p * C + c] = acc[c];
        if (lod_out) lod_out[p] = (float)m.level0 + m.frac;
    }
}

// ---- divergence analysis (profiles/simulate_warp_divergence.py): for every ray the number of QBVH2 node steps between
// consecutive leaf visits ("segments"), i.e. what one lane does between two reconvergence points of the while-while loop
void emul_trace_segments(EmulScene* s, const float* rays, int64_t R, int max_seg, int32_t* n_seg, int32_t* seg,
                         int32_t* leaf_tris) {
    for (int64_t q = 0; q < R; ++q) {
        RayQ r;
        rayq_setup(s->view, r, v3(rays[6 * q], rays[6 * q + 1], rays[6 * q + 2]),
                   v3(rays[6 * q + 3], rays[6 * q + 4], rays[6 * q + 5]));
        Hit hit;
        hit_reset(hit);
        int stack[kStackSize];
        int* sp = stack;
        int top = kQSentinel, node = 0, k = 0;
        *sp++ = kQSentinel;
        while (node != kQSentinel) {
            int steps = 0;
            while (node >= 0) {
                node = qbvh2_step(s->view, r, hit.t, node, sp, top);
                ++steps;
            }
            int cnt = 0;
            if (node != kQSentinel) {
                const int ref = ~node;
                const int first = ref >> 3;
                cnt = (ref & 7) + 1;
                for (int i = 0; i < cnt; ++i) test_slot(s->view.tris2, first + i, r.o, r.d, hit);
                node = qstack_pop(sp, top);
            }
            if (k < max_seg) {
                seg[q * max_seg + k] = steps;
                leaf_tris[q * max_seg + k] = cnt;
            }
            ++k;
        }
        n_seg[q] = k;
    }
}

// Lock-step model of one warp walking `lanes` rays (<= 32) with a per-lane queue of `Q` postponed leaves:
//   Q = 0  the kernels' while-while loop: a lane that reaches a leaf waits until no lane is at an inner node;
//   Q >= 1 speculative traversal (Aila & Laine 2009): the lane parks the leaf and keeps walking with its old t_best;
//          with stop_when_all_found the node phase ends as soon as every unfinished lane has parked a leaf.
// out[0] = warp node iterations, out[1] = warp triangle iterations, out[2] = lane node steps, out[3] = lane triangle tests
void emul_simulate_warp(EmulScene* s, const float* rays, int lanes, int Q, int stop_when_all_found, int64_t* out) {
    struct Lane {
        RayQ r;
        Hit hit;
        int stack[kStackSize];
        int* sp;
        int top, node;
        int queue[8], nq;
    };
    std::vector<Lane> L((size_t)lanes);
    for (int q = 0; q < lanes; ++q) {
        Lane& l = L[q];
        rayq_setup(s->view, l.r, v3(rays[6 * q], rays[6 * q + 1], rays[6 * q + 2]),
                   v3(rays[6 * q + 3], rays[6 * q + 4], rays[6 * q + 5]));
        hit_reset(l.hit);
        l.sp = l.stack;
        *l.sp++ = kQSentinel;
        l.top = kQSentinel;
        l.node = 0;
        l.nq = 0;
    }
    auto pop = [](Lane& l) { return qstack_pop(l.sp, l.top); };
    int64_t node_it = 0, tri_it = 0, node_steps = 0, tri_tests = 0;
    for (;;) {
        bool any_alive = false;
        for (auto& l : L) any_alive |= (l.node != kQSentinel || l.nq > 0);
        if (!any_alive) break;
        // node phase
        for (;;) {
            bool stepped = false;
            for (auto& l : L) {
                if (l.node >= 0) {
                    l.node = qbvh2_step(s->view, l.r, l.hit.t, l.node, l.sp, l.top);
                    ++node_steps;
                    stepped = true;
                    while (l.node < 0 && l.node != kQSentinel && l.nq < Q) {  // park the leaf, keep walking
                        l.queue[l.nq++] = l.node;
                        l.node = pop(l);
                    }
                }
            }
            if (!stepped) break;
            ++node_it;
            if (stop_when_all_found) {
                bool searching = false;
                for (auto& l : L) searching |= (l.node >= 0 && l.nq == 0);
                if (!searching) break;
            }
        }
        // leaf phase: every lane tests what it has parked (and the leaf it is blocked on)
        int64_t worst = 0;
        for (auto& l : L) {
            int64_t mine = 0;
            if (l.node < 0 && l.node != kQSentinel) {
                l.queue[l.nq++] = l.node;  // queue has Q + 1 slots
                l.node = pop(l);
            }
            for (int k = 0; k < l.nq; ++k) {
                const int ref = ~l.queue[k];
                const int first = ref >> 3, cnt = (ref & 7) + 1;
                for (int i = 0; i < cnt; ++i) test_slot(s->view.tris2, first + i, l.r.o, l.r.d, l.hit);
                mine += cnt;
            }
            l.nq = 0;
            tri_tests += mine;
            worst = mine > worst ? mine : worst;
        }
        tri_it += worst;
    }
    out[0] = node_it;
    out[1] = tri_it;
    out[2] = node_steps;
    out[3] = tri_tests;
}

// Two more lock-step models over a pool of `total` rays handed to 32 lanes:
//   mode 0  refill at round boundaries: a lane whose ray has finished takes the next ray of the pool when the warp
//           reconverges after a leaf phase (a ray stream with warp-wide refill points);
//   mode 1  two rays per lane: in the node phase a lane steps its first ray that stands at an inner node, so it idles
//           only when both of its rays wait at a leaf (or are finished); pairs are replaced when both are finished
//           and all lanes have finished theirs (no refill).
// out[0] = warp node iterations, out[1] = warp triangle iterations, out[2] = rounds, out[3] = rounds in which a lane finished
void emul_simulate_stream(EmulScene* s, const float* rays, int total, int mode, int64_t* out) {
    struct Ray {
        RayQ r;
        Hit hit;
        int stack[kStackSize];
        int* sp;
        int top, node;
        bool live;
    };
    const int per_lane = mode == 1 ? 2 : 1;
    std::vector<Ray> R((size_t)32 * per_lane);
    int next = 0;
    auto start = [&](Ray& l) {
        if (next >= total) {
            l.live = false;
            l.node = kQSentinel;
            return;
        }
        const float* q = rays + 6 * (size_t)next++;
        rayq_setup(s->view, l.r, v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]));
        hit_reset(l.hit);
        l.sp = l.stack;
        *l.sp++ = kQSentinel;
        l.top = kQSentinel;
        l.node = 0;
        l.live = true;
    };
    auto pop = [](Ray& l) { return qstack_pop(l.sp, l.top); };
    for (auto& l : R) start(l);
    int64_t node_it = 0, tri_it = 0, rounds = 0, finishing_rounds = 0;
    for (;;) {
        bool any = false;
        for (auto& l : R) any |= l.live;
        if (!any) break;
        for (;;) {  // node phase
            bool stepped = false;
            for (int lane = 0; lane < 32; ++lane)
                for (int k = 0; k < per_lane; ++k) {
                    Ray& l = R[(size_t)lane * per_lane + k];
                    if (l.live && l.node >= 0) {
                        l.node = qbvh2_step(s->view, l.r, l.hit.t, l.node, l.sp, l.top);
                        stepped = true;
                        break;  // one step per lane per iteration
                    }
                }
            if (!stepped) break;
            ++node_it;
        }
        int64_t worst = 0;  // leaf phase
        bool finished = false;
        for (int lane = 0; lane < 32; ++lane) {
            int64_t mine = 0;
            for (int k = 0; k < per_lane; ++k) {
                Ray& l = R[(size_t)lane * per_lane + k];
                if (!l.live) continue;
                if (l.node < 0 && l.node != kQSentinel) {
                    const int ref = ~l.node;
                    const int first = ref >> 3, cnt = (ref & 7) + 1;
                    for (int i = 0; i < cnt; ++i) test_slot(s->view.tris2, first + i, l.r.o, l.r.d, l.hit);
                    mine += cnt;
                    l.node = pop(l);
                }
                if (l.node == kQSentinel) {
                    l.live = false;
                    finished = true;
                }
            }
            worst = mine > worst ? mine : worst;
        }
        tri_it += worst;
        ++rounds;
        finishing_rounds += finished;
        if (mode == 0) {
            for (auto& l : R)
                if (!l.live) start(l);
        } else {
            bool all_done = true;
            for (auto& l : R) all_done &= !l.live;
            if (all_done)
                for (auto& l : R) start(l);
        }
    }
    out[0] = node_it;
    out[1] = tri_it;
    out[2] = rounds;
    out[3] = finishing_rounds;
}

// closest hits of R rays that all start at org[0..3), through the shared first descent (texir_walk.h)
void emul_cast_rays_front(EmulScene* s, const float* org, const float* dirs, int64_t R, float* t, uint32_t* id, float* uv,
                          int32_t* front_info) {
    uint4 entries[kMaxFront];
    int count = 0, start = 0;
    const V3 o = v3(org[0], org[1], org[2]);
    qbvh2_point_front(s->view, o, entries, &count, &start);
    if (front_info) {
        front_info[0] = count;
        front_info[1] = start;
    }
    for (int64_t q = 0; q < R; ++q) {
        Hit h = trace_qbvh2_front(s->view, o, v3(dirs[3 * q], dirs[3 * q + 1], dirs[3 * q + 2]), entries, count, start);
        t[q] = h.t;
        id[q] = h.prim;
        uv[2 * q] = h.u;
        uv[2 * q + 1] = h.v;
    }
}

// closest hits of R rays through the free-running walk (texir_walk.h, trace_qbvh2_free): the G-buffer kernel's loop
void emul_cast_rays_free(EmulScene* s, const float* rays, int64_t R, float* t, uint32_t* id, float* uv) {
    for (int64_t q = 0; q < R; ++q) {
        const float* r = rays + 6 * q;
        Hit h = trace_qbvh2_free(s->view, v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]));
        t[q] = h.t;
        id[q] = h.prim;
        uv[2 * q] = h.u;
        uv[2 * q + 1] = h.v;
    }
}

// What a traversal stack that also remembers each postponed child's entry distance would save: out[0] = node steps,
// out[1] = pops of inner nodes, out[2] = of those, pops whose stored tnear already exceeds the closest hit (today they
// cost a whole node step that finds neither child), out[3] = pops of leaves, out[4] = of those, culled the same way
// (today: their triangles are tested), out[5] = triangle tests.   (profiles/simulate_stack_tnear.py)
void emul_pop_stats(EmulScene* es, const float* rays, int64_t R, int64_t* out) {
    const SceneView& s = es->view;
    for (int k = 0; k < 6; ++k) out[k] = 0;
    for (int64_t q = 0; q < R; ++q) {
        RayQ r;
        rayq_setup(s, r, v3(rays[6 * q], rays[6 * q + 1], rays[6 * q + 2]), v3(rays[6 * q + 3], rays[6 * q + 4], rays[6 * q + 5]));
        Hit hit;
        hit_reset(hit);
        int stack[kStackSize];
        float tstack[kStackSize];
        int n = 0;
        int node = 0;
        const uint32_t fx = r.selx ^ 0x0022u, fy = r.sely ^ 0x0022u, fz = r.selz ^ 0x0022u;
        for (;;) {
            while (node >= 0) {
                ++out[0];
                const uint4* np = s.qnodes2 + 2 * (size_t)node;
                const uint4 q0 = np[0], q1 = np[1];
                const float t0n = fmaxf(fmaxf(qplane_(q0.x, r.selx, r.S.x, r.O.x), qplane_(q0.y, r.sely, r.S.y, r.O.y)),
                                        fmaxf(qplane_(q0.z, r.selz, r.S.z, r.O.z), 0.0f));
                const float t0f = fminf(fminf(qplane_(q0.x, fx, r.S.x, r.O.x), qplane_(q0.y, fy, r.S.y, r.O.y)),
                                        fminf(qplane_(q0.z, fz, r.S.z, r.O.z), hit.t));
                const float t1n = fmaxf(fmaxf(qplane_(q0.w, r.selx, r.S.x, r.O.x), qplane_(q1.x, r.sely, r.S.y, r.O.y)),
                                        fmaxf(qplane_(q1.y, r.selz, r.S.z, r.O.z), 0.0f));
                const float t1f = fminf(fminf(qplane_(q0.w, fx, r.S.x, r.O.x), qplane_(q1.x, fy, r.S.y, r.O.y)),
                                        fminf(qplane_(q1.y, fz, r.S.z, r.O.z), hit.t));
                const int c0 = (int)q1.z, c1 = (int)q1.w;
                const bool h0 = t0n <= t0f, h1 = t1n <= t1f;
                if (h0 && h1) {
                    const bool swap = t1n < t0n;
                    stack[n] = swap ? c0 : c1;
                    tstack[n++] = swap ? t0n : t1n;
                    node = swap ? c1 : c0;
                } else if (h0) {
                    node = c0;
                } else if (h1) {
                    node = c1;
                } else {
                    if (n == 0) { node = kQSentinel; break; }
                    --n;
                    node = stack[n];
                    if (node >= 0) { ++out[1]; out[2] += tstack[n] > hit.t; } else { ++out[3]; out[4] += tstack[n] > hit.t; }
                }
            }
            if (node == kQSentinel) break;
            const int ref = ~node, first = ref >> 3, cnt = (ref & 7) + 1;
            for (int i = 0; i < cnt; ++i) test_slot(s.tris2, first + i, r.o, r.d, hit);
            out[5] += cnt;
            if (n == 0) break;
            --n;
            node = stack[n];
            if (node >= 0) { ++out[1]; out[2] += tstack[n] > hit.t; } else { ++out[3]; out[4] += tstack[n] > hit.t; }
        }
    }
}

// closest hits through the walk whose stack entries carry their entry distance (texir_walk.h, trace_qbvh2_t: the render
// kernels' walk); use_front != 0: all rays start at org[0..3) and go through the shared first descent
void emul_cast_rays_tnear(EmulScene* s, const float* rays, int64_t R, int use_front, float* t, uint32_t* id, float* uv) {
    uint4 entries[kMaxFront];
    int count = 0, start = 0;
    if (use_front && R > 0) qbvh2_point_front(s->view, v3(rays[0], rays[1], rays[2]), entries, &count, &start);
    for (int64_t q = 0; q < R; ++q) {
        const float* r = rays + 6 * q;
        Hit h = trace_qbvh2_t(s->view, v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), entries, count, start);
        t[q] = h.t;
        id[q] = h.prim;
        uv[2 * q] = h.u;
        uv[2 * q + 1] = h.v;
    }
}

}  // extern "C"
