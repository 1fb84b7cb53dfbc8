// texir_bvh.h -- flattened BVH layouts in HBM and the closest-hit traversals over them.
//
// Replaces Open3D's RaycastingScene.cast_rays (Embree, CPU) at the reference call sites
// models/mat_nvdiffrast.py:98-101,294-295 and models/tracer_o3d_irt.py:85-89,243-244.
// Contract: closest t over all triangles, 0 < t < inf, no culling, ties -> lowest primitive id,
// (u, v) weight vertices 1 and 2; miss -> t = inf, id = 0xFFFFFFFF, u = v = 0.
//
// Two layouts are built by the host builder (bvh_build.cpp):
//
//  BVH2  64-byte inner nodes (4 x float4) holding BOTH children's boxes:
//          n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = same for child 1
//          n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   n3 = (child0, child1, -, -) as int bits
//        child >= 0: inner node index; child < 0: leaf, ~child = (first_slot << 3) | (count-1).
//
//  QBVH2 the same tree with 32-byte nodes (2 x uint4): the six planes of each child box as 16-bit indices
//        on a global grid  x = qmin + q * qstep  (per axis), low half = lo plane, high half = hi plane:
//          q0 = (c0.x, c0.y, c0.z, c1.x)   q1 = (c1.y, c1.z, child0, child1)
//        Half the L1 / L2 traffic per node visit of the float layout; this is the layout the kernels walk.
//
//  BVH8  80-byte compressed wide nodes (5 x uint4), 8 children with 8-bit quantised boxes on a
//        per-node power-of-two grid (Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal
//        on GPUs Through Compressed Wide BVHs", HPG 2017).  See BVH8 section below.
//
//  Triangles (both): slot-ordered, 48 bytes = 3 x float4:
//          (v0.x v0.y v0.z e1.x) (e1.y e1.z e2.x e2.y) (e2.z prim - -)
//        with e1 = v0 - v1, e2 = v2 - v0 and the original primitive id as int bits in the same record, so that
//        accepting a hit needs no second, dependent load; Ng = e2 x e1 is recomputed per test with the same
//        fma chain the builder and the oracle use (three FMA-pipe instructions, the pipe with headroom).
#pragma once
#include "texir_math.h"

#if defined(__CUDACC__)
#include <vector_types.h>
#else
struct alignas(16) float4 {
    float x, y, z, w;
};
struct alignas(16) uint4 {
    unsigned int x, y, z, w;
};
#endif

#ifdef __CUDA_ARCH__
#define TX_LDG(p) __ldg(p)
#else
#define TX_LDG(p) (*(p))
#endif

// Leaf-level data (triangles, per-corner UVs, radiance texels) is touched once per ray with no reuse inside the SM;
// TX_LEAF_NOALLOC=1 loads it with L1::no_allocate so that it does not push BVH nodes out of L1.
#ifndef TX_LEAF_NOALLOC
#define TX_LEAF_NOALLOC 0
#endif
#if defined(__CUDA_ARCH__) && TX_LEAF_NOALLOC
__device__ __forceinline__ float4 tx_ld_leaf(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 tx_ld_leaf(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
#define TX_LD_LEAF(p) tx_ld_leaf(p)
#else
#define TX_LD_LEAF(p) TX_LDG(p)
#endif

namespace texir {

constexpr int kStackSize = 64;
constexpr int kSentinel = 0x76543210;
constexpr int kStack8 = 48;

struct SceneView {
    const float4* __restrict__ nodes2;     // BVH2 nodes, 4 float4 each
    const float4* __restrict__ tris2;      // BVH2-ordered triangle slots, 3 float4 each
    const uint4* __restrict__ qnodes2;     // QBVH2 nodes, 2 uint4 each (same indices as nodes2)
    float qmin[3], qstep[3];               // the global 16-bit grid of qnodes2
    const uint4* __restrict__ nodes8;      // BVH8 nodes, 5 uint4 each
    const float4* __restrict__ tris8;      // BVH8-ordered triangle slots
    const double* __restrict__ tri_uvs;    // [T][3][2] float64, original primitive order
    const float4* __restrict__ tex;        // radiance texture, RGBA float4 [H][W]
    int tex_w, tex_h;
    uint32_t num_tris;
    int top_nodes;                         // qnodes2[0 .. top_nodes) are the top levels of the tree, breadth first
    const uint32_t* __restrict__ tex_rgbe; // the same texture as 4-byte Radiance RGBE texels (R | G<<8 | B<<16 | E<<24), or null
    int tex_rgbe_bias;                     // hdr_exposure folded into the shared exponent: value = mantissa * 2^(E - 136 + bias)
};

TX_HD float int_as_float_(int i) {
    union {
        int i;
        float f;
    } c;
    c.i = i;
    return c.f;
}
TX_HD int float_as_int_(float f) {
    union {
        int i;
        float f;
    } c;
    c.f = f;
    return c.i;
}

// reciprocal direction for slab tests; |d| is floored at 2^-80 so that no product is inf*0
TX_HD V3 safe_rcp_dir(V3 d) {
    const float tiny = 8.271806125530277e-25f;  // 2^-80
    float dx = fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x);
    float dy = fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y);
    float dz = fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z);
    return v3(1.0f / dx, 1.0f / dy, 1.0f / dz);
}

// slab test with exact subtraction before the multiply and a 3-ulp widening, so that a box
// (already padded by the builder) is never missed by rounding.
TX_HD bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, V3 o, V3 id, float tbest,
                float* tnear) {
    float ax = (lox - o.x) * id.x, bx = (hix - o.x) * id.x;
    float ay = (loy - o.y) * id.y, by = (hiy - o.y) * id.y;
    float az = (loz - o.z) * id.z, bz = (hiz - o.z) * id.z;
    float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tbest));
    *tnear = tmin;
    return tmin * 0.9999996f <= tmax * 1.0000004f;
}

TX_HD void test_slot(const float4* __restrict__ tris, int slot, V3 o, V3 d, Hit& hit) {
    float4 a = TX_LD_LEAF(&tris[3 * slot + 0]);
    float4 b = TX_LD_LEAF(&tris[3 * slot + 1]);
    float4 c = TX_LD_LEAF(&tris[3 * slot + 2]);
    const V3 e1 = v3(a.w, b.x, b.y), e2 = v3(b.z, b.w, c.x);
    float t, u, v;
    if (tri_test(v3(a.x, a.y, a.z), e1, e2, fcross(e2, e1), o, d, &t, &u, &v)) {
        if (t <= hit.t) {
            uint32_t p = (uint32_t)float_as_int_(c.y);
            if (t < hit.t || p < hit.prim) {
                hit.t = t;
                hit.u = u;
                hit.v = v;
                hit.prim = p;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// BVH2 closest hit, one ray per thread, while-while loop with a per-thread stack
// (structure after Aila & Laine, "Understanding the Efficiency of Ray Traversal on GPUs", HPG 2009)
// ---------------------------------------------------------------------------------------------
TX_HD Hit trace_bvh2(const SceneView& s, V3 o, V3 d, uint32_t* nodes_visited = nullptr,
                     uint32_t* tris_tested = nullptr) {
    Hit hit;
    hit.t = INFINITY;
    hit.u = 0.f;
    hit.v = 0.f;
    hit.prim = kInvalidId;
    V3 id = safe_rcp_dir(d);
    int stack[kStackSize];
    int sp = 0;
    stack[sp++] = kSentinel;
    int node = 0;
    uint32_t nv = 0, nt = 0;
    while (node != kSentinel) {
        while (node >= 0 && node != kSentinel) {
            const float4* np = s.nodes2 + 4 * (size_t)node;
            float4 n0 = TX_LDG(np + 0), n1 = TX_LDG(np + 1), n2 = TX_LDG(np + 2), n3 = TX_LDG(np + 3);
            ++nv;
            float t0, t1;
            bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, o, id, hit.t, &t0);
            bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, o, id, hit.t, &t1);
            int c0 = float_as_int_(n3.x), c1 = float_as_int_(n3.y);
            if (!h0 && !h1) {
                node = stack[--sp];
            } else if (h0 && h1) {
                bool swap = t1 < t0;
                node = swap ? c1 : c0;
                stack[sp++] = swap ? c0 : c1;
            } else {
                node = h0 ? c0 : c1;
            }
        }
        if (node < 0) {
            int ref = ~node;
            int first = ref >> 3, cnt = (ref & 7) + 1;
            for (int i = 0; i < cnt; ++i) test_slot(s.tris2, first + i, o, d, hit);
            nt += cnt;
            node = stack[--sp];
        }
    }
    if (nodes_visited) *nodes_visited = nv;
    if (tris_tested) *tris_tested = nt;
    return hit;
}

// ---------------------------------------------------------------------------------------------
// BVH8 compressed wide node (80 bytes = 5 x uint4)
//   q0 = (p.x, p.y, p.z as float bits, ex | ey<<8 | ez<<16 | imask<<24)
//   q1 = (child_base, tri_base, meta[0..3], meta[4..7])
//   q2 = (qlox[0..3], qlox[4..7], qloy[0..3], qloy[4..7])
//   q3 = (qloz[0..3], qloz[4..7], qhix[0..3], qhix[4..7])
//   q4 = (qhiy[0..3], qhiy[4..7], qhiz[0..3], qhiz[4..7])
// child box i: lo = p + 2^e * qlo[i], hi = p + 2^e * qhi[i] (conservative).
// meta[i]: 0 = empty slot; inner child: (1 << 5) | (24 + i), its node index is child_base +
// popc(imask & ((1 << i) - 1)); leaf: high 3 bits = unary count (1 -> 001, 2 -> 011, 3 -> 111),
// low 5 bits = offset of its first triangle from tri_base.
// ---------------------------------------------------------------------------------------------
TX_HD uint32_t popc_(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
TX_HD int ffs_msb_(uint32_t x) {  // index of highest set bit, x != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz(x);
#else
    return 31 - __builtin_clz(x);
#endif
}
TX_HD uint32_t extract_byte_(uint32_t x, int i) { return (x >> (8 * i)) & 0xFFu; }

// Straightforward (unordered-by-octant) BVH8 traversal: children hit by the ray are pushed as
// (node group: base index + bitmask of hit inner children) and (triangle group: base + bitmask).
// Inner children are visited nearest-first within a node by scanning the hit mask in order of
// entry distance.  Stack entries are 64-bit (base, mask) pairs.
struct Group8 {
    uint32_t base;
    uint32_t mask;
};

TX_HD Hit trace_bvh8(const SceneView& s, V3 o, V3 d, uint32_t* nodes_visited = nullptr,
                     uint32_t* tris_tested = nullptr) {
    Hit hit;
    hit.t = INFINITY;
    hit.u = 0.f;
    hit.v = 0.f;
    hit.prim = kInvalidId;
    V3 id = safe_rcp_dir(d);
    // Ordered traversal: the builder stores children so that slot bit a (0 = x, 1 = y, 2 = z) set
    // means "on the positive side along axis a".  A ray visits slots in descending (slot ^ oct_inv),
    // where oct_inv has bit a set iff d[a] >= 0  (so a (+,+,+) ray visits slot 0 first).
    const uint32_t oct_inv = (d.x < 0.f ? 0u : 1u) | (d.y < 0.f ? 0u : 2u) | (d.z < 0.f ? 0u : 4u);
    Group8 stack[kStack8];
    int sp = 0;
    // current node group: base = child_base of the parent; mask bits 24..31 = hit inner children
    // (bit 24 + (slot ^ oct_inv)), bits 0..7 = the parent's imask.  The root is a pseudo-group with
    // base 0, imask 0 and priority bit 31, which resolves to node 0.
    Group8 ng;
    ng.base = 0;
    ng.mask = 0x80000000u;
    Group8 tg;  // current triangle group: base = tri_base, mask = triangles still to test
    tg.base = 0;
    tg.mask = 0;
    uint32_t nv = 0, nt = 0;
    for (;;) {
        if (ng.mask > 0x00FFFFFFu) {
            // pop the highest-priority hit child of the node group
            int bit = ffs_msb_(ng.mask);
            ng.mask &= ~(1u << bit);
            uint32_t slot = (uint32_t)(bit - 24) ^ oct_inv;
            uint32_t node_idx = ng.base + popc_((ng.mask & 0xFFu) & ((1u << slot) - 1u));
            if (ng.mask > 0x00FFFFFFu) stack[sp++] = ng;  // remaining siblings
            const uint4* np = s.nodes8 + 5 * (size_t)node_idx;
            uint4 q0 = TX_LDG(np + 0), q1 = TX_LDG(np + 1), q2 = TX_LDG(np + 2), q3 = TX_LDG(np + 3),
                  q4 = TX_LDG(np + 4);
            ++nv;
            float px = int_as_float_((int)q0.x), py = int_as_float_((int)q0.y), pz = int_as_float_((int)q0.z);
            uint32_t ex = extract_byte_(q0.w, 0), ey = extract_byte_(q0.w, 1), ez = extract_byte_(q0.w, 2);
            uint32_t imask = extract_byte_(q0.w, 3);
            // scale = 2^e * idir ; origin term = (p - o) * idir  -> t = q * scale + org
            float sx = int_as_float_((int)(ex << 23)) * id.x;
            float sy = int_as_float_((int)(ey << 23)) * id.y;
            float sz = int_as_float_((int)(ez << 23)) * id.z;
            float ox = (px - o.x) * id.x, oy = (py - o.y) * id.y, oz = (pz - o.z) * id.z;
            uint32_t hit_inner = 0, hit_tris = 0;
            const uint32_t qlox[2] = {q2.x, q2.y}, qloy[2] = {q2.z, q2.w}, qloz[2] = {q3.x, q3.y};
            const uint32_t qhix[2] = {q3.z, q3.w}, qhiy[2] = {q4.x, q4.y}, qhiz[2] = {q4.z, q4.w};
            const uint32_t meta[2] = {q1.z, q1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t m = extract_byte_(meta[i >> 2], i & 3);
                if (m == 0) continue;
                float lx = (float)extract_byte_(qlox[i >> 2], i & 3), hx = (float)extract_byte_(qhix[i >> 2], i & 3);
                float ly = (float)extract_byte_(qloy[i >> 2], i & 3), hy = (float)extract_byte_(qhiy[i >> 2], i & 3);
                float lz = (float)extract_byte_(qloz[i >> 2], i & 3), hz = (float)extract_byte_(qhiz[i >> 2], i & 3);
                float ax = lx * sx + ox, bx = hx * sx + ox;
                float ay = ly * sy + oy, by = hy * sy + oy;
                float az = lz * sz + oz, bz = hz * sz + oz;
                float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
                float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), hit.t));
                if (tmin * 0.9999996f <= tmax * 1.0000004f) {
                    if ((m >> 5) == 1u && (m & 31u) >= 24u) {
                        // inner child in slot i: priority bit 24 + (i ^ oct_inv)
                        hit_inner |= 1u << (24 + ((uint32_t)i ^ oct_inv));
                    } else {
                        hit_tris |= (m >> 5) << (m & 31u);
                    }
                }
            }
            ng.base = q1.x;
            ng.mask = hit_inner | imask;
            tg.base = q1.y;
            tg.mask = hit_tris;
        } else {
            tg = ng;  // a triangle group popped from the stack (mask < 2^24)
            ng.base = 0;
            ng.mask = 0;
        }
        while (tg.mask != 0) {
            int bit = ffs_msb_(tg.mask);
            tg.mask &= ~(1u << bit);
            test_slot(s.tris8, (int)(tg.base + (uint32_t)bit), o, d, hit);
            ++nt;
        }
        if (ng.mask <= 0x00FFFFFFu) {
            if (sp == 0) break;
            ng = stack[--sp];
        }
    }
    if (nodes_visited) *nodes_visited = nv;
    if (tris_tested) *tris_tested = nt;
    return hit;
}

// ---------------------------------------------------------------------------------------------
// radiance of a hit: query_irf's mask + UV interpolation + bilinear fetch
// (models/mat_nvdiffrast.py:299-318).  Miss, or t <= 1e-4, gives zero radiance -- NOT the next hit.
// ---------------------------------------------------------------------------------------------
struct double2_ {
    double x, y;
};
// One RGBE texel -> the float32 values cv2.imread + `* 2**hdr_exposure` give for it (models/mat_nvdiffrast.py:89-93):
// OpenCV's rgbe2float is mantissa * 2^(E - 136) (black when E == 0), the exposure another power of two, so the product
// is exact in float32 -- the packed texture reproduces the fp32 one bit for bit at a quarter of the bytes.
TX_HD void rgbe_texel(uint32_t w, int bias, float* r, float* g, float* b) {
    const uint32_t e = w >> 24;
    if (e == 0u) {
        *r = *g = *b = 0.f;
        return;
    }
    const int k = (int)e - 136 + bias;
    const float fr = (float)(w & 0xFFu), fg = (float)((w >> 8) & 0xFFu), fb = (float)((w >> 16) & 0xFFu);
    if (k >= -126 && k <= 127) {
        const float scale = int_as_float_((k + 127) << 23);
        *r = fr * scale;
        *g = fg * scale;
        *b = fb * scale;
    } else {  // denormal or overflowing scale: exact scaling by the library
        *r = ldexpf(fr, k);
        *g = ldexpf(fg, k);
        *b = ldexpf(fb, k);
    }
}

TX_HD V3 fetch_radiance(const SceneView& s, const Hit& h) {
    if (!(h.t > kHitEps && h.t < INFINITY)) return v3(0.f, 0.f, 0.f);
    double tuv[6];
#ifdef __CUDA_ARCH__
    const double2* up = reinterpret_cast<const double2*>(s.tri_uvs + 6 * (size_t)h.prim);
    double2 a = TX_LD_LEAF(up + 0), b = TX_LD_LEAF(up + 1), c = TX_LD_LEAF(up + 2);
    tuv[0] = a.x; tuv[1] = a.y; tuv[2] = b.x; tuv[3] = b.y; tuv[4] = c.x; tuv[5] = c.y;
#else
    for (int i = 0; i < 6; ++i) tuv[i] = s.tri_uvs[6 * (size_t)h.prim + i];
#endif
    float gu, gv;
    interp_uv(tuv, h.u, h.v, &gu, &gv);
    Bilerp b4 = bilerp_setup(gu, gv, s.tex_w, s.tex_h);
    if (s.tex_rgbe) {
        const uint32_t* q0 = s.tex_rgbe + (size_t)b4.y0 * (size_t)s.tex_w;
        const uint32_t* q1 = s.tex_rgbe + (size_t)b4.y1 * (size_t)s.tex_w;
        float nwr, nwg, nwb, ner, neg, neb, swr, swg, swb, ser, seg, seb;
        rgbe_texel(TX_LDG(q0 + b4.x0), s.tex_rgbe_bias, &nwr, &nwg, &nwb);
        rgbe_texel(TX_LDG(q0 + b4.x1), s.tex_rgbe_bias, &ner, &neg, &neb);
        rgbe_texel(TX_LDG(q1 + b4.x0), s.tex_rgbe_bias, &swr, &swg, &swb);
        rgbe_texel(TX_LDG(q1 + b4.x1), s.tex_rgbe_bias, &ser, &seg, &seb);
        return v3(bilerp_apply(b4, nwr, ner, swr, ser), bilerp_apply(b4, nwg, neg, swg, seg), bilerp_apply(b4, nwb, neb, swb, seb));
    }
    const float4* r0 = s.tex + (size_t)b4.y0 * (size_t)s.tex_w;
    const float4* r1 = s.tex + (size_t)b4.y1 * (size_t)s.tex_w;
    float4 nw = TX_LD_LEAF(r0 + b4.x0), ne = TX_LD_LEAF(r0 + b4.x1), sw = TX_LD_LEAF(r1 + b4.x0), se = TX_LD_LEAF(r1 + b4.x1);
    return v3(bilerp_apply(b4, nw.x, ne.x, sw.x, se.x), bilerp_apply(b4, nw.y, ne.y, sw.y, se.y),
              bilerp_apply(b4, nw.z, ne.z, sw.z, se.z));
}

TX_HD Hit trace_bvh2_fast(const SceneView& s, V3 o, V3 d);  // texir_walk.h
TX_HD Hit trace_qbvh2(const SceneView& s, V3 o, V3 d);      // texir_walk.h
template <int KIND>
TX_HD Hit trace(const SceneView& s, V3 o, V3 d) {
    if (KIND == 8) return trace_bvh8(s, o, d);
    if (KIND == 4) return trace_qbvh2(s, o, d);      // 32-byte quantised BVH2 nodes
    if (KIND == 3) return trace_bvh2_fast(s, o, d);  // BVH2 layout, cheaper node step
    return trace_bvh2(s, o, d);
}

// sample-order permutation: lanes of a warp take 32 consecutive j for a fixed residue c, i.e.
// indices i = (j << a) + c.  For the Hammersley set this is a compact patch of the unit square
// (s0 in a band of 32*2^a/N, s1 in a band of 2^-a), so the 32 rays of a warp iteration are coherent.
// Patch shape: 2^a residues, at least 32 samples per residue, a <= 4 -- measured on B200 (profiles/r02_abits_sweep.txt,
// r02_prefetch_sweep2.txt): at N = 2048 a = 2 / 3 / 4 / 5 give 3.66 / 3.89 / 3.91 / 3.69e9 ray-samples/s (round 1 used the
// square patch, a = 3); the relighting shape (3072 + 1024 samples) gains 2.6 % from a = 4, N = 256 is flat.
TX_HD int coherent_bits(int N) {
    int a = 0;
    while (a < 4 && (N >> (a + 1)) >= 32) ++a;
    return a;
}

}  // namespace texir
