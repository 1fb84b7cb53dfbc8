// texir_walk.h -- the BVH2 walks the kernels use (texir_bvh.h holds the first versions and the BVH8 walk).
//
//   trace_bvh2_fast   float nodes, t = fma(plane, 1/d, -o/d), no ulp widening (the builder's padding -- 1e-5 of the
//                     scene diagonal + 4 ulp of the largest coordinate -- already covers every rounding of the slab
//                     test), all node loads issued up front
//   trace_qbvh2       the same tree with 32-byte nodes on a global 16-bit grid: half the L1/L2 traffic per visit
//
// Measured on B200 and dropped: a ray-stream variant in which idle lanes refill
// independently (better lane use, worse L1 locality) and a speculative while-while loop.
//
// Host + device: tests/cpu_emul compiles the same code with g++.
#pragma once
#include "texir_bvh.h"

namespace texir {

struct RayState {
    V3 o, d;     // origin, direction (the triangle test needs both)
    V3 id, nood; // 1/d and -o/d for the slab test
};

TX_HD void ray_setup(RayState& r, V3 o, V3 d) {
    r.o = o;
    r.d = d;
    r.id = safe_rcp_dir(d);
    r.nood = v3(-(o.x * r.id.x), -(o.y * r.id.y), -(o.z * r.id.z));
}

#ifdef __CUDA_ARCH__
// volatile asm pins the issue order: the child references are requested together with the boxes instead of
// after the box test (the compiler sinks a plain load below the branch and serialises two round trips)
TX_D float4 ld_nc_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
TX_D void ld_nc_i2(const float4* p, int* a, int* b) {
    asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(*a), "=r"(*b) : "l"(p));
}
#else
inline float4 ld_nc_f4(const float4* p) { return *p; }
inline void ld_nc_i2(const float4* p, int* a, int* b) {
    *a = float_as_int_(p->x);
    *b = float_as_int_(p->y);
}
#endif

// one inner-node step: returns the next node (inner index, leaf reference < 0, or kSentinel)
TX_HD int bvh2_step(const SceneView& s, const RayState& r, float tbest, int node, int* stack, int& sp) {
    const float4* np = s.nodes2 + 4 * (size_t)node;
    const float4 n0 = ld_nc_f4(np + 0), n1 = ld_nc_f4(np + 1), n2 = ld_nc_f4(np + 2);
    int c0, c1;
    ld_nc_i2(np + 3, &c0, &c1);
    // child 0: x = n0.xy, y = n0.zw, z = n2.xy ; child 1: x = n1.xy, y = n1.zw, z = n2.zw
    const float a0x = tx_fma(n0.x, r.id.x, r.nood.x), b0x = tx_fma(n0.y, r.id.x, r.nood.x);
    const float a0y = tx_fma(n0.z, r.id.y, r.nood.y), b0y = tx_fma(n0.w, r.id.y, r.nood.y);
    const float a0z = tx_fma(n2.x, r.id.z, r.nood.z), b0z = tx_fma(n2.y, r.id.z, r.nood.z);
    const float a1x = tx_fma(n1.x, r.id.x, r.nood.x), b1x = tx_fma(n1.y, r.id.x, r.nood.x);
    const float a1y = tx_fma(n1.z, r.id.y, r.nood.y), b1y = tx_fma(n1.w, r.id.y, r.nood.y);
    const float a1z = tx_fma(n2.z, r.id.z, r.nood.z), b1z = tx_fma(n2.w, r.id.z, r.nood.z);
    const float t0n = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
    const float t0f = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), tbest));
    const float t1n = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
    const float t1f = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), tbest));
    const bool h0 = t0n <= t0f, h1 = t1n <= t1f;
    if (h0 && h1) {
        const bool swap = t1n < t0n;
        stack[sp++] = swap ? c0 : c1;
        return swap ? c1 : c0;
    }
    if (h0) return c0;
    if (h1) return c1;
    return stack[--sp];
}

TX_HD void bvh2_leaf(const SceneView& s, const RayState& r, int node, Hit& hit) {
    const int ref = ~node;
    const int first = ref >> 3, cnt = (ref & 7) + 1;
    for (int i = 0; i < cnt; ++i) test_slot(s.tris2, first + i, r.o, r.d, hit);
}

TX_HD void hit_reset(Hit& hit) {
    hit.t = INFINITY;
    hit.u = 0.f;
    hit.v = 0.f;
    hit.prim = kInvalidId;
}

// plain closest hit with the cheaper node step (one ray, run to completion)
TX_HD Hit trace_bvh2_fast(const SceneView& s, V3 o, V3 d) {
    RayState r;
    ray_setup(r, o, d);
    Hit hit;
    hit_reset(hit);
    int stack[kStackSize];
    int sp = 0;
    stack[sp++] = kSentinel;
    int node = 0;
    while (node != kSentinel) {
        while (node >= 0 && node != kSentinel) node = bvh2_step(s, r, hit.t, node, stack, sp);
        if (node < 0) {
            bvh2_leaf(s, r, node, hit);
            node = stack[--sp];
        }
    }
    return hit;
}

// ---------------------------------------------------------------------------------------------------
// QBVH2: 32-byte nodes on a global 16-bit grid (layout in texir_bvh.h, encoder in bvh_build.cpp).
// A plane index q becomes the float 2^23 + q with ONE byte permute (0x4B00'qqqq), and the ray parameter of
// the plane is one FMA:  t = (2^23 + q) * S + O,  S = step/d,  O = (gmin - o)/d - 2^23 * S.  The permute
// selector also picks the NEAR or the FAR plane of the slab from the sign of d, so no per-axis min/max is left:
// per child 6 permutes + 6 FMAs + 2 three-input min/max.
// ---------------------------------------------------------------------------------------------------
struct RayQ {
    V3 o, d;   // for the triangle test
    V3 S, O;   // slab decode
    uint32_t selx, sely, selz;  // permute selectors of the near plane; far = sel ^ 0x0022
};

TX_HD uint32_t byte_perm_(uint32_t x, uint32_t y, uint32_t sel) {
#ifdef __CUDA_ARCH__
    // prmt.b32 directly: __byte_perm masks a run-time selector with 0x7777 first (one more ALU instruction)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(sel));
    return r;
#else
    uint64_t both = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((both >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
#endif
}

TX_HD void rayq_setup(const SceneView& s, RayQ& r, V3 o, V3 d) {
    r.o = o;
    r.d = d;
    const V3 id = safe_rcp_dir(d);
    r.S = v3(s.qstep[0] * id.x, s.qstep[1] * id.y, s.qstep[2] * id.z);
    r.O = v3((s.qmin[0] - o.x) * id.x - 8388608.0f * r.S.x, (s.qmin[1] - o.y) * id.y - 8388608.0f * r.S.y,
             (s.qmin[2] - o.z) * id.z - 8388608.0f * r.S.z);
    // low half (0x7610) = lo plane is the near plane when the ray travels in +axis
    r.selx = id.x >= 0.f ? 0x7610u : 0x7632u;
    r.sely = id.y >= 0.f ? 0x7610u : 0x7632u;
    r.selz = id.z >= 0.f ? 0x7610u : 0x7632u;
}

#ifdef __CUDA_ARCH__
#ifndef TX_NODE_EVICT_LAST
#define TX_NODE_EVICT_LAST 0
#endif
TX_D uint4 ld_nc_u4(const uint4* p) {
    uint4 v;
#if TX_NODE_EVICT_LAST
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#else
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#endif
    return v;
}
#else
inline uint4 ld_nc_u4(const uint4* p) { return *p; }
#endif

TX_HD float qplane_(uint32_t w, uint32_t sel, float S, float O) {
    return tx_fma(int_as_float_((int)byte_perm_(w, 0x4B000000u, sel)), S, O);
}

// Measured on B200 (profiles/r02_prefetch_sweep.txt, same checksum everywhere): prefetching the POSTPONED child's node
// into L1 when it is pushed gives +2.4 % on the IrT (3.891e9 -> 3.983e9) and 1.5 ... 2 % on the Mat-stage legs -- the
// pop then finds its node in L1; also prefetching a postponed leaf's triangle (= 2: -3 %) or a found leaf's triangle on
// the way out of the node loop (TX_PREFETCH_LEAF: -5 ... -8 %, the test sits in every node step) costs more than it hides.
#ifndef TX_PREFETCH_PUSH
#define TX_PREFETCH_PUSH 1
#endif
#ifndef TX_PREFETCH_UV
#define TX_PREFETCH_UV 0
#endif
#ifndef TX_PREFETCH_POP
#define TX_PREFETCH_POP 0
#endif
constexpr int kQSentinelValue = (int)0x80000000u;   // == kQSentinel (declared further down)
#ifndef TX_PREFETCH_FRONT
#define TX_PREFETCH_FRONT 0
#endif
#ifndef TX_PREFETCH_LEAF
#define TX_PREFETCH_LEAF 0
#endif
// first triangle record (48 bytes: two 32-byte sectors) of a leaf reference into L1
TX_HD void qprefetch_leaf(const SceneView& s, int leaf_ref) {
#ifdef __CUDA_ARCH__
    const char* p = (const char*)(s.tris2 + 3 * (size_t)((~leaf_ref) >> 3));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32));
#else
    (void)s; (void)leaf_ref;
#endif
}

#ifndef TX_STACK_TOP
#define TX_STACK_TOP 0
#endif
#ifndef TX_NODE_EXIT_BELOW
#define TX_NODE_EXIT_BELOW 0
#endif
TX_HD int qbvh2_step(const SceneView& s, const RayQ& r, float tbest, int node, int*& sp, int& top) {
    const uint4* np = s.qnodes2 + 2 * (size_t)node;
    // two 128-bit loads: sm_100's 256-bit LDG (LDG.E.ENL2.256) was measured 23 % slower here -- it does not
    // allocate in L1, and the upper tree lives there
    const uint4 q0 = ld_nc_u4(np + 0), q1 = ld_nc_u4(np + 1);
    const uint32_t fx = r.selx ^ 0x0022u, fy = r.sely ^ 0x0022u, fz = r.selz ^ 0x0022u;
    const float t0n = fmaxf(fmaxf(qplane_(q0.x, r.selx, r.S.x, r.O.x), qplane_(q0.y, r.sely, r.S.y, r.O.y)),
                            fmaxf(qplane_(q0.z, r.selz, r.S.z, r.O.z), 0.0f));
    const float t0f = fminf(fminf(qplane_(q0.x, fx, r.S.x, r.O.x), qplane_(q0.y, fy, r.S.y, r.O.y)),
                            fminf(qplane_(q0.z, fz, r.S.z, r.O.z), tbest));
    const float t1n = fmaxf(fmaxf(qplane_(q0.w, r.selx, r.S.x, r.O.x), qplane_(q1.x, r.sely, r.S.y, r.O.y)),
                            fmaxf(qplane_(q1.y, r.selz, r.S.z, r.O.z), 0.0f));
    const float t1f = fminf(fminf(qplane_(q0.w, fx, r.S.x, r.O.x), qplane_(q1.x, fy, r.S.y, r.O.y)),
                            fminf(qplane_(q1.y, fz, r.S.z, r.O.z), tbest));
    const int c0 = (int)q1.z, c1 = (int)q1.w;
    const bool h0 = t0n <= t0f, h1 = t1n <= t1f;
#if TX_STACK_TOP
    // the top of the stack lives in a register: a pop hands it out at once and re-fills it with a load whose
    // latency hides behind the next node step instead of sitting between the pop and the next node fetch
    if (h0 && h1) {
        const bool swap = t1n < t0n;
        *sp++ = top;
        top = swap ? c0 : c1;
        return swap ? c1 : c0;
    }
    if (h0) return c0;
    if (h1) return c1;
    const int popped = top;
    top = *--sp;
    return popped;
#else
    if (h0 && h1) {
        const bool swap = t1n < t0n;
        const int far_child = swap ? c0 : c1;
#if TX_PREFETCH_PUSH && defined(__CUDA_ARCH__)
        // the postponed child's node is fetched into L1 now, so that the pop finds it there (-DTX_PREFETCH_PUSH=1;
        // = 2: a postponed LEAF's first triangle too)
#if TX_PREFETCH_PUSH == 3   // into L2 only
        if (far_child >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(s.qnodes2 + 2 * (size_t)far_child));
#else
        if (far_child >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(s.qnodes2 + 2 * (size_t)far_child));
#endif
#if TX_PREFETCH_PUSH == 2
        else qprefetch_leaf(s, far_child);
#endif
#endif
        *sp++ = far_child;
        return swap ? c1 : c0;
    }
    if (h0) return c0;
    if (h1) return c1;
#if TX_PREFETCH_POP && defined(__CUDA_ARCH__)
    {   // the entry that will be popped NEXT is fetched into L1 now (its push-time prefetch is long evicted)
        const int popped = *--sp;
        const int below = sp[-1];
        if (popped != kQSentinelValue && below >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(s.qnodes2 + 2 * (size_t)below));
        return popped;
    }
#else
    return *--sp;
#endif
#endif
}

// ---------------------------------------------------------------------------------------------------
// North-star variant (BASELINE.json: "flattened BVH staged into shared memory via TMA"): the first `top_count` nodes
// of the array -- the top levels of the tree, breadth first -- are copied into the CTA's shared memory with ONE bulk
// asynchronous copy (cp.async.bulk, the 1-D TMA path) that completes on an mbarrier; a node step then reads a staged
// node with two LDS.128 instead of two global loads.  Measured against the default walk in profiles/ (DESIGN.md 3).
// ---------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
TX_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TX_D void tma_stage_1d(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* mbar, bool issuer) {
    // bytes: a multiple of 16; dst / src 16-byte aligned.  Call from every thread of the CTA.
    if (issuer) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (issuer) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(dst_smem)),
                     "l"(src_global), "r"(bytes), "r"(smem_u32(mbar))
                     : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_u32(mbar))
                     : "memory");
}
#endif

// node step that reads staged nodes from shared memory (`top` = staged copy of qnodes2[0 .. top_count))
TX_HD int qbvh2_step_top(const SceneView& s, const RayQ& r, float tbest, int node, int*& sp, const uint4* top, int top_count) {
    uint4 q0, q1;
    if (node < top_count) {
        q0 = top[2 * node];
        q1 = top[2 * node + 1];
    } else {
        const uint4* np = s.qnodes2 + 2 * (size_t)node;
        q0 = ld_nc_u4(np + 0);
        q1 = ld_nc_u4(np + 1);
    }
    const uint32_t fx = r.selx ^ 0x0022u, fy = r.sely ^ 0x0022u, fz = r.selz ^ 0x0022u;
    const float t0n = fmaxf(fmaxf(qplane_(q0.x, r.selx, r.S.x, r.O.x), qplane_(q0.y, r.sely, r.S.y, r.O.y)),
                            fmaxf(qplane_(q0.z, r.selz, r.S.z, r.O.z), 0.0f));
    const float t0f = fminf(fminf(qplane_(q0.x, fx, r.S.x, r.O.x), qplane_(q0.y, fy, r.S.y, r.O.y)),
                            fminf(qplane_(q0.z, fz, r.S.z, r.O.z), tbest));
    const float t1n = fmaxf(fmaxf(qplane_(q0.w, r.selx, r.S.x, r.O.x), qplane_(q1.x, r.sely, r.S.y, r.O.y)),
                            fmaxf(qplane_(q1.y, r.selz, r.S.z, r.O.z), 0.0f));
    const float t1f = fminf(fminf(qplane_(q0.w, fx, r.S.x, r.O.x), qplane_(q1.x, fy, r.S.y, r.O.y)),
                            fminf(qplane_(q1.y, fz, r.S.z, r.O.z), tbest));
    const int c0 = (int)q1.z, c1 = (int)q1.w;
    const bool h0 = t0n <= t0f, h1 = t1n <= t1f;
    if (h0 && h1) {
        const bool swap = t1n < t0n;
        *sp++ = swap ? c0 : c1;
        return swap ? c1 : c0;
    }
    if (h0) return c0;
    if (h1) return c1;
    return *--sp;
}

#ifndef TX_SPECULATIVE
#define TX_SPECULATIVE 0
#endif

// The QBVH2 walk keeps its stack pointer as an address (no index scaling per push / pop) and ends on a NEGATIVE
// sentinel that no leaf reference can equal (leaf = ~((first_slot << 3) | (count - 1)) with first_slot < 2^27), so the
// node loop tests `node >= 0` only.
constexpr int kQSentinel = (int)0x80000000u;
TX_HD int qstack_pop(int*& sp, int& top) {
#if TX_STACK_TOP
    const int n = top;
    top = *--sp;
    return n;
#else
    (void)top;
    return *--sp;
#endif
}
TX_HD void qbvh2_leaf(const SceneView& s, const RayQ& r, int node, Hit& hit) {
    const int ref = ~node;
    const int first = ref >> 3, cnt = (ref & 7) + 1;
#if TX_PREFETCH_UV && defined(__CUDA_ARCH__)
    const uint32_t before = hit.prim;
#endif
    for (int i = 0; i < cnt; ++i) test_slot(s.tris2, first + i, r.o, r.d, hit);
#if TX_PREFETCH_UV && defined(__CUDA_ARCH__)
    // a new closest hit: its UV record (48 bytes of float64, the first thing the radiance fetch reads) into L2 / L1 now
    if (hit.prim != before && s.tri_uvs) {
        const char* p = (const char*)(s.tri_uvs + 6 * (size_t)hit.prim);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32));
    }
#endif
}

TX_HD Hit trace_qbvh2(const SceneView& s, V3 o, V3 d) {
    RayQ r;
    rayq_setup(s, r, o, d);
    Hit hit;
    hit_reset(hit);
    int stack[kStackSize];
    int* sp = stack;
    *sp++ = kQSentinel;
    int top = kQSentinel;  // TX_STACK_TOP: register copy of the top entry (stack[0] then backs the final pop)
    int node = 0;
#if TX_SPECULATIVE
    // Speculative traversal (Aila & Laine 2009): the first leaf a lane finds is parked and the lane keeps walking with
    // its old t_best; the node loop ends once every lane still in it has parked a leaf.  Modelled in
    // profiles/simulate_warp_divergence.py ("1 postponed leaf, stop when all found").
    int leaf = 0;  // 0 = none (leaf references are negative)
    for (;;) {
        while (node >= 0) {
            node = qbvh2_step(s, r, hit.t, node, sp, top);
            if (node < 0 && node != kQSentinel && leaf == 0) {
                leaf = node;
                node = qstack_pop(sp, top);
            }
#ifdef __CUDA_ARCH__
            if (!__any_sync(__activemask(), leaf == 0)) break;
#endif
        }
        while (leaf < 0) {
            qbvh2_leaf(s, r, leaf, hit);
            leaf = 0;
            if (node < 0 && node != kQSentinel) {
                leaf = node;
                node = qstack_pop(sp, top);
            }
        }
        if (node == kQSentinel) break;
    }
#else
    while (node != kQSentinel) {
#if defined(__CUDA_ARCH__) && TX_NODE_EXIT_BELOW
        // leave the node loop once fewer than TX_NODE_EXIT_BELOW lanes are still in it, so that the lanes waiting at
        // a leaf do not wait for the last stragglers (they re-enter with the next round)
        while (node >= 0) {
            node = qbvh2_step(s, r, hit.t, node, sp, top);
            if (__popc(__activemask()) < TX_NODE_EXIT_BELOW) break;
        }
#else
        while (node >= 0) {
            node = qbvh2_step(s, r, hit.t, node, sp, top);
#if TX_PREFETCH_LEAF
            if (node < 0 && node != kQSentinel) qprefetch_leaf(s, node);   // fetched while the other lanes still walk
#endif
        }
#endif
        if (node < 0 && node != kQSentinel) {
            qbvh2_leaf(s, r, node, hit);
            node = qstack_pop(sp, top);
        }
    }
#endif
    return hit;
}

// Free-running walk: every lane takes a node step or a leaf test per iteration, whichever its ray stands at, and no
// lane waits for the others to reach a leaf.  A round costs both bodies whenever the warp holds both kinds of lanes, so
// on a kernel that is bound by instruction issue (the IrT, the render) the lock-step while-while loop above is faster;
// on one that is bound by the LATENCY of its slowest warp (the G-buffer: 0.65 waves, every warp resident from the start)
// the number of rounds is what counts -- max over the lanes of their own steps, instead of the sum over rounds of the
// slowest lane (profiles/simulate_gbuffer_warps.py).  Same visiting order per ray, same hits.
TX_HD Hit trace_qbvh2_free(const SceneView& s, V3 o, V3 d) {
    RayQ r;
    rayq_setup(s, r, o, d);
    Hit hit;
    hit_reset(hit);
    int stack[kStackSize];
    int* sp = stack;
    *sp++ = kQSentinel;
    int top = kQSentinel;
    int node = 0;
    while (node != kQSentinel) {
        if (node >= 0) {
            node = qbvh2_step(s, r, hit.t, node, sp, top);
        } else {
            qbvh2_leaf(s, r, node, hit);
            node = qstack_pop(sp, top);
        }
    }
    return hit;
}

// Stack entries that also carry the entry distance of the postponed child: a pop drops every entry whose distance
// already exceeds the closest hit -- otherwise such an entry costs a whole node step that finds neither child (model,
// profiles/simulate_stack_tnear.py: 1.9 of a ray's 40 node steps).  Same hits; 8-byte entries, twice the stack lines in
// L1.  Measured on B200 (profiles/r02_tnear_sweep.txt): the 40-register IrT kernel LOSES 1 % (12 CTAs per SM: the stacks
// take the L1 the nodes need), the 48-register render kernels gain 0.7 ... 3.5 % -- so the render kernels use it
// (TX_SHADE_TNEAR, default 1) and the IrT kernel does not (-DTX_STACK_TNEAR=1 forces it everywhere).
#ifndef TX_STACK_TNEAR
#define TX_STACK_TNEAR 0
#endif
#ifndef TX_SHADE_TNEAR
#define TX_SHADE_TNEAR 1
#endif
struct alignas(8) QEnt {
    int ref;
    float t;
};
TX_HD int qpop_t(QEnt*& sp, float tbest) {
    QEnt e;
    do {
        e = *--sp;
    } while (e.t > tbest);  // the sentinel carries t = 0
    return e.ref;
}
TX_HD int qbvh2_step_t(const SceneView& s, const RayQ& r, float tbest, int node, QEnt*& sp) {
    const uint4* np = s.qnodes2 + 2 * (size_t)node;
    const uint4 q0 = ld_nc_u4(np + 0), q1 = ld_nc_u4(np + 1);
    const uint32_t fx = r.selx ^ 0x0022u, fy = r.sely ^ 0x0022u, fz = r.selz ^ 0x0022u;
    const float t0n = fmaxf(fmaxf(qplane_(q0.x, r.selx, r.S.x, r.O.x), qplane_(q0.y, r.sely, r.S.y, r.O.y)),
                            fmaxf(qplane_(q0.z, r.selz, r.S.z, r.O.z), 0.0f));
    const float t0f = fminf(fminf(qplane_(q0.x, fx, r.S.x, r.O.x), qplane_(q0.y, fy, r.S.y, r.O.y)),
                            fminf(qplane_(q0.z, fz, r.S.z, r.O.z), tbest));
    const float t1n = fmaxf(fmaxf(qplane_(q0.w, r.selx, r.S.x, r.O.x), qplane_(q1.x, r.sely, r.S.y, r.O.y)),
                            fmaxf(qplane_(q1.y, r.selz, r.S.z, r.O.z), 0.0f));
    const float t1f = fminf(fminf(qplane_(q0.w, fx, r.S.x, r.O.x), qplane_(q1.x, fy, r.S.y, r.O.y)),
                            fminf(qplane_(q1.y, fz, r.S.z, r.O.z), tbest));
    const int c0 = (int)q1.z, c1 = (int)q1.w;
    const bool h0 = t0n <= t0f, h1 = t1n <= t1f;
    if (h0 && h1) {
        const bool swap = t1n < t0n;
        QEnt e;
        e.ref = swap ? c0 : c1;
        e.t = swap ? t0n : t1n;
#if TX_PREFETCH_PUSH && defined(__CUDA_ARCH__)
        if (e.ref >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(s.qnodes2 + 2 * (size_t)e.ref));
#endif
        *sp++ = e;
        return swap ? c1 : c0;
    }
    if (h0) return c0;
    if (h1) return c1;
    return qpop_t(sp, tbest);
}
TX_HD Hit trace_qbvh2_t(const SceneView& s, V3 o, V3 d, const uint4* entries, int count, int start) {
    RayQ r;
    rayq_setup(s, r, o, d);
    Hit hit;
    hit_reset(hit);
    QEnt stack[kStackSize];
    QEnt* sp = stack;
    {
        QEnt e;
        e.ref = kQSentinel;
        e.t = 0.f;
        *sp++ = e;
    }
    const uint32_t fx = r.selx ^ 0x0022u, fy = r.sely ^ 0x0022u, fz = r.selz ^ 0x0022u;
    for (int k = 0; k < count; ++k) {
        const uint4 en = entries[k];
        const float tn = fmaxf(fmaxf(qplane_(en.x, r.selx, r.S.x, r.O.x), qplane_(en.y, r.sely, r.S.y, r.O.y)),
                               fmaxf(qplane_(en.z, r.selz, r.S.z, r.O.z), 0.0f));
        const float tf = fminf(fminf(qplane_(en.x, fx, r.S.x, r.O.x), qplane_(en.y, fy, r.S.y, r.O.y)),
                               qplane_(en.z, fz, r.S.z, r.O.z));
        if (tn <= tf) {
            QEnt e;
            e.ref = (int)en.w;
            e.t = tn;
            *sp++ = e;
        }
    }
    int node = start;
    while (node != kQSentinel) {
        while (node >= 0) node = qbvh2_step_t(s, r, hit.t, node, sp);
        if (node < 0 && node != kQSentinel) {
            qbvh2_leaf(s, r, node, hit);
            node = qpop_t(sp, hit.t);
        }
    }
    return hit;
}

// ---------------------------------------------------------------------------------------------------
// Shared first descent.  All rays of a surface point start at the same origin, and a near-first walk always
// descends first into the child whose box CONTAINS the origin (its entry distance is 0): the chain of nodes from the
// root to the leaf around the origin is the same for every ray of the point, only the siblings along it are
// direction dependent.  qbvh2_point_front() walks that chain once per point and records the siblings (their three
// plane words + reference, 16 bytes each); trace_qbvh2_front() then starts a ray by testing just those boxes -- one
// box instead of two per level, read from shared memory instead of a dependent global load -- pushes the hit ones
// (shallow first, so the deepest lies on top exactly as the plain walk leaves its stack) and continues with the
// ordinary loop from the node where the chain ended.  Same visiting order, same hits, bit for bit.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxFront = 48;

TX_HD bool qbox_contains_(const SceneView& s, uint32_t wx, uint32_t wy, uint32_t wz, V3 o) {
    const float lx = tx_fma((float)(wx & 0xFFFFu), s.qstep[0], s.qmin[0]), hx = tx_fma((float)(wx >> 16), s.qstep[0], s.qmin[0]);
    const float ly = tx_fma((float)(wy & 0xFFFFu), s.qstep[1], s.qmin[1]), hy = tx_fma((float)(wy >> 16), s.qstep[1], s.qmin[1]);
    const float lz = tx_fma((float)(wz & 0xFFFFu), s.qstep[2], s.qmin[2]), hz = tx_fma((float)(wz >> 16), s.qstep[2], s.qmin[2]);
    return o.x >= lx && o.x <= hx && o.y >= ly && o.y <= hy && o.z >= lz && o.z <= hz;
}

// one thread per point: entries[0 .. *count) = siblings along the containing chain, root side first;
// *start = the node (inner index or leaf reference) where the chain ends and the ordinary walk takes over
TX_HD void qbvh2_point_front(const SceneView& s, V3 o, uint4* entries, int* count, int* start) {
    int node = 0, n = 0;
    while (node >= 0 && n < kMaxFront) {
        const uint4* np = s.qnodes2 + 2 * (size_t)node;
        const uint4 q0 = ld_nc_u4(np + 0), q1 = ld_nc_u4(np + 1);
        const bool in0 = qbox_contains_(s, q0.x, q0.y, q0.z, o), in1 = qbox_contains_(s, q0.w, q1.x, q1.y, o);
        uint4 e;
        if (in0) {  // also when the origin lies in both boxes: the plain walk breaks that tie towards child 0
            e.x = q0.w; e.y = q1.x; e.z = q1.y; e.w = q1.w;
            node = (int)q1.z;
        } else if (in1) {
            e.x = q0.x; e.y = q0.y; e.z = q0.z; e.w = q1.z;
            node = (int)q1.w;
        } else {
            break;  // the origin is in neither child: the ordinary walk starts here
        }
        entries[n++] = e;
    }
    *count = n;
    *start = node;
}

TX_HD Hit trace_qbvh2_front(const SceneView& s, V3 o, V3 d, const uint4* entries, int count, int start) {
#if TX_STACK_TNEAR
    return trace_qbvh2_t(s, o, d, entries, count, start);
#endif
    RayQ r;
    rayq_setup(s, r, o, d);
    Hit hit;
    hit_reset(hit);
    int stack[kStackSize];
    int* sp = stack;
    *sp++ = kQSentinel;
    int top = kQSentinel;
    {
        const uint32_t fx = r.selx ^ 0x0022u, fy = r.sely ^ 0x0022u, fz = r.selz ^ 0x0022u;
        for (int k = 0; k < count; ++k) {
            const uint4 e = entries[k];  // one 128-bit load (shared memory in the kernels: a broadcast)
            const uint32_t wx = e.x, wy = e.y, wz = e.z;
            const int ref = (int)e.w;
            const float tn = fmaxf(fmaxf(qplane_(wx, r.selx, r.S.x, r.O.x), qplane_(wy, r.sely, r.S.y, r.O.y)),
                                   fmaxf(qplane_(wz, r.selz, r.S.z, r.O.z), 0.0f));
            const float tf = fminf(fminf(qplane_(wx, fx, r.S.x, r.O.x), qplane_(wy, fy, r.S.y, r.O.y)),
                                   qplane_(wz, fz, r.S.z, r.O.z));
            if (tn <= tf) {
#if TX_PREFETCH_FRONT && defined(__CUDA_ARCH__)
                if (ref >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(s.qnodes2 + 2 * (size_t)ref));
#endif
#if TX_STACK_TOP
                *sp++ = top;
                top = ref;
#else
                *sp++ = ref;
#endif
            }
        }
    }
    int node = start;
    while (node != kQSentinel) {
        while (node >= 0) {
            node = qbvh2_step(s, r, hit.t, node, sp, top);
#if TX_PREFETCH_LEAF
            if (node < 0 && node != kQSentinel) qprefetch_leaf(s, node);   // fetched while the other lanes still walk
#endif
        }
        if (node < 0 && node != kQSentinel) {
            qbvh2_leaf(s, r, node, hit);
            node = qstack_pop(sp, top);
        }
    }
    return hit;
}

// trace_qbvh2_front with the staged top levels (same visiting order, same hits)
TX_HD Hit trace_qbvh2_front_top(const SceneView& s, V3 o, V3 d, const uint4* entries, int count, int start, const uint4* top,
                                int top_count) {
    RayQ r;
    rayq_setup(s, r, o, d);
    Hit hit;
    hit_reset(hit);
    int stack[kStackSize];
    int* sp = stack;
    *sp++ = kQSentinel;
    {
        const uint32_t fx = r.selx ^ 0x0022u, fy = r.sely ^ 0x0022u, fz = r.selz ^ 0x0022u;
        for (int k = 0; k < count; ++k) {
            const uint4 e = entries[k];
            const float tn = fmaxf(fmaxf(qplane_(e.x, r.selx, r.S.x, r.O.x), qplane_(e.y, r.sely, r.S.y, r.O.y)),
                                   fmaxf(qplane_(e.z, r.selz, r.S.z, r.O.z), 0.0f));
            const float tf = fminf(fminf(qplane_(e.x, fx, r.S.x, r.O.x), qplane_(e.y, fy, r.S.y, r.O.y)),
                                   qplane_(e.z, fz, r.S.z, r.O.z));
            if (tn <= tf) *sp++ = (int)e.w;
        }
    }
    int node = start;
    while (node != kQSentinel) {
        while (node >= 0) node = qbvh2_step_top(s, r, hit.t, node, sp, top, top_count);
        if (node < 0 && node != kQSentinel) {
            qbvh2_leaf(s, r, node, hit);
            node = *--sp;
        }
    }
    return hit;
}

}  // namespace texir
