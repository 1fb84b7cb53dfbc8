// bvh_build_gpu.cu -- the scene's BVH built ON THE DEVICE (replaces the Embree build that Open3D runs lazily inside
// RaycastingScene.add_triangles / the first cast_rays, models/mat_nvdiffrast.py:98-101, models/tracer_o3d_irt.py:85-89).
//
// Same algorithm as the host builder (bvh_build.cpp): top-down binned SAH, 16 bins on each of the three axes over the
// centroid bounds, <= 3 triangles per leaf, the same cost constants, the same tie rules, areas and costs in float64
// with explicitly rounded operations -- so the device builds THE SAME TREE (node boxes, split decisions and leaf
// contents are order-independent functions of a node's primitive set; tests compare the two builders ray by ray).
//
// Organisation: level-synchronous.  The nodes of one depth are processed by one launch pair -- a CTA of 1024 threads
// per large node (> 2048 primitives), a warp per small node -- each team bins its node's primitives in shared memory,
// one lane sweeps the SAH, the team partitions the primitive indices stably into the other of two index buffers
// (ping-pong by depth) and appends the children to the next level's queues.  Children inherit their boxes from the
// parent's bins, so every primitive is read twice per level and nothing is sorted.  Afterwards the inner nodes are
// numbered in depth-first order (subtree sizes bottom-up, offsets top-down, one small launch per level each), and one
// launch each writes the float BVH2 nodes + the quantised 32-byte QBVH2 nodes, and the 48-byte triangle records in
// leaf order.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cfloat>
#include <cstring>
#include <string>
#include <vector>

#include "bvh_build.h"
#include "texir_host.h"

namespace texir {
namespace {

constexpr int kBins = 16;
constexpr int kMaxLeaf = 3;
constexpr int kSahDepthLimit = 40;
constexpr double kCostNode = 1.2, kCostTri = 1.0;
constexpr int kBigTeam = 1024;
constexpr int kBigCount = 2048;   // nodes with more primitives get a whole CTA
constexpr int kMaxLevels = 60;    // deeper than this: give up (the host builder takes over)

struct DNode {
    float nb[6];   // node box: lo xyz, hi xyz
    float cb[6];   // centroid box
    int first, count;
    int left, right;  // node ids of the children, -1 = leaf
    int depth;
    int flags;        // bit 0: boxes unknown, compute them from the primitives first
    int index;        // depth-first number among the inner nodes
    int inner;        // inner nodes in this subtree
};

struct Ctx {
    const float4* pbox;   // 2 float4 per primitive: (lo, 0), (hi, 0)
    DNode* nodes;
    int* node_count;
    int* idx[2];          // primitive indices, ping-pong by depth parity
    int* final_idx;       // primitive of every leaf slot
    int* q_big[2];        // queues of node ids, by parity of the depth being processed
    int* q_small[2];
    int* q_count;         // [parity][0] big, [parity][1] small
};

// order-preserving float <-> uint32 keys, so that atomicMin / atomicMax work on floats of either sign
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

struct Bins {                 // one team's shared memory
    uint32_t lo[3][kBins][3], hi[3][kBins][3];     // primitive boxes per (axis, bin)
    uint32_t clo[3][kBins][3], chi[3][kBins][3];   // centroid boxes per (axis, bin)
    int cnt[3][kBins];
    uint32_t blo[3], bhi[3], bclo[3], bchi[3];     // whole-node bounds (flags & 1)
    int best_axis, best_split, is_leaf, nl;
    int warp_sums[32];
};

template <int TEAM>
__device__ __forceinline__ void team_sync() {
    if (TEAM == 32)
        __syncwarp();
    else
        __syncthreads();
}

// exclusive prefix count of `f` over the team, and the team's total
template <int TEAM>
__device__ __forceinline__ int team_scan(bool f, int lane, int* total, int* warp_sums) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, f);
    const int wl = lane & 31;
    const int r = __popc(m & ((1u << wl) - 1u)), wt = __popc(m);
    if (TEAM == 32) {
        *total = wt;
        return r;
    }
    const int w = lane >> 5;
    if (wl == 0) warp_sums[w] = wt;
    __syncthreads();
    int pre = 0, tot = 0;
#pragma unroll 8
    for (int i = 0; i < TEAM / 32; ++i) {
        const int v = warp_sums[i];
        if (i < w) pre += v;
        tot += v;
    }
    __syncthreads();
    *total = tot;
    return pre + r;
}

__device__ __forceinline__ void smin(uint32_t* p, uint32_t v) {
    if (v < *(volatile uint32_t*)p) atomicMin(p, v);
}
__device__ __forceinline__ void smax(uint32_t* p, uint32_t v) {
    if (v > *(volatile uint32_t*)p) atomicMax(p, v);
}

// half surface area in float64, operation by operation like Box::area() of the host builder (-ffp-contract=off)
__device__ __forceinline__ double area_of(const float* lo, const float* hi) {
    if (lo[0] > hi[0]) return 0.0;
    const double dx = __dsub_rn((double)hi[0], (double)lo[0]), dy = __dsub_rn((double)hi[1], (double)lo[1]),
                 dz = __dsub_rn((double)hi[2], (double)lo[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dy), __dmul_rn(dy, dz)), __dmul_rn(dz, dx));
}

__device__ __forceinline__ void prim_box(const Ctx& c, int p, float* lo, float* hi, float* cen) {
    const float4 a = __ldg(c.pbox + 2 * (size_t)p), b = __ldg(c.pbox + 2 * (size_t)p + 1);
    lo[0] = a.x; lo[1] = a.y; lo[2] = a.z;
    hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
    for (int k = 0; k < 3; ++k) cen[k] = __fmul_rn(0.5f, __fadd_rn(lo[k], hi[k]));
}

__device__ __forceinline__ int bin_of(float c, float lo, float scale) {
    const int b = (int)__fmul_rn(__fsub_rn(c, lo), scale);
    return b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
}

template <int TEAM>
__device__ void process_node(const Ctx& c, int nid, Bins& sb, int lane) {
    DNode* n = c.nodes + nid;
    const int first = n->first, count = n->count, depth = n->depth;
    const int* cur = c.idx[depth & 1];
    int* nxt = c.idx[(depth + 1) & 1];

    // ---- 0. boxes of the node, when the parent could not hand them down (the root; children of a median split) ----
    if (n->flags & 1) {
        if (lane < 3) {
            sb.blo[lane] = sb.bclo[lane] = 0xFFFFFFFFu;
            sb.bhi[lane] = sb.bchi[lane] = 0u;
        }
        team_sync<TEAM>();
        for (int k = lane; k < count; k += TEAM) {
            float lo[3], hi[3], cen[3];
            prim_box(c, cur[first + k], lo, hi, cen);
            for (int a = 0; a < 3; ++a) {
                smin(&sb.blo[a], fkey(lo[a]));
                smax(&sb.bhi[a], fkey(hi[a]));
                smin(&sb.bclo[a], fkey(cen[a]));
                smax(&sb.bchi[a], fkey(cen[a]));
            }
        }
        team_sync<TEAM>();
        if (lane < 3) {
            n->nb[lane] = fkey_inv(sb.blo[lane]);
            n->nb[3 + lane] = fkey_inv(sb.bhi[lane]);
            n->cb[lane] = fkey_inv(sb.bclo[lane]);
            n->cb[3 + lane] = fkey_inv(sb.bchi[lane]);
            if (lane == 0) n->flags &= ~1;
        }
        team_sync<TEAM>();
    }
    float nb[6], cb[6];
    for (int k = 0; k < 6; ++k) {
        nb[k] = *(volatile float*)&n->nb[k];
        cb[k] = *(volatile float*)&n->cb[k];
    }
    if (count == 1) {  // only a one-triangle root gets here
        if (lane == 0) {
            c.final_idx[first] = cur[first];
            n->left = n->right = -1;
        }
        return;
    }
    const float k16 = 16.0f * 0.999999f;  // (float)kBins * 0.999999f of the host builder
    float scale[3];
    for (int a = 0; a < 3; ++a) {
        const float ext = __fsub_rn(cb[3 + a], cb[a]);
        scale[a] = ext > 0.f ? __fdiv_rn(k16, ext) : 0.f;
    }
    const bool sah = depth <= kSahDepthLimit;

    // ---- 1. binning ---------------------------------------------------------------------------------
    for (int i = lane; i < 3 * kBins; i += TEAM) {
        const int a = i / kBins, b = i % kBins;
        sb.cnt[a][b] = 0;
        for (int k = 0; k < 3; ++k) {
            sb.lo[a][b][k] = sb.clo[a][b][k] = 0xFFFFFFFFu;
            sb.hi[a][b][k] = sb.chi[a][b][k] = 0u;
        }
    }
    team_sync<TEAM>();
    if (sah) {
        for (int k = lane; k < count; k += TEAM) {
            float lo[3], hi[3], cen[3];
            prim_box(c, cur[first + k], lo, hi, cen);
            uint32_t klo[3], khi[3], kc[3];
            for (int j = 0; j < 3; ++j) {
                klo[j] = fkey(lo[j]);
                khi[j] = fkey(hi[j]);
                kc[j] = fkey(cen[j]);
            }
            for (int a = 0; a < 3; ++a) {
                if (scale[a] == 0.f) continue;
                const int b = bin_of(cen[a], cb[a], scale[a]);
                atomicAdd(&sb.cnt[a][b], 1);
                for (int j = 0; j < 3; ++j) {
                    smin(&sb.lo[a][b][j], klo[j]);
                    smax(&sb.hi[a][b][j], khi[j]);
                    smin(&sb.clo[a][b][j], kc[j]);
                    smax(&sb.chi[a][b][j], kc[j]);
                }
            }
        }
    }
    team_sync<TEAM>();

    // ---- 2. SAH sweep (one lane, in the host builder's order: axis 0..2, split 15..1, strict <) --------
    if (lane == 0) {
        int best_axis = -1, best_split = -1;
        double best_cost = DBL_MAX;
        if (sah) {
            for (int a = 0; a < 3; ++a) {
                if (scale[a] == 0.f) continue;
                double larea[kBins];
                int lcount[kBins];
                float alo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ahi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                int cnt = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    if (sb.cnt[a][b] > 0)
                        for (int j = 0; j < 3; ++j) {
                            alo[j] = fminf(alo[j], fkey_inv(sb.lo[a][b][j]));
                            ahi[j] = fmaxf(ahi[j], fkey_inv(sb.hi[a][b][j]));
                        }
                    cnt += sb.cnt[a][b];
                    larea[b] = area_of(alo, ahi);
                    lcount[b] = cnt;
                }
                for (int j = 0; j < 3; ++j) {
                    alo[j] = FLT_MAX;
                    ahi[j] = -FLT_MAX;
                }
                cnt = 0;
                for (int b = kBins - 1; b >= 1; --b) {
                    if (sb.cnt[a][b] > 0)
                        for (int j = 0; j < 3; ++j) {
                            alo[j] = fminf(alo[j], fkey_inv(sb.lo[a][b][j]));
                            ahi[j] = fmaxf(ahi[j], fkey_inv(sb.hi[a][b][j]));
                        }
                    cnt += sb.cnt[a][b];
                    const int lc = lcount[b - 1];
                    if (lc == 0 || cnt == 0) continue;
                    const double cost = __dadd_rn(__dmul_rn(larea[b - 1], (double)lc), __dmul_rn(area_of(alo, ahi), (double)cnt));
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = a;
                        best_split = b;
                    }
                }
            }
        }
        int is_leaf = 0;
        if (count <= kMaxLeaf) {
            const double na = area_of(nb, nb + 3);
            const double leaf_cost = __dmul_rn(__dmul_rn(kCostTri, (double)count), na);
            const double split_cost = best_axis < 0 ? DBL_MAX : __dadd_rn(__dmul_rn(kCostNode, na), __dmul_rn(kCostTri, best_cost));
            if (leaf_cost <= split_cost) is_leaf = 1;
        }
        int nl = count / 2;  // median split by position when the SAH found nothing (identical centroids, depth limit)
        if (best_axis >= 0) {
            nl = 0;
            for (int b = 0; b < best_split; ++b) nl += sb.cnt[best_axis][b];
        }
        sb.best_axis = best_axis;
        sb.best_split = best_split;
        sb.is_leaf = is_leaf;
        sb.nl = nl;
    }
    team_sync<TEAM>();
    const int best_axis = sb.best_axis, best_split = sb.best_split, nl = sb.nl;
    if (sb.is_leaf) {
        for (int k = lane; k < count; k += TEAM) c.final_idx[first + k] = cur[first + k];
        if (lane == 0) n->left = n->right = -1;
        return;
    }

    // ---- 3. stable partition into the other index buffer -----------------------------------------------
    int base_l = 0, base_r = 0;
    for (int k0 = 0; k0 < count; k0 += TEAM) {
        const int k = k0 + lane;
        const bool valid = k < count;
        int p = 0;
        bool left = false;
        if (valid) {
            p = cur[first + k];
            if (best_axis >= 0) {
                float lo[3], hi[3], cen[3];
                prim_box(c, p, lo, hi, cen);
                left = bin_of(cen[best_axis], cb[best_axis], scale[best_axis]) < best_split;
            } else {
                left = k < nl;
            }
        }
        int nleft;
        const int rank = team_scan<TEAM>(left, lane, &nleft, sb.warp_sums);
        const int nvalid = count - k0 < TEAM ? count - k0 : TEAM;
        if (valid) nxt[left ? first + base_l + rank : first + nl + base_r + (lane - rank)] = p;
        base_l += nleft;
        base_r += nvalid - nleft;
    }
    team_sync<TEAM>();

    // ---- 4. children -----------------------------------------------------------------------------------
    if (lane == 0) {
        const int l = atomicAdd(c.node_count, 2), r = l + 1;
        n->left = l;
        n->right = r;
        for (int side = 0; side < 2; ++side) {
            DNode* ch = c.nodes + (side ? r : l);
            ch->first = side ? first + nl : first;
            ch->count = side ? count - nl : nl;
            ch->depth = depth + 1;
            ch->left = ch->right = -1;
            ch->index = -1;
            ch->inner = 0;
            ch->flags = 0;
            if (best_axis >= 0) {
                float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                const int b0 = side ? best_split : 0, b1 = side ? kBins : best_split;
                for (int b = b0; b < b1; ++b) {
                    if (sb.cnt[best_axis][b] == 0) continue;
                    for (int j = 0; j < 3; ++j) {
                        lo[j] = fminf(lo[j], fkey_inv(sb.lo[best_axis][b][j]));
                        hi[j] = fmaxf(hi[j], fkey_inv(sb.hi[best_axis][b][j]));
                        clo[j] = fminf(clo[j], fkey_inv(sb.clo[best_axis][b][j]));
                        chi[j] = fmaxf(chi[j], fkey_inv(sb.chi[best_axis][b][j]));
                    }
                }
                for (int j = 0; j < 3; ++j) {
                    ch->nb[j] = lo[j];
                    ch->nb[3 + j] = hi[j];
                    ch->cb[j] = clo[j];
                    ch->cb[3 + j] = chi[j];
                }
            } else {
                ch->flags = 1;
            }
            if (ch->count == 1) {  // a leaf at once (the host builder's `if (count == 1) return false`)
                const int p = nxt[ch->first];
                c.final_idx[ch->first] = p;
                if (ch->flags & 1) {
                    float lo[3], hi[3], cen[3];
                    prim_box(c, p, lo, hi, cen);
                    for (int j = 0; j < 3; ++j) {
                        ch->nb[j] = lo[j];
                        ch->nb[3 + j] = hi[j];
                        ch->cb[j] = ch->cb[3 + j] = cen[j];
                    }
                    ch->flags = 0;
                }
            } else {
                const int par = (depth + 1) & 1;
                if (ch->count > kBigCount)
                    c.q_big[par][atomicAdd(c.q_count + 2 * par + 0, 1)] = side ? r : l;
                else
                    c.q_small[par][atomicAdd(c.q_count + 2 * par + 1, 1)] = side ? r : l;
            }
        }
    }
}

__global__ void __launch_bounds__(kBigTeam) k_build_big(Ctx c, int parity, int count) {
    __shared__ Bins sb;
    if ((int)blockIdx.x >= count) return;
    process_node<kBigTeam>(c, c.q_big[parity][blockIdx.x], sb, threadIdx.x);
}

__global__ void __launch_bounds__(128) k_build_small(Ctx c, int parity, int count) {
    __shared__ Bins sb[4];
    const int w = threadIdx.x >> 5, team = blockIdx.x * 4 + w;
    if (team >= count) return;
    process_node<32>(c, c.q_small[parity][team], sb[w], threadIdx.x & 31);
}

// primitive boxes + input validation + scene bounds
__global__ void __launch_bounds__(256) k_prim_boxes(const float* __restrict__ verts, int64_t nv, const int32_t* __restrict__ tris,
                                                    int64_t nt, float4* __restrict__ pbox, int* __restrict__ idx0,
                                                    uint32_t* __restrict__ scene_keys, int* __restrict__ error) {
    __shared__ uint32_t s[6];
    if (threadIdx.x < 3) s[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) s[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nt) {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        bool ok = true;
        for (int k = 0; k < 3; ++k) {
            const int32_t vi = tris[3 * p + k];
            if (vi < 0 || vi >= nv) {
                ok = false;
                break;
            }
            for (int a = 0; a < 3; ++a) {
                const float x = verts[3 * (size_t)vi + a];
                if (!isfinite(x)) ok = false;
                lo[a] = fminf(lo[a], x);
                hi[a] = fmaxf(hi[a], x);
            }
        }
        if (!ok) {
            atomicExch(error, 1);
            for (int a = 0; a < 3; ++a) lo[a] = hi[a] = 0.f;
        }
        pbox[2 * p] = make_float4(lo[0], lo[1], lo[2], 0.f);
        pbox[2 * p + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        idx0[p] = (int)p;
        for (int a = 0; a < 3; ++a) {
            smin(&s[a], fkey(lo[a]));
            smax(&s[3 + a], fkey(hi[a]));
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(scene_keys + threadIdx.x, s[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(scene_keys + threadIdx.x, s[threadIdx.x]);
}

__global__ void k_init_root(Ctx c, int64_t nt) {
    DNode* n = c.nodes;
    n->first = 0;
    n->count = (int)nt;
    n->left = n->right = -1;
    n->depth = 0;
    n->flags = 1;
    n->index = 0;
    n->inner = 0;
    *c.node_count = 1;
    c.q_count[0] = c.q_count[1] = c.q_count[2] = c.q_count[3] = 0;
    if (nt > kBigCount) {
        c.q_big[0][0] = 0;
        c.q_count[0] = 1;
    } else {
        c.q_small[0][0] = 0;
        c.q_count[1] = 1;
    }
}

// Numbering of the inner nodes: the top kTopDepth levels breadth first (a contiguous prefix of the node arrays that a
// CTA can stage in shared memory with one bulk copy), every subtree below them depth first.  One thread walks the
// (at most 2^kTopDepth) top nodes; the subtrees follow from their inner-node counts (deepest level first), top down.
constexpr int kTopDepth = 10;
__global__ void k_number_top(DNode* nodes, const int* __restrict__ level_begin, int levels, int* __restrict__ top_count) {
    int next = 0;
    const int dmax = levels < kTopDepth ? levels : kTopDepth;
    for (int d = 0; d < dmax; ++d)
        for (int i = level_begin[d]; i < level_begin[d + 1]; ++i)
            if (nodes[i].left >= 0) nodes[i].index = next++;
    *top_count = next;
    if (dmax < levels)
        for (int i = level_begin[dmax]; i < level_begin[dmax + 1]; ++i)
            if (nodes[i].left >= 0) {
                nodes[i].index = next;
                next += nodes[i].inner;
            }
}

__global__ void __launch_bounds__(256) k_count_inner(DNode* nodes, int begin, int end) {
    const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    DNode& n = nodes[i];
    n.inner = n.left < 0 ? 0 : 1 + nodes[n.left].inner + nodes[n.right].inner;
}
__global__ void __launch_bounds__(256) k_number_inner(DNode* nodes, int begin, int end) {
    const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const DNode& n = nodes[i];
    if (n.left < 0) return;
    nodes[n.left].index = n.index + 1;
    nodes[n.right].index = n.index + 1 + nodes[n.left].inner;
}

struct EmitArgs {
    float pad;
    float qmin[3], qstep[3];
    float4* nodes2;
    uint4* qnodes2;
    double* sah;
    int* max_depth;
};

__global__ void __launch_bounds__(256) k_emit_nodes(const DNode* __restrict__ nodes, int total, EmitArgs e) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const DNode& n = nodes[i];
    if (n.left < 0) return;
    float lo[2][3], hi[2][3];
    int ref[2];
    double sah = 0.0;
    for (int s = 0; s < 2; ++s) {
        const DNode& ch = nodes[s ? n.right : n.left];
        for (int a = 0; a < 3; ++a) {
            lo[s][a] = __fsub_rn(ch.nb[a], e.pad);
            hi[s][a] = __fadd_rn(ch.nb[3 + a], e.pad);
        }
        const double ar = area_of(ch.nb, ch.nb + 3);
        if (ch.left >= 0) {
            ref[s] = ch.index;
            sah += kCostNode * ar;
        } else {
            ref[s] = ~((ch.first << 3) | (ch.count - 1));
            sah += kCostTri * ch.count * ar;
        }
    }
    if (i == 0) sah += kCostNode * area_of(n.nb, n.nb + 3);
    atomicAdd(e.sah, sah);
    atomicMax(e.max_depth, n.depth + 2);
    float4* o = e.nodes2 + 4 * (size_t)n.index;
    o[0] = make_float4(lo[0][0], hi[0][0], lo[0][1], hi[0][1]);
    o[1] = make_float4(lo[1][0], hi[1][0], lo[1][1], hi[1][1]);
    o[2] = make_float4(lo[0][2], hi[0][2], lo[1][2], hi[1][2]);
    o[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.f, 0.f);
    uint32_t w[6];
    for (int s = 0; s < 2; ++s)
        for (int a = 0; a < 3; ++a) {  // bvh_build.cpp quantise2: outwards onto the grid, one more unit per side
            double ql = floor(((double)lo[s][a] - (double)e.qmin[a]) / (double)e.qstep[a]) - 1.0;
            double qh = ceil(((double)hi[s][a] - (double)e.qmin[a]) / (double)e.qstep[a]) + 1.0;
            const uint32_t il = (uint32_t)fmax(0.0, fmin(65535.0, ql)), ih = (uint32_t)fmax(0.0, fmin(65535.0, qh));
            w[3 * s + a] = il | (ih << 16);
        }
    uint4* q = e.qnodes2 + 2 * (size_t)n.index;
    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
    q[1] = make_uint4(w[4], w[5], (uint32_t)ref[0], (uint32_t)ref[1]);
}

// the 48-byte triangle records in leaf order (bvh_build.cpp pack_triangle) and the primitive -> slot map
__global__ void __launch_bounds__(256) k_emit_tris(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t nt,
                                                   const int* __restrict__ final_idx, float4* __restrict__ tris2,
                                                   uint32_t* __restrict__ slot_of_prim) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nt) return;
    const int p = final_idx[s];
    const int32_t* t = tris + 3 * (size_t)p;
    const float* a = verts + 3 * (size_t)t[0];
    const float* b = verts + 3 * (size_t)t[1];
    const float* cc = verts + 3 * (size_t)t[2];
    const V3 va = v3(a[0], a[1], a[2]), vb = v3(b[0], b[1], b[2]), vc = v3(cc[0], cc[1], cc[2]);
    const V3 e1 = rn_sub(va, vb), e2 = rn_sub(vc, va);
    tris2[3 * s + 0] = make_float4(va.x, va.y, va.z, e1.x);
    tris2[3 * s + 1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    tris2[3 * s + 2] = make_float4(e2.z, __int_as_float(p), 0.f, 0.f);
    slot_of_prim[p] = (uint32_t)s;
}

struct Scratch {  // everything cudaMalloc'd for one build; freed on every exit path
    std::vector<void*> ptrs;
    ~Scratch() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    bool alloc(T** p, size_t n) {
        void* q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(n * sizeof(T), 16)) != cudaSuccess) return false;
        ptrs.push_back(q);
        *p = (T*)q;
        return true;
    }
    void release(void* p) {  // hand a buffer over to the caller
        ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), p), ptrs.end());
    }
};

}  // namespace

#define GB_CUDA(call)                                                                \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            err = std::string("device BVH build: ") + #call + ": " + cudaGetErrorString(e_); \
            return false;                                                            \
        }                                                                            \
    } while (0)

bool build_bvh_device(const float* verts_host, int64_t nv, const int32_t* tris_host, int64_t nt, void* stream_, DeviceBVH& out,
                      std::string& err) {
    auto t0 = std::chrono::steady_clock::now();
    cudaStream_t st = (cudaStream_t)stream_;
    if (nt <= 0 || nv <= 0 || !verts_host || !tris_host) {
        err = "build_bvh: empty mesh";
        return false;
    }
    if (nt >= (1ll << 27)) {
        err = "build_bvh: more than 2^27 triangles";
        return false;
    }
    Scratch S;
    float* d_verts;
    int32_t* d_tris;
    float4* d_pbox;
    DNode* d_nodes;
    int *d_idx0, *d_idx1, *d_final, *d_qb0, *d_qb1, *d_qs0, *d_qs1, *d_small;  // d_small: node_count, q_count[4], error, max_depth
    uint32_t* d_keys;
    double* d_sah;
    const size_t qcap = (size_t)nt;
    if (!S.alloc(&d_verts, 3 * (size_t)nv) || !S.alloc(&d_tris, 3 * (size_t)nt) || !S.alloc(&d_pbox, 2 * (size_t)nt) ||
        !S.alloc(&d_nodes, 2 * (size_t)nt + 2) || !S.alloc(&d_idx0, (size_t)nt) || !S.alloc(&d_idx1, (size_t)nt) ||
        !S.alloc(&d_final, (size_t)nt) || !S.alloc(&d_qb0, qcap / kBigCount + 2) || !S.alloc(&d_qb1, qcap / kBigCount + 2) ||
        !S.alloc(&d_qs0, qcap) || !S.alloc(&d_qs1, qcap) || !S.alloc(&d_small, 16) || !S.alloc(&d_keys, 8) ||
        !S.alloc(&d_sah, 2)) {
        cudaGetLastError();
        err = "device BVH build: out of device memory";
        return false;
    }
    GB_CUDA(cudaMemcpyAsync(d_verts, verts_host, 3 * (size_t)nv * sizeof(float), cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_tris, tris_host, 3 * (size_t)nt * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemsetAsync(d_small, 0, 16 * sizeof(int), st));
    GB_CUDA(cudaMemsetAsync(d_sah, 0, 2 * sizeof(double), st));
    const uint32_t keys_init[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u};
    GB_CUDA(cudaMemcpyAsync(d_keys, keys_init, sizeof(keys_init), cudaMemcpyHostToDevice, st));
    Ctx c{};
    c.pbox = d_pbox;
    c.nodes = d_nodes;
    c.node_count = d_small + 0;
    c.q_count = d_small + 1;
    int* d_error = d_small + 5;
    int* d_max_depth = d_small + 6;
    c.idx[0] = d_idx0;
    c.idx[1] = d_idx1;
    c.final_idx = d_final;
    c.q_big[0] = d_qb0;
    c.q_big[1] = d_qb1;
    c.q_small[0] = d_qs0;
    c.q_small[1] = d_qs1;
    const unsigned gprim = (unsigned)((nt + 255) / 256);
    k_prim_boxes<<<gprim, 256, 0, st>>>(d_verts, nv, d_tris, nt, d_pbox, d_idx0, d_keys, d_error);
    k_init_root<<<1, 1, 0, st>>>(c, nt);
    int host_small[8];
    uint32_t host_keys[8];
    GB_CUDA(cudaMemcpyAsync(host_keys, d_keys, sizeof(host_keys), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaMemcpyAsync(host_small, d_small, sizeof(host_small), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    if (host_small[5]) {
        err = "build_bvh: vertex index out of range or non-finite vertex";
        return false;
    }
    auto key_inv = [](uint32_t k) {
        uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
        float f;
        memcpy(&f, &b, 4);
        return f;
    };
    for (int a = 0; a < 3; ++a) {
        out.scene_lo[a] = key_inv(host_keys[a]);
        out.scene_hi[a] = key_inv(host_keys[3 + a]);
    }

    // ---- the levels ------------------------------------------------------------------------------------
    std::vector<int> level_begin;  // node ids [level_begin[d], level_begin[d+1]) have depth d
    level_begin.push_back(0);
    level_begin.push_back(1);
    int nbig = host_small[1], nsmall = host_small[2];
    int depth = 0;
    while (nbig + nsmall > 0) {
        if (depth >= kMaxLevels) {
            err = "device BVH build: tree deeper than the traversal stack allows";
            return false;
        }
        const int par = depth & 1;
        if (nbig > 0) k_build_big<<<(unsigned)nbig, kBigTeam, 0, st>>>(c, par, nbig);
        if (nsmall > 0) k_build_small<<<(unsigned)((nsmall + 3) / 4), 128, 0, st>>>(c, par, nsmall);
        GB_CUDA(cudaMemsetAsync(c.q_count + 2 * par, 0, 2 * sizeof(int), st));
        GB_CUDA(cudaMemcpyAsync(host_small, d_small, sizeof(host_small), cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        GB_CUDA(cudaGetLastError());
        level_begin.push_back(host_small[0]);
        ++depth;
        nbig = host_small[1 + 2 * (depth & 1)];
        nsmall = host_small[2 + 2 * (depth & 1)];
    }
    const int total = host_small[0];
    const int levels = (int)level_begin.size() - 1;

    // ---- depth-first numbering of the inner nodes -------------------------------------------------------
    for (int d = levels - 1; d >= 0; --d) {
        const int b = level_begin[d], e = level_begin[d + 1];
        if (e > b) k_count_inner<<<(unsigned)((e - b + 255) / 256), 256, 0, st>>>(d_nodes, b, e);
    }
    int* d_levels;
    if (!S.alloc(&d_levels, level_begin.size())) {
        cudaGetLastError();
        err = "device BVH build: out of device memory";
        return false;
    }
    GB_CUDA(cudaMemcpyAsync(d_levels, level_begin.data(), level_begin.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    int* d_top_count = d_small + 7;
    k_number_top<<<1, 1, 0, st>>>(d_nodes, d_levels, levels, d_top_count);
    for (int d = std::min(levels, kTopDepth); d < levels; ++d) {
        const int b = level_begin[d], e = level_begin[d + 1];
        if (e > b) k_number_inner<<<(unsigned)((e - b + 255) / 256), 256, 0, st>>>(d_nodes, b, e);
    }
    DNode root;
    GB_CUDA(cudaMemcpyAsync(&root, d_nodes, sizeof(DNode), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    if (root.left < 0) {
        err = "device BVH build: the whole scene is one leaf (host builder handles that)";
        return false;
    }
    const int64_t inner = root.inner;

    // ---- emit --------------------------------------------------------------------------------------------
    const double dx = (double)out.scene_hi[0] - out.scene_lo[0], dy = (double)out.scene_hi[1] - out.scene_lo[1],
                 dz = (double)out.scene_hi[2] - out.scene_lo[2];
    const double diag = std::sqrt(dx * dx + dy * dy + dz * dz);
    double maxc = 0.0;
    for (int a = 0; a < 3; ++a)
        maxc = std::max(maxc, std::max(std::fabs((double)out.scene_lo[a]), std::fabs((double)out.scene_hi[a])));
    float pad = (float)(1e-5 * diag + 4.8e-7 * maxc);
    if (!(pad > 0.f)) pad = 1e-7f;
    EmitArgs e{};
    e.pad = pad;
    for (int a = 0; a < 3; ++a) {  // bvh_build.cpp quantise2 over the padded scene box
        const double lo = (double)(out.scene_lo[a] - pad), hi = (double)(out.scene_hi[a] + pad);
        const double ext = std::max(hi - lo, 1e-30);
        float step = (float)(ext / 65527.0);
        step = std::nextafter(step, FLT_MAX);
        out.qstep[a] = e.qstep[a] = step;
        out.qmin[a] = e.qmin[a] = (float)(lo - 4.0 * (double)step);
    }
    float4 *d_nodes2, *d_tris2;
    uint4* d_qnodes2;
    uint32_t* d_slot;
    if (!S.alloc(&d_nodes2, 4 * (size_t)inner) || !S.alloc(&d_qnodes2, 2 * (size_t)inner) || !S.alloc(&d_tris2, 3 * (size_t)nt) ||
        !S.alloc(&d_slot, (size_t)nt)) {
        cudaGetLastError();
        err = "device BVH build: out of device memory";
        return false;
    }
    e.nodes2 = d_nodes2;
    e.qnodes2 = d_qnodes2;
    e.sah = d_sah;
    e.max_depth = d_max_depth;
    k_emit_nodes<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_nodes, total, e);
    k_emit_tris<<<gprim, 256, 0, st>>>(d_verts, d_tris, nt, d_final, d_tris2, d_slot);
    double h_sah[2];
    GB_CUDA(cudaMemcpyAsync(h_sah, d_sah, sizeof(h_sah), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaMemcpyAsync(host_small, d_small, sizeof(host_small), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    GB_CUDA(cudaGetLastError());
    out.depth = host_small[6];
    out.top_nodes = host_small[7];
    if (out.depth > kStackSize - 2) {
        err = "build_bvh: BVH2 deeper than the traversal stack";
        return false;
    }
    const double root_area = [&] {
        const double ax = (double)root.nb[3] - root.nb[0], ay = (double)root.nb[4] - root.nb[1], az = (double)root.nb[5] - root.nb[2];
        return ax * ay + ay * az + az * ax;
    }();
    out.sah = root_area > 0 ? h_sah[0] / root_area : 0.0;
    out.num_nodes = inner;
    out.d_nodes2 = d_nodes2;
    out.d_qnodes2 = d_qnodes2;
    out.d_tris2 = d_tris2;
    out.d_slot_of_prim = d_slot;
    S.release(d_nodes2);
    S.release(d_qnodes2);
    S.release(d_tris2);
    S.release(d_slot);
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

}  // namespace texir
