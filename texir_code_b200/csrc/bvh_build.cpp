// bvh_build.cpp -- binned-SAH binary BVH over the scene triangles, flattened into the BVH2
// layout and collapsed into the 8-wide compressed layout of texir_bvh.h.
// Must be compiled with -ffp-contract=off (triangle precomputation has to round like the oracle).
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace texir {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; ++a) {
            lo[a] = FLT_MAX;
            hi[a] = -FLT_MAX;
        }
    }
    void grow(const Box& b) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    void grow(const float* p) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    bool empty() const { return lo[0] > hi[0]; }
    double area() const {  // half surface area
        if (empty()) return 0.0;
        double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct BNode {
    Box box;
    int left = -1, right = -1;  // children (inner) or -1 (leaf)
    int first = 0, count = 0;   // range in idx[]
    int depth = 0;
};

#ifndef TX_BINS
#define TX_BINS 16
#endif
#ifndef TX_COST_NODE
#define TX_COST_NODE 1.2
#endif
constexpr int kBins = TX_BINS;
#ifndef TX_MAX_LEAF
#define TX_MAX_LEAF 3
#endif
constexpr int kMaxLeaf = TX_MAX_LEAF;  // triangles per leaf (BVH8 child slots hold at most 3)
constexpr int kSahDepthLimit = 40;  // beyond this depth fall back to object-median splits
constexpr double kCostNode = TX_COST_NODE, kCostTri = 1.0;

struct Builder {
    const float* verts;
    const int32_t* tris;
    int64_t nt;
    std::vector<Box> pbox;
    std::vector<float> pcen;
    std::vector<int> idx;
    std::vector<BNode> nodes;
    std::atomic<int> node_count{0};

    int alloc_node() { return node_count.fetch_add(1); }

    // computes the node's box, decides leaf/split, creates the children.  Returns true if split.
    bool process(int ni) {
        BNode& n = nodes[ni];
        const int first = n.first, count = n.count;
        Box nb, cb;
        nb.reset();
        cb.reset();
        for (int i = first; i < first + count; ++i) {
            int p = idx[i];
            nb.grow(pbox[p]);
            cb.grow(&pcen[3 * (size_t)p]);
        }
        n.box = nb;
        if (count == 1) return false;

        int best_axis = -1, best_split = -1;
        double best_cost = DBL_MAX;
        float scale[3];
        if (n.depth <= kSahDepthLimit) {
            int bcount[3][kBins];
            Box bbox[3][kBins];
            for (int a = 0; a < 3; ++a) {
                float ext = cb.hi[a] - cb.lo[a];
                scale[a] = ext > 0.f ? (float)kBins * 0.999999f / ext : 0.f;
                for (int b = 0; b < kBins; ++b) {
                    bcount[a][b] = 0;
                    bbox[a][b].reset();
                }
            }
            for (int i = first; i < first + count; ++i) {
                int p = idx[i];
                const float* c = &pcen[3 * (size_t)p];
                for (int a = 0; a < 3; ++a) {
                    if (scale[a] == 0.f) continue;
                    int b = std::min(kBins - 1, std::max(0, (int)((c[a] - cb.lo[a]) * scale[a])));
                    bcount[a][b]++;
                    bbox[a][b].grow(pbox[p]);
                }
            }
            for (int a = 0; a < 3; ++a) {
                if (scale[a] == 0.f) continue;
                double larea[kBins];
                int lcount[kBins];
                Box acc;
                acc.reset();
                int cnt = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    acc.grow(bbox[a][b]);
                    cnt += bcount[a][b];
                    larea[b] = acc.area();
                    lcount[b] = cnt;
                }
                acc.reset();
                cnt = 0;
                for (int b = kBins - 1; b >= 1; --b) {
                    acc.grow(bbox[a][b]);
                    cnt += bcount[a][b];
                    int lc = lcount[b - 1];
                    if (lc == 0 || cnt == 0) continue;
                    double cost = larea[b - 1] * lc + acc.area() * cnt;
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = a;
                        best_split = b;
                    }
                }
            }
        }
        if (count <= kMaxLeaf) {
            double leaf_cost = kCostTri * count * nb.area();
            double split_cost = best_axis < 0 ? DBL_MAX : kCostNode * nb.area() + kCostTri * best_cost;
            if (leaf_cost <= split_cost) return false;
        }
        int mid;
        if (best_axis >= 0) {
            const int a = best_axis;
            const float lo = cb.lo[a], sc = scale[a];
            int* b = idx.data() + first;
            int* e = std::partition(b, b + count, [&](int p) {
                int bin = std::min(kBins - 1, std::max(0, (int)((pcen[3 * (size_t)p + a] - lo) * sc)));
                return bin < best_split;
            });
            mid = first + (int)(e - b);
        } else {
            // object-median split on the widest centroid axis (all-equal centroids: split by index)
            int a = 0;
            for (int k = 1; k < 3; ++k)
                if (cb.hi[k] - cb.lo[k] > cb.hi[a] - cb.lo[a]) a = k;
            mid = first + count / 2;
            std::nth_element(idx.begin() + first, idx.begin() + mid, idx.begin() + first + count, [&](int p, int q) {
                float cp = pcen[3 * (size_t)p + a], cq = pcen[3 * (size_t)q + a];
                return cp < cq || (cp == cq && p < q);
            });
        }
        if (mid == first || mid == first + count) mid = first + count / 2;
        int l = alloc_node(), r = alloc_node();
        BNode& nn = nodes[ni];
        nn.left = l;
        nn.right = r;
        nodes[l].first = first;
        nodes[l].count = mid - first;
        nodes[l].depth = nn.depth + 1;
        nodes[r].first = mid;
        nodes[r].count = first + count - mid;
        nodes[r].depth = nn.depth + 1;
        return true;
    }

    void build_subtree(int root) {
        std::vector<int> stack;
        stack.push_back(root);
        while (!stack.empty()) {
            int ni = stack.back();
            stack.pop_back();
            if (process(ni)) {
                stack.push_back(nodes[ni].right);
                stack.push_back(nodes[ni].left);
            }
        }
    }

    void build(int threads) {
        nodes.resize(2 * (size_t)nt + 2);
        int root = alloc_node();
        nodes[root].first = 0;
        nodes[root].count = (int)nt;
        nodes[root].depth = 0;
        // serial top: split the largest open node until there is enough work to share
        std::vector<int> open{root};
        const size_t target = threads > 1 ? (size_t)threads * 16 : 1;
        while (threads > 1 && open.size() < target) {
            size_t bi = 0;
            for (size_t i = 1; i < open.size(); ++i)
                if (nodes[open[i]].count > nodes[open[bi]].count) bi = i;
            if (nodes[open[bi]].count < 4096) break;
            int ni = open[bi];
            open.erase(open.begin() + (long)bi);
            if (process(ni)) {
                open.push_back(nodes[ni].left);
                open.push_back(nodes[ni].right);
            }
        }
        std::sort(open.begin(), open.end(), [&](int a, int b) { return nodes[a].count > nodes[b].count; });
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= open.size()) break;
                build_subtree(open[i]);
            }
        };
        if (threads <= 1) {
            worker();
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
            for (auto& t : pool) t.join();
        }
    }
};

float as_float(int i);
void pack_triangle(const float* verts, const int32_t* tri, int prim, float4* out) {
    V3 a = v3(verts[3 * (size_t)tri[0] + 0], verts[3 * (size_t)tri[0] + 1], verts[3 * (size_t)tri[0] + 2]);
    V3 b = v3(verts[3 * (size_t)tri[1] + 0], verts[3 * (size_t)tri[1] + 1], verts[3 * (size_t)tri[1] + 2]);
    V3 c = v3(verts[3 * (size_t)tri[2] + 0], verts[3 * (size_t)tri[2] + 1], verts[3 * (size_t)tri[2] + 2]);
    V3 e1 = rn_sub(a, b);
    V3 e2 = rn_sub(c, a);
    out[0] = float4{a.x, a.y, a.z, e1.x};
    out[1] = float4{e1.y, e1.z, e2.x, e2.y};
    out[2] = float4{e2.z, as_float(prim), 0.f, 0.f};
}

float as_float(int i) {
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}
uint32_t as_u32(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

Box padded(const Box& b, float pad) {
    Box r = b;
    for (int a = 0; a < 3; ++a) {
        r.lo[a] = b.lo[a] - pad;
        r.hi[a] = b.hi[a] + pad;
    }
    return r;
}

// ----------------------------------------------------------------------------------------------
// BVH2 flattening: top levels breadth-first (a contiguous prefix that a CTA can stage into
// shared memory with one bulk copy), the rest depth-first so that a subtree is contiguous.
// ----------------------------------------------------------------------------------------------
void flatten2(const Builder& B, float pad, int top_depth, HostBVH& out) {
    const std::vector<BNode>& N = B.nodes;
    const int root = 0;
    out.nodes2.clear();
    out.tris2.clear();
    out.prim2.clear();
    out.depth2 = 0;
    auto emit_leaf = [&](int ni) -> int {
        const BNode& n = N[ni];
        int first_slot = (int)out.prim2.size();
        for (int i = 0; i < n.count; ++i) {
            int p = B.idx[n.first + i];
            float4 t[3];
            pack_triangle(B.verts, B.tris + 3 * (size_t)p, p, t);
            out.tris2.push_back(t[0]);
            out.tris2.push_back(t[1]);
            out.tris2.push_back(t[2]);
            out.prim2.push_back((uint32_t)p);
        }
        return ~((first_slot << 3) | (n.count - 1));
    };
    if (N[root].left < 0) {  // single-leaf scene: one node whose two children are the same leaf
        int ref = emit_leaf(root);
        Box b = padded(N[root].box, pad);
        out.nodes2.push_back(float4{b.lo[0], b.hi[0], b.lo[1], b.hi[1]});
        out.nodes2.push_back(float4{b.lo[0], b.hi[0], b.lo[1], b.hi[1]});
        out.nodes2.push_back(float4{b.lo[2], b.hi[2], b.lo[2], b.hi[2]});
        out.nodes2.push_back(float4{as_float(ref), as_float(ref), 0.f, 0.f});
        out.top_nodes2 = 1;
        out.depth2 = 1;
        return;
    }
    // pass 1: assign indices to inner nodes
    std::vector<int> index(B.node_count.load(), -1);
    std::vector<int> order;
    order.reserve(B.node_count.load() / 2 + 1);
    std::vector<int> frontier{root}, nextf;
    for (int d = 0; d < top_depth && !frontier.empty(); ++d) {
        nextf.clear();
        for (int ni : frontier) {
            index[ni] = (int)order.size();
            order.push_back(ni);
            if (N[N[ni].left].left >= 0) nextf.push_back(N[ni].left);
            if (N[N[ni].right].left >= 0) nextf.push_back(N[ni].right);
        }
        frontier.swap(nextf);
    }
    out.top_nodes2 = (int)order.size();
    std::vector<int> stack;
    for (auto it = frontier.rbegin(); it != frontier.rend(); ++it) stack.push_back(*it);
    while (!stack.empty()) {
        int ni = stack.back();
        stack.pop_back();
        index[ni] = (int)order.size();
        order.push_back(ni);
        if (N[N[ni].right].left >= 0) stack.push_back(N[ni].right);
        if (N[N[ni].left].left >= 0) stack.push_back(N[ni].left);
    }
    // pass 2: write nodes in index order; leaves are emitted in that order too
    out.nodes2.resize(4 * order.size());
    double sah = 0.0;
    const double root_area = N[root].box.area();
    for (size_t k = 0; k < order.size(); ++k) {
        const BNode& n = N[order[k]];
        out.depth2 = std::max(out.depth2, n.depth + 2);
        int cref[2];
        Box cb[2];
        const int ch[2] = {n.left, n.right};
        for (int c = 0; c < 2; ++c) {
            const BNode& cn = N[ch[c]];
            cb[c] = padded(cn.box, pad);
            if (cn.left >= 0) {
                cref[c] = index[ch[c]];
                sah += kCostNode * cn.box.area();
            } else {
                cref[c] = emit_leaf(ch[c]);
                sah += kCostTri * cn.count * cn.box.area();
            }
        }
        out.nodes2[4 * k + 0] = float4{cb[0].lo[0], cb[0].hi[0], cb[0].lo[1], cb[0].hi[1]};
        out.nodes2[4 * k + 1] = float4{cb[1].lo[0], cb[1].hi[0], cb[1].lo[1], cb[1].hi[1]};
        out.nodes2[4 * k + 2] = float4{cb[0].lo[2], cb[0].hi[2], cb[1].lo[2], cb[1].hi[2]};
        out.nodes2[4 * k + 3] = float4{as_float(cref[0]), as_float(cref[1]), 0.f, 0.f};
    }
    out.sah2 = root_area > 0 ? (sah + kCostNode * root_area) / root_area : 0.0;
}

// ----------------------------------------------------------------------------------------------
// QBVH2: the flattened BVH2 re-encoded with 32-byte nodes.  Every (already padded) child box is snapped
// outwards onto a global 16-bit grid over the scene box and widened by one more grid unit per side, which
// covers the half-unit error of the traversal's folded decode  t = fma(2^23 + q, step/d, (gmin - o)/d -
// 2^23 step/d)  (texir_walk.h).  At 65 531 steps per axis the widening is ~0.1 mm on a 10 m scene.
// ----------------------------------------------------------------------------------------------
void quantise2(HostBVH& out) {
    const size_t n = out.nodes2.size() / 4;
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (size_t k = 0; k < n; ++k) {
        const float4 &a = out.nodes2[4 * k], &b = out.nodes2[4 * k + 1], &c = out.nodes2[4 * k + 2];
        const float l[2][3] = {{a.x, a.z, c.x}, {b.x, b.z, c.z}}, h[2][3] = {{a.y, a.w, c.y}, {b.y, b.w, c.w}};
        for (int ch = 0; ch < 2; ++ch)
            for (int ax = 0; ax < 3; ++ax) {
                lo[ax] = std::min(lo[ax], (double)l[ch][ax]);
                hi[ax] = std::max(hi[ax], (double)h[ch][ax]);
            }
    }
    for (int ax = 0; ax < 3; ++ax) {
        double ext = std::max(hi[ax] - lo[ax], 1e-30);
        float step = (float)(ext / 65527.0);          // 65535 minus 4 guard units on each side
        step = std::nextafter(step, FLT_MAX);         // never smaller than the exact quotient
        out.qstep[ax] = step;
        out.qmin[ax] = (float)(lo[ax] - 4.0 * (double)step);
    }
    out.qnodes2.resize(2 * n);
    for (size_t k = 0; k < n; ++k) {
        const float4 &a = out.nodes2[4 * k], &b = out.nodes2[4 * k + 1], &c = out.nodes2[4 * k + 2],
                     &d = out.nodes2[4 * k + 3];
        const float l[2][3] = {{a.x, a.z, c.x}, {b.x, b.z, c.z}}, h[2][3] = {{a.y, a.w, c.y}, {b.y, b.w, c.w}};
        uint32_t w[6];
        for (int ch = 0; ch < 2; ++ch)
            for (int ax = 0; ax < 3; ++ax) {
                double ql = std::floor(((double)l[ch][ax] - (double)out.qmin[ax]) / (double)out.qstep[ax]) - 1.0;
                double qh = std::ceil(((double)h[ch][ax] - (double)out.qmin[ax]) / (double)out.qstep[ax]) + 1.0;
                uint32_t il = (uint32_t)std::max(0.0, std::min(65535.0, ql));
                uint32_t ih = (uint32_t)std::max(0.0, std::min(65535.0, qh));
                w[3 * ch + ax] = il | (ih << 16);
            }
        out.qnodes2[2 * k + 0] = uint4{w[0], w[1], w[2], w[3]};
        out.qnodes2[2 * k + 1] = uint4{w[4], w[5], as_u32(d.x), as_u32(d.y)};
    }
}

// ----------------------------------------------------------------------------------------------
// BVH8: greedy collapse (open the child with the largest area until 8), octant-ordered slot
// assignment, conservative 8-bit quantisation.  Nodes are emitted breadth-first so that the
// inner children of a node are contiguous (child_base + rank among inner slots).
// ----------------------------------------------------------------------------------------------
struct Wide {
    int bnode;  // binary node this wide node stands for
    int depth;
};

void build8(const Builder& B, float pad, HostBVH& out) {
    const std::vector<BNode>& N = B.nodes;
    out.nodes8.clear();
    out.tris8.clear();
    out.prim8.clear();
    out.depth8 = 0;
    std::vector<Wide> queue;
    queue.push_back(Wide{0, 1});
    out.nodes8.resize(5);
    size_t head = 0;
    double sah = 0.0;
    const double root_area = N[0].box.area();
    while (head < queue.size()) {
        const Wide w = queue[head];
        const size_t widx = head++;
        out.depth8 = std::max(out.depth8, w.depth);
        const BNode& n = N[w.bnode];
        // gather up to 8 children
        int ch[8];
        int nc = 0;
        if (n.left < 0) {
            ch[nc++] = w.bnode;  // degenerate: the root itself is a leaf
        } else {
            ch[nc++] = n.left;
            ch[nc++] = n.right;
            while (nc < 8) {
                int bi = -1;
                double ba = -1.0;
                for (int i = 0; i < nc; ++i)
                    if (N[ch[i]].left >= 0 && N[ch[i]].box.area() > ba) {
                        ba = N[ch[i]].box.area();
                        bi = i;
                    }
                if (bi < 0) break;
                int c = ch[bi];
                ch[bi] = N[c].left;
                ch[nc++] = N[c].right;
            }
        }
        // slot assignment: slot bit a set <=> child on the positive side along axis a
        Box nb = padded(n.box, pad);
        float cen[3];
        for (int a = 0; a < 3; ++a) cen[a] = 0.5f * (nb.lo[a] + nb.hi[a]);
        float cost[8][8];
        for (int i = 0; i < nc; ++i) {
            const Box& cb = N[ch[i]].box;
            float dc[3];
            for (int a = 0; a < 3; ++a) dc[a] = 0.5f * (cb.lo[a] + cb.hi[a]) - cen[a];
            for (int s = 0; s < 8; ++s) {
                float c = 0.f;
                for (int a = 0; a < 3; ++a) c += ((s >> a) & 1) ? -dc[a] : dc[a];
                cost[i][s] = c;
            }
        }
        int slot_child[8];
        for (int s = 0; s < 8; ++s) slot_child[s] = -1;
        bool assigned[8] = {false, false, false, false, false, false, false, false};
        for (int k = 0; k < nc; ++k) {
            float best = FLT_MAX;
            int bi = -1, bs = -1;
            for (int i = 0; i < nc; ++i) {
                if (assigned[i]) continue;
                for (int s = 0; s < 8; ++s) {
                    if (slot_child[s] >= 0) continue;
                    if (cost[i][s] < best) {
                        best = cost[i][s];
                        bi = i;
                        bs = s;
                    }
                }
            }
            assigned[bi] = true;
            slot_child[bs] = ch[bi];
        }
        // quantisation grid: p = box.lo, 2^e * 255 >= extent
        uint32_t ebyte[3];
        float grid[3];
        for (int a = 0; a < 3; ++a) {
            float ext = nb.hi[a] - nb.lo[a];
            int e = ext > 0.f ? (int)std::ceil(std::log2((double)ext / 255.0)) : -126;
            e = std::max(-126, std::min(127, e));
            while (e < 127 && (double)nb.lo[a] + std::ldexp(255.0, e) < (double)nb.hi[a]) ++e;
            ebyte[a] = (uint32_t)(e + 127);
            grid[a] = std::ldexp(1.0f, e);
        }
        uint8_t meta[8], qlo[3][8], qhi[3][8];
        uint32_t imask = 0;
        const uint32_t child_base = (uint32_t)queue.size();
        const uint32_t tri_base = (uint32_t)out.prim8.size();
        uint32_t tri_off = 0;
        for (int s = 0; s < 8; ++s) {
            meta[s] = 0;
            for (int a = 0; a < 3; ++a) qlo[a][s] = qhi[a][s] = 0;
            int c = slot_child[s];
            if (c < 0) continue;
            const BNode& cn = N[c];
            Box cb = padded(cn.box, pad);
            for (int a = 0; a < 3; ++a) {
                double l = std::floor(((double)cb.lo[a] - (double)nb.lo[a]) / (double)grid[a]);
                double h = std::ceil(((double)cb.hi[a] - (double)nb.lo[a]) / (double)grid[a]);
                qlo[a][s] = (uint8_t)std::max(0.0, std::min(255.0, l));
                qhi[a][s] = (uint8_t)std::max(0.0, std::min(255.0, h));
            }
            if (cn.left >= 0) {
                imask |= 1u << s;
                meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
                queue.push_back(Wide{c, w.depth + 1});
                sah += kCostNode * cn.box.area();
            } else {
                uint32_t unary = (1u << cn.count) - 1u;  // 1 -> 001, 2 -> 011, 3 -> 111
                meta[s] = (uint8_t)((unary << 5) | tri_off);
                for (int i = 0; i < cn.count; ++i) {
                    int p = B.idx[cn.first + i];
                    float4 t[3];
                    pack_triangle(B.verts, B.tris + 3 * (size_t)p, p, t);
                    out.tris8.push_back(t[0]);
                    out.tris8.push_back(t[1]);
                    out.tris8.push_back(t[2]);
                    out.prim8.push_back((uint32_t)p);
                }
                tri_off += (uint32_t)cn.count;
                sah += kCostTri * cn.count * cn.box.area();
            }
        }
        out.nodes8.resize(5 * queue.size());
        auto pack4 = [](const uint8_t* b) {
            return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        };
        uint4* q = &out.nodes8[5 * widx];
        q[0] = uint4{as_u32(nb.lo[0]), as_u32(nb.lo[1]), as_u32(nb.lo[2]),
                     ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (imask << 24)};
        q[1] = uint4{child_base, tri_base, pack4(meta), pack4(meta + 4)};
        q[2] = uint4{pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4)};
        q[3] = uint4{pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4)};
        q[4] = uint4{pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4)};
    }
    out.sah8 = root_area > 0 ? (sah + kCostNode * root_area) / root_area : 0.0;
}

}  // namespace

bool build_bvh(const float* verts, int64_t nv, const int32_t* tris, int64_t nt, int threads, HostBVH& out,
               std::string& err, bool want8) {
    auto t0 = std::chrono::steady_clock::now();
    if (nt <= 0 || nv <= 0 || !verts || !tris) {
        err = "build_bvh: empty mesh";
        return false;
    }
    if (nt >= (1ll << 27)) {
        err = "build_bvh: more than 2^27 triangles";
        return false;
    }
    Builder B;
    B.verts = verts;
    B.tris = tris;
    B.nt = nt;
    B.pbox.resize((size_t)nt);
    B.pcen.resize(3 * (size_t)nt);
    B.idx.resize((size_t)nt);
    Box scene;
    scene.reset();
    for (int64_t p = 0; p < nt; ++p) {
        Box b;
        b.reset();
        for (int k = 0; k < 3; ++k) {
            int32_t vi = tris[3 * p + k];
            if (vi < 0 || vi >= nv) {
                err = "build_bvh: vertex index out of range";
                return false;
            }
            const float* v = verts + 3 * (size_t)vi;
            if (!(std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]))) {
                err = "build_bvh: non-finite vertex";
                return false;
            }
            b.grow(v);
        }
        B.pbox[(size_t)p] = b;
        for (int a = 0; a < 3; ++a) B.pcen[3 * (size_t)p + a] = 0.5f * (b.lo[a] + b.hi[a]);
        B.idx[(size_t)p] = (int)p;
        scene.grow(b);
    }
    for (int a = 0; a < 3; ++a) {
        out.scene_lo[a] = scene.lo[a];
        out.scene_hi[a] = scene.hi[a];
    }
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = std::min(threads, 64);
    auto tp0 = std::chrono::steady_clock::now();
    B.build(threads);
    auto tp1 = std::chrono::steady_clock::now();
    // boxes are padded by 1e-5 of the scene diagonal plus 4 ulp of the largest coordinate: larger
    // than any rounding in the slab tests or in the triangle test's edge decisions
    double dx = (double)scene.hi[0] - scene.lo[0], dy = (double)scene.hi[1] - scene.lo[1],
           dz = (double)scene.hi[2] - scene.lo[2];
    double diag = std::sqrt(dx * dx + dy * dy + dz * dz);
    double maxc = 0.0;
    for (int a = 0; a < 3; ++a) maxc = std::max(maxc, std::max(std::fabs((double)scene.lo[a]), std::fabs((double)scene.hi[a])));
    float pad = (float)(1e-5 * diag + 4.8e-7 * maxc);
    if (!(pad > 0.f)) pad = 1e-7f;
    flatten2(B, pad, /*top_depth=*/10, out);
    auto tp2 = std::chrono::steady_clock::now();
    quantise2(out);
    auto tp3 = std::chrono::steady_clock::now();
    if (want8) build8(B, pad, out);
    auto tp4 = std::chrono::steady_clock::now();
    if (getenv("TEXIR_BUILD_TIMING")) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "bvh build: prims %.1f ms, tree %.1f ms, flatten %.1f ms, quantise %.1f ms, bvh8 %.1f ms\n",
                ms(t0, tp0), ms(tp0, tp1), ms(tp1, tp2), ms(tp2, tp3), ms(tp3, tp4));
    }
    if (out.depth2 > kStackSize - 2) {
        err = "build_bvh: BVH2 deeper than the traversal stack";
        return false;
    }
    if (want8 && out.depth8 > kStack8 - 2) {
        err = "build_bvh: BVH8 deeper than the traversal stack";
        return false;
    }
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

}  // namespace texir
