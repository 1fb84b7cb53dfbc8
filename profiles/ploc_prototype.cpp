// ploc_prototype.cpp -- sequential prototype of the device BVH builder's algorithm (PLOC: parallel locally-ordered
// clustering, Meister & Bittner 2018), used to judge tree QUALITY against the host binned-SAH builder before the
// CUDA version was written (profiles/ploc_quality.py).  Not part of the product: it includes bvh_build.cpp to reuse
// its flattening.   g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -pthread -I texir_code_b200/csrc ...
#include "../texir_code_b200/csrc/bvh_build.cpp"

#include <numeric>

namespace texir {
namespace {

uint64_t spread21(uint64_t v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

struct PNode {
    Box box;
    int left = -1, right = -1;  // -1: single primitive
    int prim = -1;
    int tris = 1;
    double cost = 0.0;
    bool leaf = true;  // after the SAH collapse: this subtree is emitted as one leaf
};

// builds Builder-compatible nodes (B.nodes / B.idx / node_count) from a PLOC tree
void ploc_build(Builder& B, int radius, int max_leaf) {
    const int64_t nt = B.nt;
    Box scene;
    scene.reset();
    for (int64_t p = 0; p < nt; ++p) scene.grow(B.pbox[p]);
    std::vector<uint64_t> code((size_t)nt);
    for (int64_t p = 0; p < nt; ++p) {
        uint64_t c = 0;
        for (int a = 0; a < 3; ++a) {
            double ext = (double)scene.hi[a] - scene.lo[a];
            double q = ext > 0 ? ((double)B.pcen[3 * p + a] - scene.lo[a]) / ext : 0.0;
            uint64_t cell = (uint64_t)std::min(2097151.0, std::max(0.0, q * 2097152.0));
            c |= spread21(cell) << a;
        }
        code[p] = c;
    }
    std::vector<int> order((size_t)nt);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b] || (code[a] == code[b] && a < b); });
    std::vector<PNode> N;
    N.reserve(2 * (size_t)nt);
    std::vector<int> C((size_t)nt);
    for (int64_t i = 0; i < nt; ++i) {
        PNode n;
        n.box = B.pbox[order[i]];
        n.prim = order[i];
        n.cost = kCostTri * n.box.area();
        N.push_back(n);
        C[i] = (int)i;
    }
    std::vector<int> nn, next;
    int iters = 0;
    while (C.size() > 1) {
        const int n = (int)C.size();
        nn.assign(n, -1);
        for (int i = 0; i < n; ++i) {
            double best = DBL_MAX;
            int bj = -1;
            const Box& bi = N[C[i]].box;
            for (int j = std::max(0, i - radius); j <= std::min(n - 1, i + radius); ++j) {
                if (j == i) continue;
                Box u = bi;
                u.grow(N[C[j]].box);
                double a = u.area();
                if (a < best) {
                    best = a;
                    bj = j;
                }
            }
            nn[i] = bj;
        }
        next.clear();
        for (int i = 0; i < n; ++i) {
            const int j = nn[i];
            if (nn[j] == i) {
                if (i < j) {
                    PNode m;
                    m.box = N[C[i]].box;
                    m.box.grow(N[C[j]].box);
                    m.left = C[i];
                    m.right = C[j];
                    m.tris = N[C[i]].tris + N[C[j]].tris;
                    const double leaf_cost = kCostTri * m.tris * m.box.area();
                    const double split_cost = kCostNode * m.box.area() + N[C[i]].cost + N[C[j]].cost;
                    m.leaf = m.tris <= max_leaf && leaf_cost <= split_cost;
                    m.cost = m.leaf ? leaf_cost : split_cost;
                    N.push_back(m);
                    next.push_back((int)N.size() - 1);
                }
            } else {
                next.push_back(C[i]);
            }
        }
        C.swap(next);
        ++iters;
    }
    fprintf(stderr, "ploc: %d iterations, %zu nodes\n", iters, N.size());
    // convert: DFS from the root, collapsed subtrees become leaves with contiguous idx ranges
    B.nodes.assign(2 * (size_t)nt + 2, BNode());
    B.node_count.store(0);
    B.idx.clear();
    struct Item {
        int pn, bn, depth;
    };
    std::vector<Item> st;
    const int root = B.alloc_node();
    st.push_back({C[0], root, 0});
    std::vector<int> tmp;
    while (!st.empty()) {
        Item it = st.back();
        st.pop_back();
        const PNode& p = N[it.pn];
        BNode& b = B.nodes[it.bn];
        b.box = p.box;
        b.depth = it.depth;
        if (p.leaf) {
            b.first = (int)B.idx.size();
            tmp.assign(1, it.pn);
            while (!tmp.empty()) {
                int q = tmp.back();
                tmp.pop_back();
                if (N[q].left < 0) {
                    B.idx.push_back(N[q].prim);
                } else {
                    tmp.push_back(N[q].right);
                    tmp.push_back(N[q].left);
                }
            }
            b.count = (int)B.idx.size() - b.first;
            b.left = b.right = -1;
        } else {
            const int l = B.alloc_node(), r = B.alloc_node();
            B.nodes[it.bn].left = l;
            B.nodes[it.bn].right = r;
            B.nodes[it.bn].count = p.tris;
            st.push_back({p.right, r, it.depth + 1});
            st.push_back({p.left, l, it.depth + 1});
        }
    }
}

}  // namespace
}  // namespace texir

using namespace texir;

// mode 0: host binned SAH; mode 1: PLOC(radius).  rays [R,6].  out: [0] sah, [1] nodes/ray, [2] tris/ray, [3] depth, [4] inner nodes
extern "C" int ploc_quality(const float* verts, int64_t nv, const int32_t* tris, int64_t nt, int mode, int radius, int max_leaf,
                            int threads, const float* rays, int64_t R, double* out, uint32_t* prim_out, float* t_out) {
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
        for (int k = 0; k < 3; ++k) b.grow(verts + 3 * (size_t)tris[3 * p + k]);
        B.pbox[p] = b;
        for (int a = 0; a < 3; ++a) B.pcen[3 * p + a] = 0.5f * (b.lo[a] + b.hi[a]);
        B.idx[p] = (int)p;
        scene.grow(b);
    }
    if (mode == 0)
        B.build(threads);
    else
        ploc_build(B, radius, max_leaf);
    double dx = (double)scene.hi[0] - scene.lo[0], dy = (double)scene.hi[1] - scene.lo[1], dz = (double)scene.hi[2] - scene.lo[2];
    double diag = std::sqrt(dx * dx + dy * dy + dz * dz), maxc = 0.0;
    for (int a = 0; a < 3; ++a) maxc = std::max(maxc, std::max(std::fabs((double)scene.lo[a]), std::fabs((double)scene.hi[a])));
    float pad = (float)(1e-5 * diag + 4.8e-7 * maxc);
    HostBVH hb;
    flatten2(B, pad, 10, hb);
    SceneView s{};
    s.nodes2 = hb.nodes2.data();
    s.tris2 = hb.tris2.data();
    s.num_tris = (uint32_t)nt;
    double nvs = 0, nts = 0;
    for (int64_t r = 0; r < R; ++r) {
        uint32_t a = 0, b = 0;
        Hit h = trace_bvh2(s, v3(rays[6 * r], rays[6 * r + 1], rays[6 * r + 2]), v3(rays[6 * r + 3], rays[6 * r + 4], rays[6 * r + 5]), &a, &b);
        nvs += a;
        nts += b;
        if (prim_out) prim_out[r] = h.prim;
        if (t_out) t_out[r] = h.t;
    }
    out[0] = hb.sah2;
    out[1] = nvs / (double)R;
    out[2] = nts / (double)R;
    out[3] = hb.depth2;
    out[4] = (double)hb.nodes2.size() / 4;
    return 0;
}
