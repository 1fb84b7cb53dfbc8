/* cast_rays_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the one piece of the TexIR hot path whose arithmetic is not in the reference
 * repository: Open3D's RaycastingScene.cast_rays (open3d, version unpinned in the reference's
 * requirements.txt:4; Open3D 0.15-0.17 bundle Embree 3.13.x), called at
 * models/mat_nvdiffrast.py:294-295 and models/tracer_o3d_irt.py:243-244.
 *
 * Restated algorithm = Embree 3's single-ray Moeller-Trumbore triangle intersector
 * (kernels/geometry/triangle_intersector_moeller.h, MoellerTrumboreIntersector1) applied to every
 * triangle, closest hit wins:
 *     e1 = v0 - v1, e2 = v2 - v0, Ng = e2 x e1, C = v0 - O, R = C x D, den = Ng . D,
 *     U = (R . e2) sgn(den), V = (R . e1) sgn(den), T = (Ng . C) sgn(den),
 *     hit iff den != 0, U >= 0, V >= 0, U + V <= |den|, |den| * tnear < T <= |den| * tfar,
 *     u = U/|den|, v = V/|den|, t = T/|den|        (u weights v1, v if v2; Open3D's primitive_uvs)
 * with Open3D's ray setup tnear = 0, tfar = inf, no culling (default scene flags), and Open3D's miss
 * values t_hit = inf, primitive_ids = 0xFFFFFFFF, primitive_uvs = 0.  Embree computes 1/|den| with a
 * Newton-refined rcp; here it is one IEEE division.  Tie-break between equal t: lowest primitive id
 * (Embree's choice is unspecified).  PARITY AGAINST REAL EMBREE IS UNPINNED: open3d cannot be
 * installed in this environment and the reference holds no golden vectors for this call.
 *
 * Two evaluation modes that must agree bit-for-bit (tests/test_oracle.py):
 *   brute = 1: every ray against every triangle in primitive order;
 *   brute = 0: a median-split BVH (leaf <= 4) with boxes padded by 1e-5 of the scene diagonal, so
 *              1M-triangle scenes are usable as the CPU baseline (bench.py, cpu_baseline leg).
 * Every product-sum is an explicit fmaf chain; compile with -ffp-contract=off.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float v0[3], e1[3], e2[3], ng[3];
} otri_t;

typedef struct {
    float lo[3], hi[3];
    int32_t left, right; /* inner: children; leaf: left = -1 */
    int32_t first, count;
} onode_t;

typedef struct {
    int64_t nt;
    otri_t* tris;    /* primitive order */
    int32_t* order;  /* BVH leaf order -> primitive id */
    onode_t* nodes;
    int32_t num_nodes;
    float pad;
} oscene_t;

static inline float dot3(const float* a, const float* b) { return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])); }
static inline void cross3(const float* a, const float* b, float* r) {
    r[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    r[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    r[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}

static inline int tri_hit(const otri_t* tr, const float* o, const float* d, float* t, float* u, float* v) {
    float c[3] = {tr->v0[0] - o[0], tr->v0[1] - o[1], tr->v0[2] - o[2]};
    float r[3];
    cross3(c, d, r);
    float den = dot3(tr->ng, d);
    float aden = fabsf(den);
    float sgn = den < 0.f ? -1.0f : 1.0f;
    float U = dot3(r, tr->e2) * sgn;
    float V = dot3(r, tr->e1) * sgn;
    float T = dot3(tr->ng, c) * sgn;
    if (!(den != 0.f && U >= 0.f && V >= 0.f && U + V <= aden && T > 0.f)) return 0;
    float rcp = 1.0f / aden;
    *t = T * rcp;
    *u = U * rcp;
    *v = V * rcp;
    return 1;
}

/* ------------------------------------------------------------------------------------------- */
static void tri_bounds(const otri_t* tr, const float* verts, const int32_t* idx, float* lo, float* hi) {
    (void)tr;
    for (int a = 0; a < 3; ++a) {
        float x0 = verts[3 * (size_t)idx[0] + a], x1 = verts[3 * (size_t)idx[1] + a], x2 = verts[3 * (size_t)idx[2] + a];
        lo[a] = fminf(x0, fminf(x1, x2));
        hi[a] = fmaxf(x0, fmaxf(x1, x2));
    }
}

typedef struct {
    const float* cen;
    int axis;
} sortctx_t;
static sortctx_t g_sort; /* build is single-threaded */
static int cmp_centroid(const void* pa, const void* pb) {
    int32_t a = *(const int32_t*)pa, b = *(const int32_t*)pb;
    float ca = g_sort.cen[3 * (size_t)a + g_sort.axis], cb = g_sort.cen[3 * (size_t)b + g_sort.axis];
    if (ca < cb) return -1;
    if (ca > cb) return 1;
    return (a > b) - (a < b);
}

static int32_t build_rec(oscene_t* s, const float* plo, const float* phi, const float* cen, int32_t first,
                         int32_t count) {
    int32_t ni = s->num_nodes++;
    onode_t* n = &s->nodes[ni];
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int a = 0; a < 3; ++a) {
        n->lo[a] = INFINITY;
        n->hi[a] = -INFINITY;
    }
    for (int32_t i = first; i < first + count; ++i) {
        int32_t p = s->order[i];
        for (int a = 0; a < 3; ++a) {
            n->lo[a] = fminf(n->lo[a], plo[3 * (size_t)p + a]);
            n->hi[a] = fmaxf(n->hi[a], phi[3 * (size_t)p + a]);
            clo[a] = fminf(clo[a], cen[3 * (size_t)p + a]);
            chi[a] = fmaxf(chi[a], cen[3 * (size_t)p + a]);
        }
    }
    for (int a = 0; a < 3; ++a) {
        n->lo[a] -= s->pad;
        n->hi[a] += s->pad;
    }
    n->first = first;
    n->count = count;
    n->left = n->right = -1;
    if (count <= 4) return ni;
    int axis = 0;
    for (int a = 1; a < 3; ++a)
        if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
    g_sort.cen = cen;
    g_sort.axis = axis;
    qsort(s->order + first, (size_t)count, sizeof(int32_t), cmp_centroid);
    int32_t half = count / 2;
    int32_t l = build_rec(s, plo, phi, cen, first, half);
    int32_t r = build_rec(s, plo, phi, cen, first + half, count - half);
    s->nodes[ni].left = l;
    s->nodes[ni].right = r;
    return ni;
}

void* oracle_scene_create(const float* verts, int64_t nv, const int32_t* tris, int64_t nt) {
    (void)nv;
    if (nt <= 0) return NULL;
    oscene_t* s = (oscene_t*)calloc(1, sizeof(oscene_t));
    s->nt = nt;
    s->tris = (otri_t*)malloc(sizeof(otri_t) * (size_t)nt);
    s->order = (int32_t*)malloc(sizeof(int32_t) * (size_t)nt);
    s->nodes = (onode_t*)malloc(sizeof(onode_t) * (size_t)(2 * nt + 1));
    float* plo = (float*)malloc(sizeof(float) * 3 * (size_t)nt);
    float* phi = (float*)malloc(sizeof(float) * 3 * (size_t)nt);
    float* cen = (float*)malloc(sizeof(float) * 3 * (size_t)nt);
    float slo[3] = {INFINITY, INFINITY, INFINITY}, shi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t p = 0; p < nt; ++p) {
        const int32_t* ix = tris + 3 * p;
        otri_t* t = &s->tris[p];
        for (int a = 0; a < 3; ++a) {
            float x0 = verts[3 * (size_t)ix[0] + a], x1 = verts[3 * (size_t)ix[1] + a], x2 = verts[3 * (size_t)ix[2] + a];
            t->v0[a] = x0;
            t->e1[a] = x0 - x1;
            t->e2[a] = x2 - x0;
        }
        cross3(t->e2, t->e1, t->ng);
        tri_bounds(t, verts, ix, plo + 3 * p, phi + 3 * p);
        for (int a = 0; a < 3; ++a) {
            cen[3 * p + a] = 0.5f * (plo[3 * p + a] + phi[3 * p + a]);
            slo[a] = fminf(slo[a], plo[3 * p + a]);
            shi[a] = fmaxf(shi[a], phi[3 * p + a]);
        }
        s->order[p] = (int32_t)p;
    }
    double dx = (double)shi[0] - slo[0], dy = (double)shi[1] - slo[1], dz = (double)shi[2] - slo[2];
    double maxc = 0.0;
    for (int a = 0; a < 3; ++a) maxc = fmax(maxc, fmax(fabs((double)slo[a]), fabs((double)shi[a])));
    s->pad = (float)(1e-5 * sqrt(dx * dx + dy * dy + dz * dz) + 4.8e-7 * maxc);
    if (!(s->pad > 0.f)) s->pad = 1e-7f;
    s->num_nodes = 0;
    build_rec(s, plo, phi, cen, 0, (int32_t)nt);
    free(plo);
    free(phi);
    free(cen);
    return s;
}

void oracle_scene_destroy(void* h) {
    oscene_t* s = (oscene_t*)h;
    if (!s) return;
    free(s->tris);
    free(s->order);
    free(s->nodes);
    free(s);
}

static inline void consider(const oscene_t* s, int32_t p, const float* o, const float* d, float* bt, float* bu,
                            float* bv, uint32_t* bid) {
    float t, u, v;
    if (tri_hit(&s->tris[p], o, d, &t, &u, &v)) {
        if (t < *bt || (t == *bt && (uint32_t)p < *bid)) {
            *bt = t;
            *bu = u;
            *bv = v;
            *bid = (uint32_t)p;
        }
    }
}

static void cast_one(const oscene_t* s, const float* ray, int brute, float* t_out, uint32_t* id_out, float* uv_out) {
    const float* o = ray;
    const float* d = ray + 3;
    float bt = INFINITY, bu = 0.f, bv = 0.f;
    uint32_t bid = 0xFFFFFFFFu;
    if (brute) {
        for (int64_t p = 0; p < s->nt; ++p) consider(s, (int32_t)p, o, d, &bt, &bu, &bv, &bid);
    } else {
        float id[3];
        for (int a = 0; a < 3; ++a) {
            float x = d[a];
            if (!(fabsf(x) > 1e-24f)) x = copysignf(1e-24f, x);
            id[a] = 1.0f / x;
        }
        int32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const onode_t* n = &s->nodes[stack[--sp]];
            float tmin = 0.f, tmax = bt;
            for (int a = 0; a < 3; ++a) {
                float t0 = (n->lo[a] - o[a]) * id[a], t1 = (n->hi[a] - o[a]) * id[a];
                tmin = fmaxf(tmin, fminf(t0, t1));
                tmax = fminf(tmax, fmaxf(t0, t1));
            }
            if (!(tmin * 0.9999996f <= tmax * 1.0000004f)) continue;
            if (n->left < 0) {
                for (int32_t i = n->first; i < n->first + n->count; ++i) consider(s, s->order[i], o, d, &bt, &bu, &bv, &bid);
            } else {
                stack[sp++] = n->left;
                stack[sp++] = n->right;
            }
        }
    }
    *t_out = bt;
    *id_out = bid;
    uv_out[0] = bu;
    uv_out[1] = bv;
}

typedef struct {
    const oscene_t* s;
    const float* rays;
    int64_t begin, end;
    int brute;
    float* t;
    uint32_t* id;
    float* uv;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (int64_t r = j->begin; r < j->end; ++r) cast_one(j->s, j->rays + 6 * r, j->brute, j->t + r, j->id + r, j->uv + 2 * r);
    return NULL;
}

/* rays [R,6] -> t [R], id [R], uv [R,2].  threads <= 1: run in the calling thread. */
void oracle_cast_rays(void* h, const float* rays, int64_t R, int brute, int threads, float* t, uint32_t* id, float* uv) {
    const oscene_t* s = (const oscene_t*)h;
    if (threads > 256) threads = 256;
    if (threads <= 1 || R < 1024) {
        job_t j = {s, rays, 0, R, brute, t, id, uv};
        worker(&j);
        return;
    }
    pthread_t th[256];
    job_t jobs[256];
    int64_t chunk = (R + threads - 1) / threads;
    int n = 0;
    for (int k = 0; k < threads; ++k) {
        int64_t b = k * chunk, e = b + chunk < R ? b + chunk : R;
        if (b >= e) break;
        jobs[n] = (job_t){s, rays, b, e, brute, t, id, uv};
        pthread_create(&th[n], NULL, worker, &jobs[n]);
        ++n;
    }
    for (int k = 0; k < n; ++k) pthread_join(th[k], NULL);
}
