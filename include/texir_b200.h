/* texir_b200.h -- C ABI of libtexir_b200.so: the B200 (sm_100a) implementation of TexIR's
 * irradiance-texture / per-pixel BRDF hot path.
 *
 * The reference (LZleejean/TexIR_code) is pure Python; it has no FFI of its own.  Each entry
 * point below replaces one piece of reference Python (file:line cited per function) and is what
 * the ctypes binding in texir_code_b200/_lib.py -- or a binding a reference maintainer would add,
 * see INTEGRATION.md -- calls.
 *
 * Conventions
 *   - every pointer is a CUDA DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     every call is asynchronous on that stream and performs no host synchronisation,
 *     except texir_scene_create / texir_scene_set_texture(host source), which upload;
 *   - tensors are dense row-major float32 unless stated otherwise;
 *   - return value 0 = ok; non-zero = error, text from texir_last_error() (thread-local);
 *   - no C++ exception crosses this boundary;
 *   - a scene is immutable after creation apart from set_texture / set_bvh, which must not race
 *     with launches that use the scene.
 */
#ifndef TEXIR_B200_H
#define TEXIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TEXIR_API __declspec(dllexport)
#else
#define TEXIR_API __attribute__((visibility("default")))
#endif

typedef struct texir_scene texir_scene;

/* ranks of one node that can map each other's buffers (texir_peer_*), and the size of an exported buffer handle */
#define TEXIR_MAX_PEERS 8
#define TEXIR_PEER_HANDLE_BYTES 64

/* sampling modes of generate_dir (utils/sample_util.py:115-143) */
enum { TEXIR_MODE_UNIFORM = 0, TEXIR_MODE_COSINE = 1, TEXIR_MODE_IMPORTANCE = 2 };
/* radiance texture source layouts */
enum { TEXIR_TEX_CHW = 0 /* [3,H,W], the reference's self.texture[0] */, TEXIR_TEX_HWC = 1 /* [H,W,3] */ };
/* BVH layout a scene's kernels traverse */
enum { TEXIR_BVH2 = 2, TEXIR_BVH8 = 8 };
/* who builds the tree: AUTO = the device for scenes of >= 4096 triangles (environment TEXIR_BUILDER=1 / 2 forces host / device) */
enum { TEXIR_BUILDER_AUTO = 0, TEXIR_BUILDER_HOST = 1, TEXIR_BUILDER_DEVICE = 2 };

typedef struct texir_scene_info {
    int64_t num_tris;
    int64_t num_nodes2, num_nodes8;
    int32_t depth2, depth8;
    int32_t tex_h, tex_w;
    int32_t bvh_kind;       /* TEXIR_BVH2 or TEXIR_BVH8 */
    int32_t device;
    double sah2, sah8;
    double build_seconds;   /* wall time of texir_scene_create: build + uploads */
    int64_t device_bytes;   /* HBM held by the scene (BVHs + triangles + uvs + texture) */
    int32_t builder;        /* TEXIR_BUILDER_HOST or TEXIR_BUILDER_DEVICE: who built the resident tree */
    int32_t reserved;
} texir_scene_info;

TEXIR_API const char* texir_last_error(void);
TEXIR_API int texir_version(void);
/* number of kernels this library has launched in this process (all scenes, all streams) */
TEXIR_API int64_t texir_launch_count(void);

/* ---- scene: replaces o3d.t.geometry.RaycastingScene() + add_triangles + self.triangle_uvs
 *      (models/mat_nvdiffrast.py:87,97-101; models/tracer_o3d_irt.py:75,85-89) ------------------
 * verts_host [nv,3] f32, tris_host [nt,3] i32, tri_uvs_host [3*nt,2] f64 (np.asarray(mesh.triangle_uvs),
 * may be NULL for a pure cast_rays scene).  Builds the BVH -- on `device` itself for scenes of >= 4096 triangles,
 * else on the host (build_threads <= 0: all cores) -- and keeps it resident there. */
TEXIR_API int texir_scene_create(const float* verts_host, int64_t nv, const int32_t* tris_host, int64_t nt,
                                 const double* tri_uvs_host, int device, int build_threads, texir_scene** out);
/* The same with the builder chosen explicitly (TEXIR_BUILDER_*).  Both builders run the same binned-SAH algorithm with
 * the same arithmetic and produce the same tree (node boxes, splits, leaf contents); the device one (csrc/
 * bvh_build_gpu.cu: level-synchronous, a CTA or a warp per node) takes milliseconds where the host one takes most of a
 * second.  The 8-wide layout is built lazily, on the host, by the first texir_scene_set_bvh(scene, TEXIR_BVH8). */
TEXIR_API int texir_scene_create_ex(const float* verts_host, int64_t nv, const int32_t* tris_host, int64_t nt,
                                    const double* tri_uvs_host, int device, int build_threads, int builder,
                                    texir_scene** out);
TEXIR_API int texir_scene_destroy(texir_scene* scene);
TEXIR_API int texir_scene_get_info(const texir_scene* scene, texir_scene_info* info);
TEXIR_API int texir_scene_set_bvh(texir_scene* scene, int bvh_kind);

/* ---- radiance texture: replaces self.texture (mat_nvdiffrast.py:89-93), already flipped and
 *      exposed by the caller.  tex is f32, layout TEXIR_TEX_*; src_on_device != 0: device pointer. */
TEXIR_API int texir_scene_set_texture(texir_scene* scene, const float* tex, int H, int W, int layout,
                                      int src_on_device, void* stream);

/* The same texture kept as the 4-byte Radiance RGBE texels of its .hdr file (the reference's texture always comes from
 * hdr_texture.hdr, mat_nvdiffrast.py:89): rgbe [H,W,4] bytes in file order (R, G, B mantissas, shared exponent; what
 * texir_hdr_read_host returns), flip_rows != 0 = cv2.flip(texture, 0) (:90), exposure = train.hdr_exposure (an integer:
 * a power-of-two factor folds into the exponent).  The kernels decode mantissa * 2^(E - 136 + exposure) per tap, which
 * is EXACTLY the float32 value of the cv2.imread + multiply path -- results are bit-identical to
 * texir_scene_set_texture of the decoded texture, with a quarter of the texture bytes in HBM / L2.  The texture set
 * last (this call or texir_scene_set_texture) is the active one. */
TEXIR_API int texir_scene_set_texture_rgbe(texir_scene* scene, const uint8_t* rgbe, int H, int W, int flip_rows,
                                           int exposure, int src_on_device, void* stream);

/* ---- cast_rays: replaces RaycastingScene.cast_rays (mat_nvdiffrast.py:294-295).
 * rays [R,6] = (origin, direction).  Outputs: t_hit [R] (inf on miss, units of |d|),
 * prim_id [R] u32 (0xFFFFFFFF on miss), prim_uv [R,2] (0 on miss), prim_normal [R,3] (unit
 * geometric normal e1 x e2 orientation of Open3D, 0 on miss; may be NULL). */
TEXIR_API int texir_cast_rays(const texir_scene* scene, const float* rays, int64_t R, float* t_hit,
                              uint32_t* prim_id, float* prim_uv, float* prim_normal, void* stream);
/* same, plus per-ray traversal counters (nodes visited, triangles tested) for profiling */
TEXIR_API int texir_cast_rays_stats(const texir_scene* scene, const float* rays, int64_t R, float* t_hit,
                                    uint32_t* prim_id, uint32_t* nodes_visited, uint32_t* tris_tested,
                                    void* stream);

/* ---- generate_dir: utils/sample_util.py:63-146 (pre_mode='Hammersley').
 * nrm [P,3], rough [P] (importance mode only, else NULL), shift [P,2] = the torch.rand(b,1,2)
 * Cranley-Patterson shift drawn by the caller; out [P,N,3]. */
TEXIR_API int texir_generate_dir(const float* nrm, const float* rough, const float* shift, int64_t P, int N,
                                 int mode, float* out, void* stream);
/* same with caller-drawn i.i.d. samples [P,N,2] (pre_mode != 'Hammersley', utils/sample_util.py:109-111:
 * `samples = torch.rand((b, N, 2))`, used as drawn -- no shift, wrap or clamp). */
TEXIR_API int texir_generate_dir_samples(const float* nrm, const float* rough, const float* samples, int64_t P,
                                         int N, int mode, float* out, void* stream);
/* gradient of generate_dir w.r.t. roughness (importance mode): g_out [P,N,3] -> g_rough [P] */
TEXIR_API int texir_generate_dir_bwd(const float* nrm, const float* rough, const float* shift, int64_t P, int N,
                                     const float* g_out, float* g_rough, void* stream);

/* diagnostics: s = sin(x), c = cos(x) for x [n] exactly as the kernels evaluate them inside generate_dir (variant 0) or
 * as two separate library calls (variant 1); used to pin the direction arithmetic to the reference's torch-CUDA run */
TEXIR_API int texir_debug_sincos(const float* x, int64_t n, int variant, float* s, float* c, void* stream);

/* ---- query_irf: models/mat_nvdiffrast.py:292-320 / models/tracer_o3d_irt.py:240-269.
 * org [P,3] (one origin per point, the reference expands it), dir [P,N,3]; rad [P,N,3]. */
TEXIR_API int texir_query_irf(const texir_scene* scene, const float* org, const float* dir, int64_t P, int N,
                              float* rad, void* stream);

/* ---- fused IrT batch: models/tracer_o3d_irt.py:165-173 for P points at once.
 * E[p] = sum_i L_i * clamp(n.l_i, 0, 1) * 2*pi / N  with l_i = generate_dir(nrm, N, mode).
 * order (nullable): a permutation of [0,P) from texir_point_order; warp w then works on point order[w].  The
 * result is independent of it (every point is computed by one warp from its own inputs); it only makes
 * neighbouring warps walk neighbouring parts of the BVH (measured +7 % on scattered points). */
TEXIR_API int texir_irradiance_fwd(const texir_scene* scene, const float* pos, const float* nrm,
                                   const float* shift, const int32_t* order, int64_t P, int N, int mode, float* E,
                                   void* stream);
/* The same launch for ONE RANK'S SHARE of a multi-GPU IrT (the reference's loop models/tracer_o3d_irt.py:156-173
 * has no multi-GPU form, the texels are independent).  pos / nrm / shift are the FULL [P_total,.] arrays, `order` lists
 * the P point indices this rank computes, and every result is stored to slot order[i] of ALL n_peers buffers in
 * E_peers_host (a HOST array of device pointers: this rank's own [P_total,3] buffer and the other ranks' buffers mapped
 * with texir_peer_open) -- the all-gather is the kernel's epilogue, over NVLink peer stores.  Follow it with
 * texir_peer_barrier before any rank reads its buffer. */
TEXIR_API int texir_irradiance_fwd_peers(const texir_scene* scene, const float* pos, const float* nrm,
                                         const float* shift, const int32_t* order, int64_t P, int N, int mode,
                                         float* const* E_peers_host, int n_peers, void* stream);
/* order [P] (int32, device): the points sorted along a 30-bit Morton curve over the scene box */
TEXIR_API int texir_point_order(const texir_scene* scene, const float* pos, int64_t P, int32_t* order, void* stream);

/* ---- fused radiance-field samples: the target generator of the NIrF / irradiance-field stages,
 *      models/tracer_o3d.py:73-116 and models/tracer_o3d_irrf.py:72-122: for every point the N
 *      shifted-Hammersley directions of generate_dir(nrm, N, mode) are traced and the hit radiance returned
 *      un-reduced.  rad [P,N,3]; dir [P,N,3] (nullable) receives the directions. */
TEXIR_API int texir_radiance_samples(const texir_scene* scene, const float* pos, const float* nrm, const float* shift,
                                     int64_t P, int N, int mode, float* rad, float* dir, void* stream);

/* ---- G-buffer from primary rays (replaces dr.rasterize + dr.interpolate of position / normal / uv,
 *      models/mat_nvdiffrast.py:119-126, models/tracer_o3d_irt.py:99-112).
 * rays [R,6].  Outputs (each nullable): pos [R,3] hit point, nrm [R,3] unit geometric normal turned against the
 * ray, uv [R,2] interpolated per-corner UVs (the reference's texc), prim [R] u32, bary [R,2], mask [R] f32
 * (1 = covered).  Uncovered pixels get pos = nrm = (1,0,0) like mat_nvdiffrast.py:125, uv = 0. */
TEXIR_API int texir_gbuffer(const texir_scene* scene, const float* rays, int64_t R, float* pos, float* nrm, float* uv,
                            uint32_t* prim, float* bary, float* mask, void* stream);

/* The same G-buffer straight from the reference's camera matrices, ONE launch for all views (replaces
 * `clip = vertices @ mvp; dr.rasterize; dr.interpolate x2`, models/mat_nvdiffrast.py:116-128): mvp [V,4,4] in the
 * reference's row-vector convention (clip = [v,1] @ mvp[i]; datasets/dataset.py:461-463), cam [3], res x res pixels
 * per view, pixel row 0 at ndc y = -1.  Outputs as texir_gbuffer, [V*res*res, .], plus uv_da [.,4] = the reference's
 * texd (du/dX, du/dY, dv/dX, dv/dY) and rays_out [.,6] (origin, unit direction); each nullable. */
TEXIR_API int texir_gbuffer_mvp(const texir_scene* scene, const float* mvp, const float* cam, int V, int res, float* pos,
                                float* nrm, float* uv, float* uv_da, uint32_t* prim, float* bary, float* mask,
                                float* rays_out, void* stream);
/* The same for one rank of a pixel-sharded Mat stage, and with the reference's texture-coordinate convention:
 * rank `shard_rank` of `shard_world` takes the 128-pixel blocks b = shard_rank (mod shard_world) of every view and
 * stores them compactly (outputs [V*res*res/shard_world, .]; res*res must be a multiple of 128*shard_world).
 * flip_v != 0: uv = (u, 1 - v) and dv/dX, dv/dY negated -- pyredner.load_obj's flip_tex_coords=True default, which
 * is what the reference's self.uvs hold (models/mat_nvdiffrast.py:194) and its dr.texture calls address with. */
TEXIR_API int texir_gbuffer_mvp_shard(const texir_scene* scene, const float* mvp, const float* cam, int V, int res,
                                      int shard_rank, int shard_world, int flip_v, float* pos, float* nrm, float* uv,
                                      float* uv_da, uint32_t* prim, float* bary, float* mask, float* rays_out,
                                      void* stream);
/* screen-space derivatives of the interpolated UVs (the reference's texd = dr.interpolate(..., rast_db,
 * diff_attrs='all'), models/mat_nvdiffrast.py:128), as a ray differential on the hit triangle's plane.
 * ddir [R,6] = (d dir/dX, d dir/dY) per pixel, prim [R] from texir_gbuffer -> uv_da [R,4] =
 * (du/dX, du/dY, dv/dX, dv/dY); uncovered pixels get 0. */
TEXIR_API int texir_gbuffer_uv_da(const texir_scene* scene, const float* rays, const float* ddir, const uint32_t* prim,
                                  int64_t R, float* uv_da, void* stream);

/* ---- fused render: MaterialModel.render, models/mat_nvdiffrast.py:201-249 (and the relighting
 *      variant models/test_nvdiffrast.py:256-303 when N_diff > 0).
 * nrm, albedo, pos, irr [P,3]; rough [P]; cam [3]; shift_spec [P,2]; shift_diff [P,2] (N_diff > 0).
 * eps = 1e-14 (mat_nvdiffrast.py) or 1e-6 (test_nvdiffrast.py and the baseline variants).
 * rgb [P,3].  Optional stashes that make the backward pass free of ray casts:
 *   drgb_drough [P,3] = d rgb / d roughness,  ediff [P,3] = MC diffuse irradiance term (N_diff > 0). */
TEXIR_API int texir_shade_fwd(const texir_scene* scene, const float* nrm, const float* albedo, const float* rough,
                              const float* pos, const float* cam, const float* irr, const float* shift_spec,
                              const float* shift_diff, int64_t P, int N_spec, int N_diff, int mode_spec,
                              int mode_diff, float eps, float* rgb, float* drgb_drough, float* ediff,
                              void* stream);
/* backward of texir_shade_fwd.  g_rgb [P,3] -> g_albedo [P,3], g_rough [P], g_irr [P,3] (each may
 * be NULL).  If drgb_drough (and ediff when N_diff > 0) from the forward call are given they are
 * used; otherwise the kernel re-generates the samples and re-casts every ray. */
TEXIR_API int texir_shade_bwd(const texir_scene* scene, const float* nrm, const float* albedo, const float* rough,
                              const float* pos, const float* cam, const float* irr, const float* shift_spec,
                              const float* shift_diff, int64_t P, int N_spec, int N_diff, int mode_spec,
                              int mode_diff, float eps, const float* g_rgb, const float* drgb_drough,
                              const float* ediff, float* g_albedo, float* g_rough, float* g_irr, void* stream);

/* The same backward, accumulated straight into TEXTURE SPACE by the same kernel: instead of per-pixel g_albedo /
 * g_rough, every pixel's gradient is scattered with atomics (red.global.add.f32) into the gradient textures of
 * materials_a [Ha,Wa,3] and materials_r [Hr,Wr,1] at its uv [P,2] -- the transpose of the dr.texture fetch of
 * models/mat_nvdiffrast.py:131-139 ('linear-mipmap-linear' with uv_da [P,4] and La / Lr pyramid levels in g_mips_*,
 * plain bilinear when uv_da is NULL and La = Lr = 0).  Gradient textures and pyramids are zero-initialised by the
 * caller; texir_texture_mip_fold then folds each pyramid into its level 0.  Either texture may be NULL. */
TEXIR_API int texir_shade_bwd_tex(const texir_scene* scene, const float* nrm, const float* albedo, const float* rough,
                                  const float* pos, const float* cam, const float* irr, const float* shift_spec,
                                  const float* shift_diff, int64_t P, int N_spec, int N_diff, int mode_spec,
                                  int mode_diff, float eps, const float* g_rgb, const float* drgb_drough,
                                  const float* ediff, const float* uv, const float* uv_da, float* g_tex_a,
                                  float* g_mips_a, int Ha, int Wa, int La, float* g_tex_r, float* g_mips_r, int Hr,
                                  int Wr, int Lr, void* stream);

/* The whole backward of a Mat-stage step (models/mat_nvdiffrast.py:131-139,179-182 + the reference loss's direct terms
 * on preds['albedo'] / ['roughness'] / ['roughness_womipmap'], models/loss.py:94,103,113) in ONE launch: as
 * texir_shade_bwd_tex, plus per-pixel gradients of the fetched materials added before the scatter -- add_a [P,3] and
 * add_r [P] through the mip-mapped transpose, add_rw [P] through the bilinear one (each nullable).  render_albedo = 0:
 * the render saw albedo.detach() (stage 1); render_rough = 0: roughness entered detached; render_rough_bilinear = 1: the
 * render saw roughness_womipmap (stage 1), so its roughness gradient takes the bilinear transpose. */
TEXIR_API int texir_shade_bwd_mat(const texir_scene* scene, const float* nrm, const float* albedo, const float* rough,
                                  const float* pos, const float* cam, const float* irr, const float* shift_spec,
                                  const float* shift_diff, int64_t P, int N_spec, int N_diff, int mode_spec,
                                  int mode_diff, float eps, const float* g_rgb, const float* drgb_drough,
                                  const float* ediff, const float* uv, const float* uv_da, const float* add_a,
                                  const float* add_r, const float* add_rw, int render_albedo, int render_rough,
                                  int render_rough_bilinear, float* g_tex_a, float* g_mips_a, int Ha, int Wa, int La,
                                  float* g_tex_r, float* g_mips_r, int Hr, int Wr, int Lr, void* stream);

/* ---- stand-alone BRDF reductions with the reference's tensor signatures --------------------
 * specular_reflectance(lighting, h, n, v, l, roughness): mat_nvdiffrast.py:260-279.
 * lighting, h, l [P,N,3]; n, v [P,3]; rough [P]; out [P,3] (NOT divided by N, as in the reference). */
TEXIR_API int texir_specular_reflectance(const float* lighting, const float* h, const float* n, const float* v,
                                         const float* l, const float* rough, int64_t P, int N, float eps,
                                         float* out, void* stream);
/* g_out [P,3] -> g_lighting, g_h, g_l [P,N,3], g_rough [P] (each may be NULL) */
TEXIR_API int texir_specular_reflectance_bwd(const float* lighting, const float* h, const float* n, const float* v,
                                             const float* l, const float* rough, int64_t P, int N, float eps,
                                             const float* g_out, float* g_lighting, float* g_h, float* g_l,
                                             float* g_rough, void* stream);
/* diffuse_reflectance(lighting, l, n, albedo, sample_type): mat_nvdiffrast.py:252-258; cosine != 0
 * selects the 'cosine' branch. */
TEXIR_API int texir_diffuse_reflectance(const float* lighting, const float* l, const float* n, const float* albedo,
                                        int64_t P, int N, int cosine, float* out, void* stream);
TEXIR_API int texir_diffuse_reflectance_bwd(const float* lighting, const float* l, const float* n,
                                            const float* albedo, int64_t P, int N, int cosine, const float* g_out,
                                            float* g_lighting, float* g_l, float* g_albedo, void* stream);

/* ---- texture-space material fetch / gradient scatter (dr.texture(..., filter_mode='linear'),
 *      models/mat_nvdiffrast.py:131-139) ------------------------------------
 * tex [H,W,C] f32, uv [P,2] in [0,1] (nvdiffrast convention: texel centres at (i+0.5)/W, wrap
 * boundary mode), out [P,C]. */
TEXIR_API int texir_texture_fetch(const float* tex, int H, int W, int C, const float* uv, int64_t P, float* out,
                                  void* stream);
/* g_out [P,C] scattered with atomics into g_tex [H,W,C] (caller zero-initialises) */
TEXIR_API int texir_texture_scatter(const float* g_out, int H, int W, int C, const float* uv, int64_t P,
                                    float* g_tex, void* stream);

/* ---- mip-mapped fetch: dr.texture(tex[None], texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=M),
 *      models/mat_nvdiffrast.py:131-139 (albedo, roughness, irr).  Rules restated in csrc/texir_tex.h.
 * A pyramid is the caller's level-0 texture [H,W,C] plus `mips`: levels 1..L packed back to back ([h,w,C] each,
 * h = max(H >> l, 1)).  texir_texture_mip_levels: L for a texture and a max_mip_level cap (< 0: none), -1 when an odd
 * side would have to be halved (nvdiffrast raises there too).  texir_texture_mip_elems: floats in `mips`. */
TEXIR_API int texir_texture_mip_levels(int H, int W, int max_mip_level);
TEXIR_API int64_t texir_texture_mip_elems(int H, int W, int C, int levels);
/* 2x2 box averages, level by level */
TEXIR_API int texir_texture_mip_build(const float* tex, int H, int W, int C, int levels, float* mips, void* stream);
/* uv [P,2] in [0,1] (wrap), uv_da [P,4] = (du/dX, du/dY, dv/dX, dv/dY) per screen pixel -> out [P,C];
 * lod_out [P] (nullable) receives the fractional level of detail */
TEXIR_API int texir_texture_fetch_mip(const float* tex, const float* mips, int H, int W, int C, int levels,
                                      const float* uv, const float* uv_da, int64_t P, float* out, float* lod_out,
                                      void* stream);
/* transpose of the fetch: g_out [P,C] scattered with atomics into g_tex / g_mips (caller zero-initialises both) ... */
TEXIR_API int texir_texture_scatter_mip(const float* g_out, int H, int W, int C, int levels, const float* uv,
                                        const float* uv_da, int64_t P, float* g_tex, float* g_mips, void* stream);
/* ... and transpose of the build: g_mips folded down into g_tex, coarsest level first (g_mips is consumed) */
TEXIR_API int texir_texture_mip_fold(float* g_tex, float* g_mips, int H, int W, int C, int levels, void* stream);

/* The per-step forms of the three calls above (same values, fewer passes; csrc/texir_matstage.cu): the whole pyramid
 * in two launches; the fold as one gather launch that writes the folded level 0 to `out` (out == g_tex: in place) and,
 * with clear != 0, leaves g_tex / g_mips zeroed for the next step's scatter; and the four per-pixel fetches of a Mat-stage
 * step (albedo and irradiance [P,3] mip-mapped, roughness [P] mip-mapped and bilinear; models/mat_nvdiffrast.py:131-139)
 * in one launch, each output nullable. */
TEXIR_API int texir_texture_mip_build_fast(const float* tex, int H, int W, int C, int levels, float* mips, void* stream);
TEXIR_API int texir_texture_mip_fold_into(float* g_tex, float* g_mips, int H, int W, int C, int levels, float* out,
                                          int clear, void* stream);
TEXIR_API int texir_mat_fetch(const float* tex_a, const float* mips_a, int Ha, int Wa, int La, const float* tex_r,
                              const float* mips_r, int Hr, int Wr, int Lr, const float* tex_i, const float* mips_i, int Hi,
                              int Wi, int Li, const float* uv, const float* uv_da, int64_t P, float* albedo, float* rough,
                              float* rough_womip, float* irr, void* stream);

/* The texture-space half of the Mat-stage backward on its own: P rows of 11 floats = (u, v, du/dX, du/dY, dv/dX, dv/dY,
 * g_albedo[3], g_roughness taken through the mip-mapped fetch, g_roughness taken through the bilinear fetch) scattered
 * with atomics into the gradient pyramids (zero-initialised by the caller; fold afterwards).  Used when the per-pixel
 * gradients of all ranks were all-gathered (4.3 MB at the reference's 98 304 pixels) instead of all-reducing the
 * 112 MiB texture gradients: every rank then scatters the same rows. */
TEXIR_API int texir_mat_scatter_rows(const float* rows, int64_t P, float* g_tex_a, float* g_mips_a, int Ha, int Wa, int La,
                                     float* g_tex_r, float* g_mips_r, int Hr, int Wr, int Lr, void* stream);

/* ---- the reference's shift stream on the device: n values of torch.rand (float32, CPU default generator:
 * at::mt19937 + uniform_real_distribution<float>, utils/sample_util.py:102) from the generator state
 * (state_in: the 624 state words; left / next: the engine's counters, all from torch.get_rng_state()), bit for bit, and
 * the advanced state (state_out: 624 words, then left, next) for torch.set_rng_state.  One CTA; device pointers.
 * left == -1: the counters are read from state_in[624], state_in[625] instead (the form a CUDA graph replays: kernel
 * arguments are frozen at capture); an inconsistent pair yields zeros and state_out[624] = state_out[625] = 0xFFFFFFFF. */
TEXIR_API int texir_mt19937_uniform(const uint32_t* state_in, int left, int next, int64_t n, float* out,
                                    uint32_t* state_out, void* stream);
/* the engine bookkeeping of the same n draws on the HOST (no numbers produced): state[624], *left, *next advanced in
 * place to where torch.rand(n) leaves the generator -- lets the caller hand torch its state without a device round trip. */
TEXIR_API int texir_mt19937_advance_host(uint32_t* state, int* left, int* next, int64_t n);

/* ---- peer memory of the multi-GPU path: one process per GPU; a rank exports a device buffer as a 64-byte
 *      handle (CUDA IPC), the other ranks map it over NVLink / NVSwitch and kernels store into it directly.
 * texir_peer_alloc: zeroed buffer on `device` + its handle (handle_out: TEXIR_PEER_HANDLE_BYTES, HOST memory);
 * texir_peer_open: map another rank's buffer; _close / _free undo them.
 * texir_peer_barrier: flags_host[r] = rank r's flag array (world + 1 uint32, from texir_peer_alloc, own or mapped):
 * every rank stores `epoch` (monotonically increasing, > 0) into its slot of every array and waits until its own array
 * holds `epoch` from everybody -- stream-ordered, so stores of earlier kernels on `stream` are visible to every rank
 * after it.  The wait is bounded (about 4 s); on expiry slot [world] of the own array is set to `epoch`. */
TEXIR_API int texir_peer_alloc(int device, int64_t bytes, void** dptr, unsigned char* handle_out);
TEXIR_API int texir_peer_free(int device, void* dptr);
TEXIR_API int texir_peer_open(int device, const unsigned char* handle, void** dptr);
TEXIR_API int texir_peer_close(int device, void* dptr);
TEXIR_API int texir_peer_barrier(uint32_t* const* flags_host, int rank, int world, uint32_t epoch, int device,
                                 void* stream);

/* ---- IrT post-processing ---------------------------------------------
 * padding_texture (tools/padding_texture.py:49-80): every empty texel (channel sum == 0) takes the value of
 * its nearest non-empty texel -- scipy.ndimage.distance_transform_edt(return_indices=True) followed by
 * F.grid_sample(mode='nearest'), including both of their tie / rounding rules (csrc/texir_post.h).
 * img, out [H,W,C] f32 (C <= 4, out != img); index [H,W,2] int32 (row, col of the nearest non-empty texel,
 * scipy's `indices`; may be NULL); scratch [H,W] int32. */
TEXIR_API int texir_pad_texture(const float* img, int H, int W, int C, float* out, int32_t* index, int32_t* scratch,
                                void* stream);
/* Radiance RGBE pixel codec of cv2.imwrite / cv2.imread for .hdr (trainer/generate_ir_texture.py:80-82,
 * models/mat_nvdiffrast.py:73,89; OpenCV rgbe.cpp float2rgbe / rgbe2float).  img [n,3] f32 in RGB order, or in
 * cv2's BGR order when bgr != 0; rgbe [n,4] bytes in file order (R, G, B mantissas, exponent). */
TEXIR_API int texir_rgbe_encode(const float* img, int64_t n, int bgr, uint8_t* rgbe, void* stream);
TEXIR_API int texir_rgbe_decode(const uint8_t* rgbe, int64_t n, int bgr, float* img, void* stream);
/* the .hdr container around it, HOST memory and file I/O only (no GPU work): header "#?RADIANCE / FORMAT=32-bit_rle_rgbe /
 * -Y H +X W" and OpenCV's per-scanline run-length coding; byte-identical to cv2.imwrite for the same pixels. */
TEXIR_API int texir_hdr_write_host(const char* path, const uint8_t* rgbe_host /*[H,W,4]*/, int H, int W);
TEXIR_API int texir_hdr_read_info(const char* path, int* H, int* W);
TEXIR_API int texir_hdr_read_host(const char* path, uint8_t* rgbe_host /*[H,W,4]*/, int H, int W);
/* the 16-bit PNG index texture (tools/padding_texture.py:8-43 reads and writes it through cv2): undo the PNG scanline
 * filters of an inflated, non-interlaced IDAT stream.  filtered = H rows of 1 + row_bytes bytes, out = H rows of
 * row_bytes, bpp = bytes per pixel.  HOST memory only; chunk parsing and zlib stay with the caller. */
TEXIR_API int texir_png_unfilter_host(const uint8_t* filtered, int H, int row_bytes, int bpp, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* TEXIR_B200_H */
