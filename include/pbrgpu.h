/* pbrgpu — C ABI of the B200 (sm_100a) backend for pbrlab's path-tracing hot path.
 *
 * This is the drop-in boundary: plain C, opaque handle, plain pointers and sizes, int error codes, no exceptions
 * and no C++/torch types.  It is what a reference-side binding would call from
 *   - Scene::CommitScene()            (reference src/scene.cc:96-104)      -> pbrgpu_set_* + pbrgpu_commit
 *   - pbrlab::Render()                (reference src/render.h:14-17, src/render.cc:192-241) -> pbrgpu_render
 *   - Scene::TraceFirstHit1 / AnyHit1 (reference src/scene.h:89-91, src/scene.cc:261-268)   -> pbrgpu_trace / _occluded
 * (INTEGRATION.md shows the stub).  Our own C++ mirror of the reference API (pbrlab_b200/host/) and the Python
 * tests / bench both go through exactly these entry points.
 *
 * Threading: a context is thread-compatible, not thread-safe.  All calls are blocking.
 * Every function returning int returns PBRGPU_OK (0) on success; on failure pbrgpu_last_error() describes it.
 * There is NO CPU fallback: without a CUDA device pbrgpu_create() fails. */
#ifndef PBRGPU_H_
#define PBRGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBRGPU_OK 0
#define PBRGPU_ERR_INVALID 1   /* bad argument / call order */
#define PBRGPU_ERR_CUDA 2      /* CUDA runtime error (no device, out of memory, launch failure) */
#define PBRGPU_ERR_BUILD 3     /* acceleration structure could not be built */

#define PBRGPU_INVALID_ID 0xFFFFFFFFu

typedef struct pbrgpu_ctx pbrgpu_ctx;

/* Material parameter block.  type 0 mirrors CyclesPrincipledBsdfParameter (reference src/material-param.h:24-49):
 *   p[0..2] base_color, p[3] subsurface, p[4..6] subsurface_radius, p[7..9] subsurface_color, p[10] metallic,
 *   p[11] specular, p[12] specular_tint, p[13] roughness, p[14] anisotropic, p[15] anisotropic_rotation,
 *   p[16] sheen, p[17] sheen_tint, p[18] clearcoat, p[19] clearcoat_roughness, p[20] ior, p[21] transmission,
 *   p[22] transmission_roughness; tex_id = {base_color_tex_id, subsurface_color_tex_id}.
 * type 1 mirrors HairBsdfParameter (src/material-param.h:51-72):
 *   p[0] coloring_hair (0 = kRGB, 1 = kMelanin), p[1..3] base_color, p[4] melanin, p[5] melanin_redness,
 *   p[6] melanin_randomize, p[7] roughness, p[8] azimuthal_roughness, p[9] ior, p[10] shift (degrees),
 *   p[11..13] specular_tint, p[14..16] second_specular_tint, p[17..19] transmission_tint. */
typedef struct pbrgpu_material {
  uint32_t type;
  uint32_t tex_id[2];
  uint32_t reserved;
  float p[24];
} pbrgpu_material;

/* One ray, the layout of the SoA ray queues: origin + tnear, direction + tfar (reference src/ray.h:9-14). */
typedef struct pbrgpu_ray {
  float org[3];
  float tmin;
  float dir[3];
  float tmax;
} pbrgpu_ray;

/* One closest-hit result: the reference's TraceResult (src/raytracer/raytracer.h:9-17).
 * Miss <=> instance_id == PBRGPU_INVALID_ID. */
typedef struct pbrgpu_hit {
  float normal_g[3]; /* normalised geometric normal (triangles) / curve tangent (curves) */
  float t, u, v;
  uint32_t instance_id, geom_id, prim_id;
} pbrgpu_hit;

/* One texture of Scene::AddTexture (reference src/scene.h:45-51, src/texture.h:10-44): row-major float pixels with
 * `channels` (1..4) interleaved values per texel, already in linear colour (the loader de-gammas sRGB files,
 * reference src/io/triangle-mesh-io.cc:80-104).  Sampled as Texture::FetchFloat3 does: bilinear, clamp addressing
 * (src/texture.cc:43-72, src/image-utils.cc:99-167). */
typedef struct pbrgpu_texture {
  const float* pixels;
  uint32_t width, height, channels;
  uint32_t reserved;
} pbrgpu_texture;

/* Flattened area-light tables, exactly the numbers LightManager holds after Commit()
 * (reference src/light-manager.cc:29-184): one entry per light in global light order, and per light one entry per
 * primitive of its mesh. */
typedef struct pbrgpu_light_tables {
  uint32_t num_lights;
  const float* light_probability;   /* [num_lights] choose_light_probability */
  const float* light_cdf;           /* [num_lights] cumulative_probability_ */
  const uint32_t* light_prim_offset;/* [num_lights + 1] offsets into the per-primitive arrays */
  uint32_t num_light_prims;
  const float* prim_probability;    /* [num_light_prims] choose_primitive_probability */
  const float* prim_cdf;            /* [num_light_prims] AreaLight::cumulative_probability_ */
  const float* prim_area_pdf;       /* [num_light_prims] prim_area_measure_pdf (0 for non-emissive) */
  const float* prim_emission;       /* [num_light_prims * 3] AreaLightParameter::emission (0 for non-emissive) */
  const uint32_t* prim_is_emissive; /* [num_light_prims] light_param_ids[prim] != -1 */
  const uint32_t* prim_triangle;    /* [num_light_prims] index of the triangle in pbrgpu_set_triangles order */
} pbrgpu_light_tables;

typedef struct pbrgpu_stats {
  uint64_t paths;          /* camera samples this process accumulated in the last pbrgpu_render* call (its share of
                              the job; fewer after a cancel) / paths traced by the last hook call */
  uint64_t closest_rays;   /* closest-hit rays (path vertices) */
  uint64_t shadow_rays;    /* any-hit rays */
  uint64_t sss_rays;       /* closest-hit rays issued inside random-walk subsurface scattering */
  uint64_t kernel_launches;
  double   seconds;        /* wall time of the call, device work included */
  double   trace_closest_ms, trace_any_ms, shade_ms, sss_ms;  /* CUDA-event sums per kernel family (profiling mode) */
  uint64_t nodes_visited, prims_tested;   /* only filled by pbrgpu_trace* with stats enabled */
  double   regen_ms;       /* retire + regenerate kernel (profiling mode) */
  double   device_ms;      /* CUDA-event time from the first to the last kernel of the call on the launching stream
                              (max over the context's devices) */
  uint64_t trace_closest_launches;
  uint64_t sss_skipped;    /* random-walk segments answered by the clearance grid instead of a ray query */
  uint64_t shade_vertices; /* path vertices shaded (Shader() calls incl. the exit vertex of a random walk) */
  uint64_t iterations;     /* wavefront iterations (= launches of each kernel family) of the last render */
} pbrgpu_stats;

/* ---- life cycle.  device_ids == NULL / n_devices == 0: the current device only. */
pbrgpu_ctx* pbrgpu_create(const int* device_ids, int n_devices);
void pbrgpu_destroy(pbrgpu_ctx* ctx);
const char* pbrgpu_last_error(const pbrgpu_ctx* ctx);   /* ctx may be NULL: error of the last failed create */
int pbrgpu_device_count(void);

/* ---- scene upload (host pointers; data is copied).  All instances are flattened by the caller into one triangle
 * soup and one curve soup; the (instance_id, geom_id, prim_id) triple the reference reports is kept per primitive.
 *   xyzw      [nverts*4]   positions, stride 16 B as the reference hands them to Embree (raytracer_impl.cc:124-127)
 *   vidx      [ntris*3]    vertex indices
 *   nxyzw     [nnormals*4] shading normals (may be NULL), nidx [ntris*3] normal indices, 0xFFFFFFFF = none
 *   uv        [nuv*2]      texcoords (may be NULL), tidx [ntris*3], 0xFFFFFFFF = none
 *   material_id / instance_id / geom_id / prim_id  [ntris]  */
int pbrgpu_set_triangles(pbrgpu_ctx* ctx, const float* xyzw, uint32_t nverts, const uint32_t* vidx,
                         const float* nxyzw, uint32_t nnormals, const uint32_t* nidx, const float* uv, uint32_t nuv,
                         const uint32_t* tidx, const uint32_t* material_id, const uint32_t* instance_id,
                         const uint32_t* geom_id, const uint32_t* prim_id, uint64_t ntris);
/*   xyzr      [nverts*4]   Bezier control points: position + radius (CyHair thickness passed through unchanged,
 *                          reference src/io/curve-mesh-io.cc:107-110)
 *   first_cp  [nsegs]      index of the first of 4 consecutive control points of each cubic segment */
int pbrgpu_set_curves(pbrgpu_ctx* ctx, const float* xyzr, uint32_t nverts, const uint32_t* first_cp,
                      const uint32_t* material_id, const uint32_t* instance_id, const uint32_t* geom_id,
                      const uint32_t* prim_id, uint64_t nsegs);
/* May be called again between renders (live material edits, reference pc/pbrlab-gui.cc:207-238). */
int pbrgpu_set_materials(pbrgpu_ctx* ctx, const pbrgpu_material* materials, uint32_t n);
int pbrgpu_set_lights(pbrgpu_ctx* ctx, const pbrgpu_light_tables* tables);
/* Texture table the materials' tex_id index (pixels are copied).  Call before pbrgpu_commit; a material whose
 * tex_id is neither PBRGPU_INVALID_ID nor inside this table fails the commit. */
int pbrgpu_set_textures(pbrgpu_ctx* ctx, const pbrgpu_texture* textures, uint32_t n);
/* Builds the acceleration structure and uploads everything to every device of the context.  bmin/bmax: the scene
 * AABB the camera is derived from (reference src/render.cc:132-158); pass NULL to use the bounds of the uploaded
 * geometry (triangle vertices in use; curve bounds as Embree's accurateFlatBounds). */
int pbrgpu_commit(pbrgpu_ctx* ctx, const float* bmin, const float* bmax);
int pbrgpu_scene_bounds(const pbrgpu_ctx* ctx, float* bmin, float* bmax);
/* What the last pbrgpu_commit() did: wall seconds in total and of its parts (acceleration-structure build of both BVHs
 * incl. transfers, clearance field of the subsurface meshes, upload of all tables to the devices), and which builder
 * made the triangle BVH: 0 = host binned SAH, 1 = PLOC on the host, 2 = PLOC on the device. */
typedef struct pbrgpu_commit_info {
  double commit_s, bvh_s, clearance_s, upload_s;
  uint32_t tri_builder, tri_nodes, curve_nodes, tri_depth;
} pbrgpu_commit_info;
int pbrgpu_get_commit_info(const pbrgpu_ctx* ctx, pbrgpu_commit_info* out);

/* ---- rendering: the body of pbrlab::Render().  Blocking.  Accumulates SUMS exactly like RenderLayer:
 *   rgba_out [w*h*4] += (L.r, L.g, L.b, 1) per sample, count_out [w*h] += 1 per sample (both are overwritten,
 *   i.e. cleared first, as PrepareRendering does).  `seed` selects the per-path PCG32 streams:
 *   path (pixel p, sample s) uses pcg32_srandom(initstate = seed + s, initseq = p).
 *   sample_offset/sample_stride: this call renders samples s = sample_offset + k*sample_stride < spp (normally 0, 1;
 *   a launcher that does its own reduction gives rank r of R the pair (r, R); see pbrgpu_nccl_init for the built-in
 *   multi-process split).
 *   cancel may be NULL; polled once per wavefront iteration.  When it is raised the call stops handing out samples,
 *   drops the paths still in flight and returns PBRGPU_OK with what was accumulated so far: every pixel holds whole
 *   samples only (rgba.a == count), as a cancelled reference render does (src/render.cc:217-231).
 *   finish_pass (may be NULL) is raised monotonically and never exceeds spp. */
int pbrgpu_render(pbrgpu_ctx* ctx, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed,
                  uint32_t sample_offset, uint32_t sample_stride, const volatile int* cancel, float* rgba_out,
                  uint32_t* count_out, size_t* finish_pass);
/* Same, but the result stays on the device: d_rgba / d_count are DEVICE pointers on the context's first device
 * (e.g. a tensor's data_ptr) and nothing is copied to the host. */
int pbrgpu_render_device(pbrgpu_ctx* ctx, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed,
                         uint32_t sample_offset, uint32_t sample_stride, const volatile int* cancel, float* d_rgba,
                         uint32_t* d_count, size_t* finish_pass);
/* ---- one frame split over several processes (one process per GPU, e.g. torchrun; north_star: "a frame is
 * partitioned across the 8 GPUs ... the accumulation buffers summed by a single NCCL reduce").  Rank 0 calls
 * pbrgpu_nccl_unique_id() and hands the 128 bytes to every rank by whatever means the launcher has; then EVERY rank
 * calls pbrgpu_nccl_init() (collective).  From then on pbrgpu_render / pbrgpu_render_device are collective: rank r of
 * R renders the samples sample_offset + (r + k*R)*sample_stride of the job — the reference's (tile, sample) job split,
 * src/render.cc:211-222, with the sample as the unit — and the sums are added onto rank 0 by ONE ncclReduce over
 * NVLink at frame end (16 B per pixel: count is the alpha sum).  Rank 0 receives the complete frame; the other ranks'
 * buffers hold their own partial sums.  A context that owns several devices (pbrgpu_create with n_devices > 1) splits
 * and reduces the same way inside one process (ncclCommInitAll).  NCCL is loaded on first use (dlopen), the library
 * does not link against it. */
#define PBRGPU_NCCL_ID_BYTES 128
int pbrgpu_nccl_unique_id(uint8_t* id128);
int pbrgpu_nccl_init(pbrgpu_ctx* ctx, const uint8_t* id128, int rank, int world);
int pbrgpu_job_rank(const pbrgpu_ctx* ctx, int* rank, int* world);

/* Output stage of the reference CLI on the device (pc/pbrlab-cli.cc:47-57, src/image-utils.cc:26-38,72-90,
 * src/io/image-io.cc:172-210): colour = rgba / count -> LinerToSrgb on r,g,b (alpha untouched) -> 8 bit as
 * WritePNG quantises, (unsigned char)clamp(v * 256, 0, 255).  Reads the accumulators the last pbrgpu_render* call
 * left on the context's first device (for pbrgpu_render_device: the caller's d_rgba / d_count).  rgba8_out is a HOST
 * buffer of width*height*4 bytes; pbrgpu_resolve_srgb8_device writes a DEVICE buffer instead. */
int pbrgpu_resolve_srgb8(pbrgpu_ctx* ctx, uint32_t width, uint32_t height, uint8_t* rgba8_out);
int pbrgpu_resolve_srgb8_device(pbrgpu_ctx* ctx, const float* d_rgba, const uint32_t* d_count, uint32_t width,
                                uint32_t height, uint8_t* d_rgba8_out);
int pbrgpu_get_stats(const pbrgpu_ctx* ctx, pbrgpu_stats* out);
/* profiling mode: every kernel family of every iteration is bracketed by CUDA events on the launching stream and the
 * sums are reported in pbrgpu_stats (adds ~10 event records per iteration) */
int pbrgpu_set_profiling(pbrgpu_ctx* ctx, int enabled);
/* samples of one pixel traced concurrently per wave (0 = choose from free memory) */
int pbrgpu_set_wave_spp(pbrgpu_ctx* ctx, uint32_t spp_per_wave);

/* Measurement hook for the roofline of the L2-resident scenes ("L2 bandwidth is not in
 * MEASURED_PEAKS.json — measure it with a resident-working-set gather microbenchmark"): every thread of a persistent
 * grid fetches node-sized records (80 bytes = five 128-bit loads, the traversal's access unit) at pseudo-random
 * positions of a working set of `working_set_bytes` (resident in L2 when it is a few tens of MB, in HBM when it is
 * GBs), `records_per_thread` times, with `chains` independent dependent-load chains per thread (1 = latency bound like
 * a single traversal, 8 = bandwidth bound).  Reports the achieved GB/s of the best of 5 launches (CUDA events). */
int pbrgpu_measure_gather(pbrgpu_ctx* ctx, uint64_t working_set_bytes, uint32_t records_per_thread, uint32_t chains,
                          double* gbytes_per_s);

/* ---- test hooks for the parity gates (host pointers) */
/* closest hit / occlusion for a batch of rays: Scene::TraceFirstHit1 / AnyHit1 semantics */
int pbrgpu_trace(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, uint64_t n, pbrgpu_hit* hits);
int pbrgpu_occluded(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, uint64_t n, uint8_t* occluded);
/* same on DEVICE buffers (timing the traversal kernels alone); hits/occluded may be NULL to discard results */
int pbrgpu_trace_device(pbrgpu_ctx* ctx, const pbrgpu_ray* d_rays, uint64_t n, pbrgpu_hit* d_hits, int collect_stats);
int pbrgpu_occluded_device(pbrgpu_ctx* ctx, const pbrgpu_ray* d_rays, uint64_t n, uint8_t* d_occluded);
/* GetRadiance (reference src/render.cc:24-90) for caller-supplied rays and PCG32 seeds:
 * seeds[2*i] = initstate, seeds[2*i+1] = initseq; radiance_out [n*3]. */
int pbrgpu_radiance(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* radiance_out);
/* same paths through a one-thread-per-path megakernel that traces its own shadow rays in the reference's order: the
 * cross-check for the wavefront scheduling (both run the same per-vertex device functions) */
int pbrgpu_radiance_mega(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n,
                         float* radiance_out);
/* one shading vertex: closest hit -> Shader() (reference src/shader/shader.cc:8-34), no emission / roulette.
 * out [n*16]: hit flag, wi[3], throughput[3], contribute[3] (shadow test applied), pdf, position[3],
 * face_direction, t */
int pbrgpu_shade(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* out16);
/* closure-level known-answer hooks; `op` and the argument packing are listed in pbrlab_b200/csrc/kat.cuh.
 * params: 32 floats (unused tail ignored); in: n records of in_stride floats; out: n records of out_stride floats */
int pbrgpu_eval_closure(pbrgpu_ctx* ctx, int op, const float* params, const float* in, uint32_t in_stride, uint64_t n,
                        float* out, uint32_t out_stride);

#ifdef __cplusplus
}
#endif
#endif /* PBRGPU_H_ */
