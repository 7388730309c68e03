/*
 * rt_b200.h -- C ABI of the B200-native renderer core (librt_b200.so).
 *
 * This is the drop-in boundary for the render hot path of ACEfanatic02/par_raytracer.
 * The reference has no plugin / FFI layer; the seam is the set of static C++ functions
 *
 *     Render(Camera*, Scene*, u32 w, u32 h) -> Framebuffer      main.cpp:301-358
 *     RenderTask(RenderJob*, DebugCounters*)                    main.cpp:267-283
 *     RenderPixel(RenderJob*, DebugCounters*, x, y) -> Vector4  main.cpp:224-265
 *     TraceRayColor(Ray, Scene*, iters, dbg, rng) -> Vector4    raytracer.cpp:413-577
 *     TraceRay(Ray, Scene*, RaycastHit*, dbg) -> bool           raytracer.cpp:159-232
 *
 * Each entry point below names the reference function it replaces. Plain pointers and
 * sizes only; no C++ / torch types. All structs are POD with the field order of the
 * reference struct they mirror so the host shim is a field copy.
 *
 * There is NO CPU fallback behind this interface: every compute entry point runs CUDA
 * kernels on the device the scene was created on and fails with RT_ERR_CUDA otherwise.
 */
#ifndef RT_B200_H_
#define RT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 1

/* status codes (reference has no error convention: brt.h:38-41 asserts print and continue) */
#define RT_OK          0
#define RT_ERR_ARG    -1
#define RT_ERR_CUDA   -2
#define RT_ERR_NOMEM  -3
#define RT_ERR_STATE  -4

/* ---- geometry.h:4-12 ------------------------------------------------------------ */
typedef struct rt_ray {
    float origin[3];
    float direction[3];
} rt_ray;                                   /* == Ray, 24 B */

/* ---- main.cpp:133-143 ----------------------------------------------------------- */
typedef struct rt_camera {
    float tan_a2;
    float aspect;
    float inv_width;
    float inv_height;
    float position[3];
    float forward[3];
    float right[3];
    float up[3];
} rt_camera;                                /* == Camera, 64 B */

/* ---- bsphere.cpp:316-320 -------------------------------------------------------- */
typedef struct rt_bsphere {
    float center[3];
    float radius;
    uint32_t c0;                            /* 0 == leaf sentinel (index 0 is the root) */
    uint32_t c1;
} rt_bsphere;                               /* == BoundingSphere, 24 B */

/* ---- scene.h:3-15 --------------------------------------------------------------- */
#define RT_LIGHT_DIRECTIONAL 0
#define RT_LIGHT_POINT       1
typedef struct rt_light {
    int32_t type;
    float color[4];
    float position[3];
    float facing[3];
    float falloff;
} rt_light;                                 /* == LightSource, 48 B */

/* ---- mesh.h:8-13 ---------------------------------------------------------------- */
typedef struct rt_texture {
    uint32_t size_x;
    uint32_t size_y;
    uint32_t channels;                      /* 1..4, u8 per channel, row-major, as stbi_load returns */
    const uint8_t *texels;
} rt_texture;                               /* == Texture, 24 B */

/* ---- mesh.h:15-32 (texture pointers become indices into rt_scene_desc.textures) - */
typedef struct rt_material {
    float specular_intensity;
    float index_of_refraction;
    float alpha;
    float ambient_color[4];
    float diffuse_color[4];
    float specular_color[4];
    float emissive_color[4];                /* carried, unused by the reference's integrator */
    int32_t ambient_texture;                /* -1 == NULL */
    int32_t diffuse_texture;
    int32_t specular_texture;
    int32_t alpha_texture;
    int32_t bump_texture;                   /* already converted to a 3-channel normal map (texture.cpp:102-144) */
} rt_material;

/* ---- globals.h:3-7 -------------------------------------------------------------- */
typedef struct rt_counters {
    uint64_t ray_count;                     /* every TraceRay call: primary + shadow + bounce + alpha continuation */
    uint64_t sphere_check_count;            /* bounding-sphere tests in OUR cluster hierarchy (differs from the reference's) */
    uint64_t mesh_check_count;              /* leaf (triangle-cluster) scans in OUR hierarchy */
} rt_counters;                              /* == DebugCounters */

/* device-side measurements of the last call on a scene */
typedef struct rt_stats {
    double   gpu_ms;                        /* CUDA-event time of the render region on the scene's stream */
    double   trace_ms;                      /* CUDA-event time summed over the closest-hit kernel launches (0 unless RT_FLAG_TIME_KERNELS) */
    double   shadow_ms;                     /* same for the occlusion kernel launches */
    double   logic_ms;                      /* same for the shade / bounce-generation kernel launches */
    uint64_t kernel_launches;               /* kernels launched by the call */
    uint64_t waves;                         /* wavefront iterations */
    uint64_t closest_rays;                  /* rays through the closest-hit kernel */
    uint64_t shadow_rays;                   /* rays through the occlusion kernel */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
} rt_stats;

/* ---- globals.h:9-22 (the members the hot path reads) + main.cpp:308-309 ----------
 * base_seed: the per-(pixel,sample) seeding contract. Sample s of
 * linear pixel index p draws from a RandomState seeded with
 *     Random_Seed(base_seed ^ (p * 0x9E3779B97F4A7C15 + s))          (u64 wrap-around)
 * random.h arithmetic unchanged. The oracle harness uses the same contract.
 * Limits (RT_ERR_ARG otherwise): bounce_depth <= 200; reflection_samples + spec_samples <= 1e6; and the recursion tree of ONE sample
 * -- (reflection + spec + 1)^level nodes per level, 1 + reflection_samples draws per node -- must not be able to take more than
 * 65535 random draws (the per-path draw counter is 16 bits; the defaults take at most 2 + 13 * 2 = 28). */
typedef struct rt_params {
    float    ray_bias;
    uint32_t reflection_samples;
    uint32_t spec_samples;
    uint32_t bounce_depth;
    float    background_color[4];
    uint32_t min_samples;                   /* main.cpp:308; fixed spp when min == max */
    uint32_t max_samples;                   /* main.cpp:309 */
    uint64_t base_seed;
} rt_params;

/* ---- raytracer.cpp:20-30 ------------------------------------------------------- */
typedef struct rt_hit {
    float    t;                             /* FLT_MAX on miss, as best_hit = { FLT_MAX } (raytracer.cpp:166) */
    float    bw[3];
    uint32_t vertex0;                       /* index into the group's index buffer, multiple of 3 */
    float    position[3];
    float    normal[3];                     /* geometric, normalised */
    int32_t  object;                        /* index of the leaf BoundingSphere / SceneObject; -1 on miss (object == NULL) */
    uint32_t hit;                           /* TraceRay's bool return */
} rt_hit;

/* ---- mesh.h:36-55, scene.h:22-36, bsphere.cpp:322-326 flattened to POD ---------- */
typedef struct rt_scene_desc {
    /* Mesh vertex streams (Mesh::positions / texcoords / normals / tangents) */
    uint32_t n_positions;  const float *positions;   /* xyz */
    uint32_t n_texcoords;  const float *texcoords;   /* uv  */
    uint32_t n_normals;    const float *normals;     /* xyz */
    const float *tangents;                           /* xyz per NORMAL index (mesh.h:115-117); NULL if no bump maps */

    /* Mesh groups: the three index buffers of every MeshGroup concatenated; group g owns
     * indices [group_first[g], group_first[g+1]) of each buffer (a multiple of 3). */
    uint32_t n_groups;
    const uint32_t *group_first;                     /* n_groups + 1 */
    const uint32_t *idx_positions;
    const uint32_t *idx_texcoords;
    const uint32_t *idx_normals;
    const int32_t  *group_material;                  /* per group; -1 == scene.default_mat (material 0 is NOT implied) */

    /* BoundingHierarchy::spheres (pre-order, bsphere.cpp:328-350) and, per sphere, the mesh
     * group it holds (BoundingHierarchy::mesh_groups / Scene::objects): -1 for internal nodes.
     * Only the leaf visit ORDER (c1 subtree before c0, raytracer.cpp:208-209) is used by the
     * GPU core, to reproduce the reference's first-encountered tie-break among equal-t hits;
     * traversal runs on a GPU-built cluster hierarchy. */
    uint32_t n_spheres;
    const rt_bsphere *spheres;
    const int32_t    *sphere_group;

    uint32_t n_materials;  const rt_material *materials;
    rt_material default_material;                    /* Scene::default_mat (main.cpp:579) */
    uint32_t n_textures;   const rt_texture  *textures;
    uint32_t n_lights;     const rt_light    *lights;  /* Scene::lights[0 .. light_count) */
} rt_scene_desc;

typedef struct rt_scene rt_scene;           /* opaque device-resident scene + GPU-built hierarchy */

/* flags for rt_render* */
#define RT_OUT_MEAN        0u               /* out = sum / samples, w = 1   (RenderPixel, main.cpp:262-263) */
#define RT_OUT_SUM         1u               /* out = raw sample sum, w = number of samples (for sample-range splits) */
#define RT_OUT_FULLFRAME   2u               /* device output is a W*H frame; pixel p is written at p, others untouched */
#define RT_FLAG_COUNTERS   4u               /* also count sphere / cluster tests (slower) */
#define RT_FLAG_TIME_KERNELS 8u             /* CUDA-event-time the trace kernels (adds events, no syncs) */
#define RT_FLAG_PIN_HOST   32u              /* rt_render / rt_render_combined / rt_render_multi: page-lock the caller's output buffer (cudaHostRegister) the
                                               first time it is seen and keep it registered while the same pointer comes back, so the download is one
                                               DMA at PCIe speed instead of a staged pageable copy. The buffer must stay allocated until the scene / comm
                                               is destroyed or a different buffer is passed. Falls back silently if the registration fails */
#define RT_FLAG_ADAPTIVE   16u              /* RenderPixel's adaptive second loop (main.cpp:245-258): min_samples fixed samples, then up to
                                               max_samples with the variance test; sample_count is ignored */

/* ---- lifecycle ------------------------------------------------------------------ */

/* Copies every array of `desc` to device `device`, gathers per-triangle SoA records,
 * builds the bounding-sphere cluster hierarchy on the GPU (replaces the per-ray use of
 * BuildHierarchy's output, bsphere.cpp:379-444) and the host-computed decode tables
 * (sRGB->linear, color.h:13-21; Hammersley directions, raytracer.cpp:273-341).
 * The caller keeps ownership of its arrays. */
int  rt_scene_create(const rt_scene_desc *desc, int device, rt_scene **out_scene);
void rt_scene_destroy(rt_scene *scene);

/* Number of CUDA devices this process sees (0 if none / no driver): lets a host that does not link the CUDA runtime pick
 * `device = rank % rt_device_count()`. */
int  rt_device_count(void);

/* Thread-local description of the last error on this thread ("" if none). */
const char *rt_last_error(void);
int  rt_abi_version(void);

/* ---- Render / RenderTask / RenderPixel (main.cpp:224-358) ----------------------- */

/* Renders pixels of a width x height frame and returns them in HOST memory, 4 floats per
 * pixel, in the order of the pixel list. If pixel_ids is NULL the pixels are the linear
 * row-major range [pixel_begin, pixel_begin + pixel_count) exactly like a RenderJob
 * (main.cpp:316-317); otherwise pixel_ids[0 .. pixel_count) (pixel_begin ignored).
 * Samples [sample_begin, sample_begin + sample_count) of every pixel are traced. */
int rt_render(rt_scene *scene, const rt_camera *cam, const rt_params *params,
              uint32_t width, uint32_t height,
              const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count,
              uint32_t sample_begin, uint32_t sample_count, uint32_t flags,
              float *out_rgba_host, rt_counters *out_counters);

/* Same, output left in DEVICE memory (for the NCCL combine that replaces MPI_Gather,
 * main.cpp:345-347). `stream` is the caller's cudaStream_t; NULL (handle 0) means the LEGACY
 * default stream. The render runs on the scene's own stream, ordered AFTER everything already
 * enqueued on `stream` (the caller's memset of the frame, a previous combine reading it), and `stream`
 * is made to wait for the result; the call itself returns after the scene's stream has drained
 * (the counters are read back). pixel_ids is a HOST pointer. */
int rt_render_device(rt_scene *scene, const rt_camera *cam, const rt_params *params,
                     uint32_t width, uint32_t height,
                     const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count,
                     uint32_t sample_begin, uint32_t sample_count, uint32_t flags,
                     float *out_rgba_device, void *stream, rt_counters *out_counters);

/* ---- Render's partition + MPI_Gather (main.cpp:311-319, 345-347) on N GPUs ---------
 *
 * The reference shards contiguous pixel-index ranges over MPI ranks (scene replicated per rank) and gathers the
 * ranges on rank 0. Here one "rank" is one GPU; the combine is an NCCL reduce over NVLink (or, inside one process,
 * direct peer-memory stores / loads), followed on the root by the resolve it implies -- all inside the library:
 *
 *   RT_PART_TILES    interleaved tile x tile squares, round-robin over ranks (contiguous ranges load-balance badly:
 *                    NOTES.txt:25). Ranks write disjoint pixels of zero-initialised frames, so the reduce(SUM) is
 *                    bit-identical to the reference's MPI_Gather (x + 0 == x).
 *   RT_PART_RANGES   the reference's own split: rank r renders [r * cpp, (r + 1) * cpp), cpp = ceil(W*H / ranks).
 *   RT_PART_SAMPLES  every rank renders all pixels for a sample sub-range as raw sums; reduce(SUM), then / total
 *                    samples, w = 1 (main.cpp:262-263). Not the reference's partition: toleranced (2e-6), not bit-exact.
 */
typedef struct rt_comm rt_comm;             /* one rank's handle on a group of GPUs (NCCL communicator + combine buffers) */

#define RT_COMM_ID_BYTES 128
#define RT_PART_TILES   0
#define RT_PART_RANGES  1
#define RT_PART_SAMPLES 2

/* ncclGetUniqueId: called by ONE rank, which hands the bytes to the others through the host's own channel
 * (the reference host has MPI: one MPI_Bcast; bench.py uses torch.distributed's store). */
int rt_comm_unique_id(uint8_t out_id[RT_COMM_ID_BYTES]);
/* ncclCommInitRank for `rank` of `n_ranks` on CUDA device `device`; collective over the ranks (one process or thread each). */
int rt_comm_create(int n_ranks, int rank, const uint8_t id[RT_COMM_ID_BYTES], int device, rt_comm **out_comm);
/* One process driving n GPUs: out_comms[i] is rank i on devices[i] (devices == NULL: 0 .. n-1). Peer access between the
 * devices is enabled where the hardware allows it; rt_render_multi then gathers through peer memory and needs no NCCL. */
int rt_comm_create_local(int n, const int *devices, rt_comm **out_comms);
void rt_comm_destroy(rt_comm *comm);
int rt_comm_rank(const rt_comm *comm);
int rt_comm_size(const rt_comm *comm);

/* Host-only (no device needed): the linear pixel ids of the tiles rank `rank` of `n_ranks` owns under RT_PART_TILES --
 * tile t = ty * tiles_x + tx goes to rank t % n_ranks; inside a tile pixels are listed row by row. out_ids may be NULL
 * (count only). *out_count receives the number of ids; out_ids must hold that many. */
int rt_partition_tiles(uint32_t width, uint32_t height, uint32_t tile, int rank, int n_ranks, uint32_t *out_ids, uint32_t *out_count);

/* This rank's share of Render() AND the combine, in one call (collective over the ranks of `comm`): partition
 * (main.cpp:311-317 or the tile / sample form), render on `scene`'s GPU, ncclReduce of the float4 frames to `root` on the
 * render stream, the resolve the partition implies (samples: / total, w = 1), and -- optionally, root only -- the tone map
 * (rt_tonemap_device) and the download. Samples per pixel = params->min_samples (RT_FLAG_ADAPTIVE: min..max, pixel
 * partitions only). flags: RT_FLAG_ADAPTIVE | RT_FLAG_TIME_KERNELS | RT_FLAG_COUNTERS.
 *   out_rgba_host   root: W*H*4 floats (Framebuffer.pixels) or NULL (frame stays on the device, see rt_comm_frame)
 *   out_rgba8_host  root: W*H*4 bytes after the reference's tone map, or NULL;  out_scene_luma: its LogAverageLuma
 *   out_counters    root: counters summed over ranks; other ranks: their own */
int rt_render_combined(rt_scene *scene, rt_comm *comm, const rt_camera *cam, const rt_params *params,
                       uint32_t width, uint32_t height, int partition, uint32_t tile, uint32_t flags, int root,
                       float *out_rgba_host, uint8_t *out_rgba8_host, float *out_scene_luma, rt_counters *out_counters);

/* Render() on n GPUs of ONE process (what RenderB200 calls when a rank sees several GPUs): scenes[i] lives on the device of
 * comms[i] (rt_comm_create_local). One host thread per GPU; with peer access the tiles are stored straight into the root GPU's
 * frame over NVLink (no reduce, no zero-fill) and sample sums are added by ONE kernel on the root reading its peers' frames in
 * rank order (deterministic); without peer access the ranks fall back to rt_render_combined's NCCL reduce. */
int rt_render_multi(rt_scene *const *scenes, rt_comm *const *comms, int n, const rt_camera *cam, const rt_params *params,
                    uint32_t width, uint32_t height, int partition, uint32_t tile, uint32_t flags,
                    float *out_rgba_host, uint8_t *out_rgba8_host, float *out_scene_luma, rt_counters *out_counters);

/* The combined W*H float4 frame of the last rt_render_combined / rt_render_multi on this rank (device pointer; meaningful on root). */
const float *rt_comm_frame(const rt_comm *comm);
/* out[0] = ms of zero-fill + reduce + resolve (CUDA events on the render stream), out[1] = ms of tone map + download,
 * out[2] = bytes this rank contributed to the reduce, out[3] = 1 if the last combine went through peer memory, 0 = NCCL */
int rt_comm_get_stats(const rt_comm *comm, double out[4]);

/* ---- TraceRay (raytracer.cpp:159-232) ------------------------------------------- */

/* TraceRay for n arbitrary rays (host in, host out). mode:
 *   RT_TRACE_CLOSEST  closest hit, every rt_hit field as TraceRay leaves it;
 *   RT_TRACE_ANY      the occlusion traversal used for ShadeLight's directional shadow rays
 *                     (raytracer.cpp:385): only `hit` (TraceRay's bool) is meaningful;
 *   RT_TRACE_BRUTE    closest hit by testing every triangle, no hierarchy (diagnostic: checks that
 *                     the GPU hierarchy's pruning is conservative). */
#define RT_TRACE_CLOSEST 0
#define RT_TRACE_ANY     1
#define RT_TRACE_BRUTE   2
int rt_trace_rays(rt_scene *scene, const rt_params *params, const rt_ray *rays, uint64_t n,
                  int mode, rt_hit *out_hits, rt_counters *out_counters);

/* Primary rays of RenderPixel's first loop (main.cpp:237-241) for the given pixels and
 * samples: ray generation (Random_Seed + two jitter draws + MakeCameraRay) and closest hit.
 * out_rays / out_hits hold pixel_count * sample_count entries, pixel-major. Either may be NULL. */
int rt_trace_primary(rt_scene *scene, const rt_camera *cam, const rt_params *params,
                     uint32_t width, uint32_t height,
                     const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count,
                     uint32_t sample_begin, uint32_t sample_count,
                     rt_ray *out_rays, rt_hit *out_hits);

/* ---- TraceRayColor (raytracer.cpp:413-577) -------------------------------------- */

/* Radiance of n arbitrary rays with iters = params->bounce_depth; ray i draws from
 * Random_Seed(seeds[i]). out_rgba: 4 floats per ray (w as the reference leaves it is not
 * reproduced: w = 1 on hit paths is not guaranteed by the reference either; we return 0). */
int rt_trace_color(rt_scene *scene, const rt_params *params, const rt_ray *rays,
                   const uint64_t *seeds, uint64_t n, float *out_rgba, rt_counters *out_counters);

/* ---- BuildHierarchy (bsphere.cpp:379-444) ------------------------------------------ */

/* The reference's OWN hierarchy over mesh groups -- leaf sphere per group (EigenSphere + Ritter_Iterative), greedy merge of
 * the pair with the smallest parent radius, pre-order flattening -- built on the GPU with bit-identical output (the host
 * build is O(groups^3): ~3 min at 5,000 groups). Fills out_spheres / out_sphere_group (2 * n_groups - 1 entries) in the
 * layout rt_scene_desc.spheres / sphere_group expects. n_groups <= 65535, no empty group. */
int rt_build_group_hierarchy(int device, const float *positions, uint32_t n_positions, uint32_t n_groups, const uint32_t *group_first,
                             const uint32_t *idx_positions, rt_bsphere *out_spheres, int32_t *out_sphere_group, uint32_t *out_count);

/* ---- load-time preprocessing (mesh.h:59-129, texture.cpp:85-144) ------------------- */

/* CalculateTangents: per-triangle UV-delta tangents of the bump-mapped groups accumulated on the NORMAL index in triangle
 * order, then normalised -- bit-identical to the reference (the accumulation order is reproduced). out_tangents: 3 floats per
 * normal (what rt_scene_desc.tangents expects). group_has_bump[g] != 0 iff the group's material has a bump texture. */
int rt_calculate_tangents(int device, const float *positions, uint32_t n_positions, const float *texcoords, uint32_t n_texcoords,
                          uint32_t n_normals, uint32_t n_groups, const uint32_t *group_first, const uint32_t *idx_positions,
                          const uint32_t *idx_texcoords, const uint32_t *idx_normals, const uint8_t *group_has_bump, float *out_tangents);
/* ConvertHeightMapToNormalMap + WriteNormal: 1-channel height map -> 3-channel normal map (stored sRGB-encoded like the
 * reference does). The encode's powf is CUDA's double pow rounded to float: a texel on a truncation boundary may differ by 1. */
int rt_height_to_normal_map(int device, uint32_t size_x, uint32_t size_y, const uint8_t *height_host, uint8_t *out_rgb_host);

/* ---- WriteFramebufferImage minus the PNG (main.cpp:78-127, color.h:94-111) -------- */

/* Global log-average-luma Reinhard tone map (key 0.18) + Color_Pack to RGBA8 of a W*H float4 frame that is already on the
 * device (e.g. rt_render_device's output): the step right after Render(). out_scene_luma_host (optional) receives
 * LogAverageLuma. Downloading the 4-byte pixels instead of the 16-byte ones is the point; stbi_write_png stays on the host. */
int rt_tonemap_device(int device, const float *rgba_device, uint32_t width, uint32_t height, uint8_t *out_rgba8_device,
                      float *out_scene_luma_host, void *stream);
/* Same with host buffers (upload, tone map, download). */
int rt_tonemap(int device, const float *rgba_host, uint32_t width, uint32_t height, uint8_t *out_rgba8_host, float *out_scene_luma);

/* ---- introspection --------------------------------------------------------------- */
int rt_get_stats(const rt_scene *scene, rt_stats *out);

/* Per-pixel sample counts -- RenderPixel's final `samp` (main.cpp:262) -- of the last render on this scene that ran
 * with RT_FLAG_ADAPTIVE and min_samples < max_samples; n <= that render's pixel_count, order of its pixel list. */
int rt_get_sample_counts(rt_scene *scene, uint32_t *out_host, uint32_t n);

/* GPU-built hierarchy facts: out[0]=triangles, [1]=clusters(leaves), [2]=nodes, [3]=max depth,
 * [4]=bytes of node array, [5]=bytes of triangle records, [6]=build microseconds, [7]=reserved */
int rt_get_hierarchy_info(const rt_scene *scene, uint64_t out[8]);

/* Host-only (no device needed): the 15-bit quantisation grid rt_scene_create lays over scene bounds [lo, hi] for the default child
 * bound of the GPU hierarchy (no reference counterpart: the reference's BoundingSphere nodes, bsphere.cpp:316-320, are floats).
 * Plane q in [0, 32767] of axis a sits at mid[a] + (32768 + q) * step[a]. Returns RT_OK and *ok = 1 when the grid covers the bounds,
 * *ok = 0 when no grid can (non-finite bounds: the scene then uses float boxes). Test hook for the placement's edge cases. */
int rt_quant_grid(const float lo[3], const float hi[3], float step[3], float mid[3], int *ok);

/* Random_Seed + n x Random_Next on the device (random.h:9-42); KAT hook for the parity tests. */
int rt_rng_kat(int device, uint64_t seed, uint32_t n, uint64_t *out_host);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H_ */
