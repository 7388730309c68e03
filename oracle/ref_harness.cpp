// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the UNMODIFIED reference (ACEfanatic02/par_raytracer) from where it lies under
// /root/reference into oracle/_ref/libref_harness.so (recipe: oracle/Makefile) and exposes
// its static functions through a small C interface, so that
//   * the C restatement in oracle/rt_oracle.c can be pinned against the real reference, and
//   * golden vectors for tests/golden/ can be generated (tests/golden/make_golden.py), and
//   * bench.py --impl reference / cpu_baseline(kind="reference") can time the reference's own
//     TraceRayColor/TraceRay on the host cores.
// No reference source is copied: main.cpp is #included by path with `main` renamed. The only
// code here is glue: loops that call the reference's functions and copy results out.
// Nothing under par_raytracer_b200/ (the product) links or loads this file.
#include <thread>
#include <vector>
#include <map>
#include <chrono>
#include <string>
#include <cstring>

#define main reference_main
#include "main.cpp"            // found through -I/root/reference
#undef main

#include "rt_b200.h"            // POD mirrors used for export (include/)

namespace {

struct Loaded {
    Mesh *mesh = nullptr;
    BoundingHierarchy hierarchy;
    Scene scene;
    std::vector<Material *> materials;   // index -> Material*, order of first use by a group
    std::map<Material *, int> material_index;
    std::vector<Texture *> textures;
    std::map<Texture *, int> texture_index;
    bool ok = false;
};
Loaded *g = nullptr;

int TextureIndex(Texture *t) {
    if (!t) return -1;
    auto it = g->texture_index.find(t);
    if (it != g->texture_index.end()) return it->second;
    int idx = (int)g->textures.size();
    g->textures.push_back(t);
    g->texture_index[t] = idx;
    return idx;
}

int MaterialIndex(Material *m) {
    if (!m) return -1;
    auto it = g->material_index.find(m);
    if (it != g->material_index.end()) return it->second;
    int idx = (int)g->materials.size();
    g->materials.push_back(m);
    g->material_index[m] = idx;
    TextureIndex(m->ambient_texture);
    TextureIndex(m->diffuse_texture);
    TextureIndex(m->specular_texture);
    TextureIndex(m->alpha_texture);
    TextureIndex(m->bump_texture);
    return idx;
}

void FillMaterial(rt_material *o, Material *m) {
    o->specular_intensity = m->specular_intensity;
    o->index_of_refraction = m->index_of_refraction;
    o->alpha = m->alpha;
    memcpy(o->ambient_color, &m->ambient_color, 16);
    memcpy(o->diffuse_color, &m->diffuse_color, 16);
    memcpy(o->specular_color, &m->specular_color, 16);
    memcpy(o->emissive_color, &m->emissive_color, 16);
    o->ambient_texture = TextureIndex(m->ambient_texture);
    o->diffuse_texture = TextureIndex(m->diffuse_texture);
    o->specular_texture = TextureIndex(m->specular_texture);
    o->alpha_texture = TextureIndex(m->alpha_texture);
    o->bump_texture = TextureIndex(m->bump_texture);
}

Camera ToCamera(const rt_camera *c) {
    Camera cam;
    static_assert(sizeof(Camera) == sizeof(rt_camera), "Camera layout");
    memcpy(&cam, c, sizeof(cam));
    return cam;
}

void FillHit(rt_hit *o, bool hit, const RaycastHit &h) {
    o->t = h.t;
    o->bw[0] = h.bw.x; o->bw[1] = h.bw.y; o->bw[2] = h.bw.z;
    o->vertex0 = h.vertex0;
    memcpy(o->position, &h.position, 12);
    memcpy(o->normal, &h.normal, 12);
    o->object = -1;
    if (hit && h.object) {
        for (size_t i = 0; i < g->scene.objects.size(); ++i) {
            if (g->scene.objects[i] == h.object) { o->object = (int32_t)i; break; }
        }
    }
    o->hit = hit ? 1u : 0u;
}

inline u64 SampleSeed(u64 base_seed, u32 pixel, u32 sample) {
    // The per-(pixel, sample) seeding contract (include/rt_b200.h, rt_params.base_seed).
    return base_seed ^ ((u64)pixel * 0x9E3779B97F4A7C15ULL + (u64)sample);
}

// One sample of RenderPixel (main.cpp:237-243 / 246-251) with its own freshly seeded stream.
// g++ evaluates the two Random_NextFloat11 arguments of `Vector2 sample_offset(...)` right to
// left (checked by ref_check_jitter_order below): y gets the first draw.
inline Vector4 SeededSample(Camera *cam, Scene *scene, u32 x, u32 y, u64 seed, float jitter_scale,
                            DebugCounters *debug) {
    RandomState rng;
    Random_Seed(&rng, seed);
    float jy = Random_NextFloat11(&rng);
    float jx = Random_NextFloat11(&rng);
    Vector2 sample_offset(jx, jy);
    Vector2 base_position(x, y);
    Ray ray = MakeCameraRay(cam, base_position + sample_offset * jitter_scale);
    return TraceRayColor(ray, scene, gParams.bounce_depth, debug, &rng);
}

// RenderPixel (main.cpp:224-265) with per-sample reseeding. min == max gives the fixed-spp mean.
Vector4 SeededPixel(Camera *cam, Scene *scene, u32 width, u32 x, u32 y, u32 sample_begin, u32 min_samples,
                    u32 max_samples, u64 base_seed, bool sum_only, DebugCounters *debug, u32 *out_samples) {
    u32 pixel = y * width + x;
    std::vector<Vector4> scratch(max_samples ? max_samples : 1);
    Vector4 color;
    u32 samp = 0;
    for (; samp < min_samples; ++samp) {
        scratch[samp] = SeededSample(cam, scene, x, y, SampleSeed(base_seed, pixel, sample_begin + samp), 0.5f, debug);
        color += scratch[samp];
    }
    if (min_samples < max_samples) {
        float var = CalculateVariance(scratch.data(), samp);
        (void)var;
        for (; samp < max_samples; ++samp) {
            scratch[samp] = SeededSample(cam, scene, x, y, SampleSeed(base_seed, pixel, sample_begin + samp), 1.0f, debug);
            color += scratch[samp];
            var = CalculateVariance(scratch.data(), samp);
            if (var <= 0.01f) break;
        }
    }
    if (out_samples) *out_samples = samp;
    if (!sum_only) {
        color /= samp;
        color.w = 1.0f;
    }
    return color;
}

} // namespace

extern "C" {

// ---- scene -----------------------------------------------------------------------------------

// ParseOBJ + CalculateTangents + BuildHierarchy + InitScene + objects, exactly main.cpp:544-599.
// `dir` must contain sponza.obj (file name hard-coded at main.cpp:553).
int ref_load_scene(const char *dir) {
    if (g) { /* reference never frees; neither do we */ }
    g = new Loaded;
    char *argv0[] = { (char *)"ref", nullptr };
    InitParams(1, argv0);
    Matrix33 transform;
    transform.SetIdentity();
    char *d = strdup(dir);
    char fname[] = "sponza.obj";
    g->mesh = ParseOBJ(d, fname, transform);
    free(d);
    if (!g->mesh) return -1;
    CalculateTangents(g->mesh);
    BuildHierarchy(&g->hierarchy, g->mesh);
    g->scene = InitScene();
    g->scene.hierarchy = &g->hierarchy;
    g->scene.default_mat = MakeMaterial(Vector4(0.75f, 0.5f, 0.75f, 1.0f));
    for (u32 i = 0; i < g->hierarchy.mesh_groups.size(); ++i) {
        MeshGroup *mg = g->hierarchy.mesh_groups[i];
        SceneObject *obj = (SceneObject *)calloc(1, sizeof(SceneObject));
        obj->mesh_group = mg;
        obj->mesh = g->mesh;
        obj->type = ObjectType_MeshGroup;
        obj->material = g->scene.default_mat;
        if (mg && mg->material) obj->material = mg->material;
        g->scene.objects.push_back(obj);
    }
    for (u32 i = 0; i < g->mesh->groups.size(); ++i) MaterialIndex(g->mesh->groups[i].material);
    g->ok = true;
    return 0;
}

void ref_set_params(float ray_bias, u32 reflection_samples, u32 spec_samples, u32 bounce_depth, const float *bg) {
    gParams.ray_bias = ray_bias;
    gParams.reflection_samples = reflection_samples;
    gParams.spec_samples = spec_samples;
    gParams.bounce_depth = bounce_depth;
    gParams.background_color = Vector4(bg[0], bg[1], bg[2], bg[3]);
}

void ref_get_params(rt_params *p) {
    p->ray_bias = gParams.ray_bias;
    p->reflection_samples = gParams.reflection_samples;
    p->spec_samples = gParams.spec_samples;
    p->bounce_depth = gParams.bounce_depth;
    memcpy(p->background_color, &gParams.background_color, 16);
    p->min_samples = 10;   // main.cpp:308
    p->max_samples = 50;   // main.cpp:309
    p->base_seed = gRNGInitTable[0];
}

void ref_set_lights(u32 n, const rt_light *lights) {
    static_assert(sizeof(LightSource) == sizeof(rt_light), "LightSource layout");
    LightSource *l = (LightSource *)calloc(n ? n : 1, sizeof(LightSource));
    memcpy(l, lights, n * sizeof(LightSource));
    g->scene.lights = l;
    g->scene.light_count = n;
}

// sizes: [0]=positions [1]=texcoords [2]=normals [3]=groups [4]=total indices [5]=spheres
//        [6]=materials [7]=textures [8]=lights [9]=has tangents
void ref_export_sizes(u64 *out) {
    Mesh *m = g->mesh;
    u64 total_idx = 0;
    for (auto &mg : m->groups) total_idx += mg.idx_positions.size();
    out[0] = m->positions.size();
    out[1] = m->texcoords.size();
    out[2] = m->normals.size();
    out[3] = m->groups.size();
    out[4] = total_idx;
    out[5] = g->hierarchy.spheres.size();
    out[6] = g->materials.size();
    out[7] = g->textures.size();
    out[8] = g->scene.light_count;
    out[9] = m->tangents.size() == m->normals.size() ? 1 : 0;
}

void ref_texture_info(u32 idx, u32 *out3) {
    Texture *t = g->textures[idx];
    out3[0] = t->size_x; out3[1] = t->size_y; out3[2] = t->channels;
}

void ref_texture_copy(u32 idx, u8 *dst) {
    Texture *t = g->textures[idx];
    memcpy(dst, t->texels, (size_t)t->size_x * t->size_y * t->channels);
}

void ref_export_fill(float *positions, float *texcoords, float *normals, float *tangents, u32 *group_first,
                     u32 *idx_p, u32 *idx_t, u32 *idx_n, int32_t *group_material, rt_bsphere *spheres,
                     int32_t *sphere_group, rt_material *materials, rt_material *default_material,
                     rt_light *lights) {
    Mesh *m = g->mesh;
    memcpy(positions, m->positions.data(), m->positions.size() * 12);
    memcpy(texcoords, m->texcoords.data(), m->texcoords.size() * 8);
    memcpy(normals, m->normals.data(), m->normals.size() * 12);
    if (tangents && m->tangents.size() == m->normals.size()) memcpy(tangents, m->tangents.data(), m->tangents.size() * 12);
    u32 at = 0;
    for (size_t gi = 0; gi < m->groups.size(); ++gi) {
        MeshGroup &mg = m->groups[gi];
        group_first[gi] = at;
        size_t n = mg.idx_positions.size();
        memcpy(idx_p + at, mg.idx_positions.data(), n * 4);
        memcpy(idx_t + at, mg.idx_texcoords.data(), n * 4);
        memcpy(idx_n + at, mg.idx_normals.data(), n * 4);
        group_material[gi] = MaterialIndex(mg.material);
        at += (u32)n;
    }
    group_first[m->groups.size()] = at;
    static_assert(sizeof(BoundingSphere) == sizeof(rt_bsphere), "BoundingSphere layout");
    memcpy(spheres, g->hierarchy.spheres.data(), g->hierarchy.spheres.size() * sizeof(BoundingSphere));
    for (size_t i = 0; i < g->hierarchy.mesh_groups.size(); ++i) {
        MeshGroup *mg = g->hierarchy.mesh_groups[i];
        sphere_group[i] = mg ? (int32_t)(mg - &m->groups[0]) : -1;
    }
    for (size_t i = 0; i < g->materials.size(); ++i) FillMaterial(&materials[i], g->materials[i]);
    FillMaterial(default_material, g->scene.default_mat);
    memcpy(lights, g->scene.lights, g->scene.light_count * sizeof(LightSource));
}

// ---- function-level probes -----------------------------------------------------------------

void ref_rng_next(u64 seed, u32 n, u64 *out) {
    RandomState s;
    Random_Seed(&s, seed);
    for (u32 i = 0; i < n; ++i) out[i] = Random_Next(&s);
}

void ref_rng_float(u64 seed, u32 n, int which, float *out) {
    RandomState s;
    Random_Seed(&s, seed);
    for (u32 i = 0; i < n; ++i) out[i] = which ? Random_NextFloat11(&s) : Random_NextFloat01(&s);
}

u64 ref_rng_table(u32 process_id, u32 thread_id) {
    u64 seed_idx = (process_id << 3 | thread_id) % array_count(gRNGInitTable);
    return gRNGInitTable[seed_idx];
}

void ref_make_camera(float fov, u32 w, u32 h, const float *pos, const float *facing, rt_camera *out) {
    gParams.camera_position = Vector3(pos[0], pos[1], pos[2]);
    gParams.camera_facing = Vector3(facing[0], facing[1], facing[2]);
    gParams.camera_fov = fov;
    Camera cam = MakeCamera(fov, w, h);
    memcpy(out, &cam, sizeof(cam));
}

void ref_camera_rays(const rt_camera *c, u32 n, const float *xy, rt_ray *out) {
    Camera cam = ToCamera(c);
    for (u32 i = 0; i < n; ++i) {
        Ray r = MakeCameraRay(&cam, Vector2(xy[2 * i], xy[2 * i + 1]));
        memcpy(&out[i], &r, 24);
    }
}

// tri: 9 floats (a, b, c). best_t: value of out_hit->t on entry. out: hit, t, bw.xyz, normal.xyz, position.xyz
void ref_intersect_triangle(u32 n, const rt_ray *rays, const float *tri, const float *best_t, u32 *out_hit,
                            float *out10) {
    for (u32 i = 0; i < n; ++i) {
        Ray r; memcpy(&r, &rays[i], 24);
        const float *p = tri + 9 * i;
        RaycastHit h = { best_t[i] };
        bool hit = IntersectRayTriangle(r, Vector3(p[0], p[1], p[2]), Vector3(p[3], p[4], p[5]), Vector3(p[6], p[7], p[8]), &h);
        out_hit[i] = hit;
        float *o = out10 + 10 * i;
        o[0] = h.t; o[1] = h.bw.x; o[2] = h.bw.y; o[3] = h.bw.z;
        o[4] = h.normal.x; o[5] = h.normal.y; o[6] = h.normal.z;
        o[7] = h.position.x; o[8] = h.position.y; o[9] = h.position.z;
    }
}

// sphere: 4 floats. out: hit, t
void ref_intersect_sphere(u32 n, const rt_ray *rays, const float *sph, u32 *out_hit, float *out_t) {
    for (u32 i = 0; i < n; ++i) {
        Ray r; memcpy(&r, &rays[i], 24);
        Sphere s; memcpy(&s, sph + 4 * i, 16);
        RaycastHit h = {};
        out_hit[i] = IntersectRaySphere(r, s, &h);
        out_t[i] = h.t;
    }
}

void ref_hammersley(u32 n, const u32 *i, const u32 *N, float *out2) {
    for (u32 k = 0; k < n; ++k) {
        Vector2 v = Hammersley(i[k], N[k]);
        out2[2 * k] = v.x; out2[2 * k + 1] = v.y;
    }
}

void ref_diffuse_rays(u32 n, const float *origin, const float *normal, const float *xi, rt_ray *out) {
    for (u32 k = 0; k < n; ++k) {
        Ray r = GetDiffuseReflectionRay(Vector3(origin[3 * k], origin[3 * k + 1], origin[3 * k + 2]),
                                        Vector3(normal[3 * k], normal[3 * k + 1], normal[3 * k + 2]),
                                        Vector2(xi[2 * k], xi[2 * k + 1]));
        memcpy(&out[k], &r, 24);
    }
}

void ref_specular_rays(u32 n, const float *origin, const float *normal, const float *spec, const float *xi,
                       rt_ray *out) {
    for (u32 k = 0; k < n; ++k) {
        Ray r = GetSpecularReflectionRay(Vector3(origin[3 * k], origin[3 * k + 1], origin[3 * k + 2]),
                                         Vector3(normal[3 * k], normal[3 * k + 1], normal[3 * k + 2]), spec[k],
                                         Vector2(xi[2 * k], xi[2 * k + 1]));
        memcpy(&out[k], &r, 24);
    }
}

void ref_fresnel(u32 n, const float *ior_exit, const float *ior_enter, const float *normal, const float *incident,
                 float *out) {
    for (u32 k = 0; k < n; ++k) {
        out[k] = FresnelAmount(ior_exit[k], ior_enter[k], Vector3(normal[3 * k], normal[3 * k + 1], normal[3 * k + 2]),
                               Vector3(incident[3 * k], incident[3 * k + 1], incident[3 * k + 2]));
    }
}

void ref_texture_sample(u32 tex, u32 n, const float *uv, float *out4) {
    Texture *t = g->textures[tex];
    for (u32 k = 0; k < n; ++k) {
        Vector4 c = Texture_SampleBilinear(t, uv[2 * k], uv[2 * k + 1]);
        memcpy(out4 + 4 * k, &c, 16);
    }
}

// Texture_SampleBilinear on a caller-supplied texture (no scene needed).
void ref_texture_sample_raw(u32 sx, u32 sy, u32 ch, const u8 *texels, u32 n, const float *uv, float *out4) {
    Texture t; t.size_x = sx; t.size_y = sy; t.channels = ch; t.texels = (u8 *)texels;
    for (u32 k = 0; k < n; ++k) {
        Vector4 c = Texture_SampleBilinear(&t, uv[2 * k], uv[2 * k + 1]);
        memcpy(out4 + 4 * k, &c, 16);
    }
}

// ConvertHeightMapToNormalMap (texture.cpp:102-144) on a 1-channel map; dst = sx*sy*3 bytes.
void ref_height_to_normal(u32 sx, u32 sy, const u8 *height, u8 *dst) {
    Texture t; t.size_x = sx; t.size_y = sy; t.channels = 1; t.texels = (u8 *)height;
    Texture *r = ConvertHeightMapToNormalMap(&t);
    memcpy(dst, r->texels, (size_t)sx * sy * 3);
    free(r->texels); free(r);
}

void ref_srgb_lut(float *out256) {
    for (int i = 0; i < 256; ++i) out256[i] = Color_SRGBToLinear((float)i * gOneOver255);
}

// ---- TraceRay / TraceRayColor --------------------------------------------------------------

void ref_trace_rays(u64 n, const rt_ray *rays, rt_hit *out, rt_counters *counters) {
    DebugCounters dbg = {};
    for (u64 i = 0; i < n; ++i) {
        Ray r; memcpy(&r, &rays[i], 24);
        RaycastHit h;
        bool hit = TraceRay(r, &g->scene, &h, &dbg);
        FillHit(&out[i], hit, h);
    }
    if (counters) memcpy(counters, &dbg, sizeof(dbg));
}

void ref_trace_color(u64 n, const rt_ray *rays, const u64 *seeds, float *out_rgba, rt_counters *counters) {
    DebugCounters dbg = {};
    for (u64 i = 0; i < n; ++i) {
        Ray r; memcpy(&r, &rays[i], 24);
        RandomState rng;
        Random_Seed(&rng, seeds[i]);
        Vector4 c = TraceRayColor(r, &g->scene, gParams.bounce_depth, &dbg, &rng);
        memcpy(out_rgba + 4 * i, &c, 16);
    }
    if (counters) memcpy(counters, &dbg, sizeof(dbg));
}

// Primary rays + closest hits under the per-(pixel, sample) contract. Entries pixel-major.
void ref_trace_primary(const rt_camera *c, u32 width, u32 height, const u32 *pixel_ids, u32 pixel_begin,
                       u32 pixel_count, u32 sample_begin, u32 sample_count, u64 base_seed, rt_ray *out_rays,
                       rt_hit *out_hits) {
    Camera cam = ToCamera(c);
    DebugCounters dbg = {};
    (void)height;
    for (u32 k = 0; k < pixel_count; ++k) {
        u32 pixel = pixel_ids ? pixel_ids[k] : pixel_begin + k;
        u32 x = pixel % width, y = pixel / width;
        for (u32 s = 0; s < sample_count; ++s) {
            RandomState rng;
            Random_Seed(&rng, SampleSeed(base_seed, pixel, sample_begin + s));
            float jy = Random_NextFloat11(&rng);
            float jx = Random_NextFloat11(&rng);
            Ray ray = MakeCameraRay(&cam, Vector2(x, y) + Vector2(jx, jy) * 0.5f);
            size_t o = (size_t)k * sample_count + s;
            if (out_rays) memcpy(&out_rays[o], &ray, 24);
            if (out_hits) {
                RaycastHit h;
                bool hit = TraceRay(ray, &g->scene, &h, &dbg);
                FillHit(&out_hits[o], hit, h);
            }
        }
    }
}

// The seeded render (RenderPixel semantics, per-(pixel,sample) streams) on `threads` host threads,
// pixels split into equal contiguous chunks like MPI ranks (main.cpp:311-319). Returns seconds.
double ref_render_seeded(const rt_camera *c, u32 width, u32 height, const u32 *pixel_ids, u32 pixel_begin,
                         u32 pixel_count, u32 sample_begin, u32 min_samples, u32 max_samples, u64 base_seed,
                         int sum_only, u32 threads, float *out_rgba, u32 *out_nsamples, rt_counters *counters) {
    Camera cam = ToCamera(c);
    (void)height;
    if (threads < 1) threads = 1;
    std::vector<DebugCounters> dbg(threads);
    memset(dbg.data(), 0, sizeof(DebugCounters) * threads);
    std::vector<std::thread> pool;
    u32 per = (pixel_count + threads - 1) / threads;
    auto t0 = std::chrono::steady_clock::now();
    for (u32 t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            Camera local = cam;
            u32 b = t * per, e = (t + 1) * per;
            if (e > pixel_count) e = pixel_count;
            for (u32 k = b; k < e; ++k) {
                u32 pixel = pixel_ids ? pixel_ids[k] : pixel_begin + k;
                u32 ns = 0;
                Vector4 col = SeededPixel(&local, &g->scene, width, pixel % width, pixel / width, sample_begin,
                                          min_samples, max_samples, base_seed, sum_only != 0, &dbg[t], &ns);
                memcpy(out_rgba + 4 * (size_t)k, &col, 16);
                if (out_nsamples) out_nsamples[k] = ns;
            }
        });
    }
    for (auto &th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (auto &d : dbg) {
            counters->ray_count += d.ray_count;
            counters->sphere_check_count += d.sphere_check_count;
            counters->mesh_check_count += d.mesh_check_count;
        }
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// The reference's own mode: K threads each running the UNMODIFIED RenderTask on an equal contiguous
// pixel range with GetRNG(k, 0) (== K MPI ranks + MPI_Gather, main.cpp:311-347). Returns seconds.
double ref_render_ranks(const rt_camera *c, u32 width, u32 height, u32 pixel_begin, u32 pixel_count, u32 min_samples,
                        u32 max_samples, u32 threads, float *out_rgba, rt_counters *counters) {
    Camera cam = ToCamera(c);
    if (threads < 1) threads = 1;
    RenderSharedData shared;
    shared.cam = &cam;
    shared.scene = &g->scene;
    shared.width = width;
    shared.height = height;
    shared.min_samples = min_samples;
    shared.max_samples = max_samples;
    std::vector<DebugCounters> dbg(threads);
    memset(dbg.data(), 0, sizeof(DebugCounters) * threads);
    std::vector<RenderJob> jobs(threads);
    u32 per = (pixel_count + threads - 1) / threads;
    for (u32 t = 0; t < threads; ++t) {
        jobs[t].shared = &shared;
        jobs[t].start_idx = pixel_begin + t * per;
        jobs[t].end_idx = pixel_begin + ((t + 1) * per > pixel_count ? pixel_count : (t + 1) * per);
        jobs[t].buffer = (Vector4 *)(out_rgba + 4 * (size_t)(t * per));
        jobs[t].rng = GetRNG(t, 0);
    }
    std::vector<std::thread> pool;
    auto t0 = std::chrono::steady_clock::now();
    for (u32 t = 0; t < threads; ++t) pool.emplace_back([&, t]() { RenderTask(&jobs[t], &dbg[t]); });
    for (auto &th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (auto &d : dbg) {
            counters->ray_count += d.ray_count;
            counters->sphere_check_count += d.sphere_check_count;
            counters->mesh_check_count += d.mesh_check_count;
        }
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// Self-check of the jitter-order assumption: the UNMODIFIED RenderPixel with one sample, its
// job.rng seeded with `seed`, must equal SeededSample with the same seed. Returns 1 if identical.
int ref_check_jitter_order(const rt_camera *c, u32 width, u32 height, u32 x, u32 y, u64 seed) {
    Camera cam = ToCamera(c);
    RenderSharedData shared;
    shared.cam = &cam; shared.scene = &g->scene; shared.width = width; shared.height = height;
    shared.min_samples = 1; shared.max_samples = 1;
    RenderJob job;
    job.shared = &shared; job.start_idx = 0; job.end_idx = 1; job.buffer = nullptr;
    Random_Seed(&job.rng, seed);
    DebugCounters d0 = {}, d1 = {};
    Vector4 a = RenderPixel(&job, &d0, x, y);
    Vector4 b = SeededSample(&cam, &g->scene, x, y, seed, 0.5f, &d1);
    b.w = 1.0f;
    return memcmp(&a, &b, 16) == 0 && d0.ray_count == d1.ray_count;
}

// The reference's own LogAverageLuma and WriteFramebufferImage (main.cpp:78-131) on a caller-supplied frame; the PNG it
// writes is read back with the stb_image the reference already compiles in. Returns scene_luma.
float ref_tonemap_png(u32 width, u32 height, const float *rgba, const char *tmp_png, u8 *out_rgba8) {
    Framebuffer fb; fb.pixels = (Vector4 *)rgba; fb.width = width; fb.height = height;
    float luma = LogAverageLuma(&fb);
    s32 saved = gMPI_CommRank; gMPI_CommRank = 0;
    char *fn = strdup(tmp_png);
    WriteFramebufferImage(&fb, fn);
    gMPI_CommRank = saved;
    s32 x = 0, y = 0, ch = 0;
    u8 *img = stbi_load(fn, &x, &y, &ch, 4);
    free(fn);
    if (img && (u32)x == width && (u32)y == height) memcpy(out_rgba8, img, (size_t)width * height * 4);
    if (img) stbi_image_free(img);
    return luma;
}

} // extern "C"
