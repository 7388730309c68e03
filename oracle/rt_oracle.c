/*
 * rt_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the render hot path of ACEfanatic02/par_raytracer, operating on the
 * flattened scene of include/rt_b200.h. It exists to CHECK the CUDA path (tests/, __graft_entry__.smoke(),
 * bench.py's cpu_baseline / --impl reference legs). Nothing under par_raytracer_b200/ links, loads or
 * calls it, and it is never the thing measured as the product.
 *
 * Parity status: PINNED. tests/test_oracle_golden.py checks every function below bit-for-bit against
 * vectors produced by the unmodified reference compiled in the authoring container
 * (oracle/ref_harness.cpp -> oracle/_ref/libref_harness.so, vectors in tests/golden/, generator
 * tests/golden/make_golden.py). The reference ships no tests or golden vectors of its own.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march, no fast-math), i.e. the arithmetic the reference's
 * build.sh:6 produces on x86-64: IEEE binary32, no FMA contraction, glibc libm.
 *
 * Every function cites the reference file:line it follows. The code is written against flat arrays
 * (float[3], index buffers) rather than the reference's Vector3 / std::vector types; the ORDER of
 * floating-point operations is the reference's.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rt_b200.h"

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

/* ---- mathlib.h:136-262 (Vector3) ------------------------------------------------------- */
static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_ld(const float *p) { v3 r = { p[0], p[1], p[2] }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_neg(v3 a) { return v3_scale(a, -1.0f); }                       /* mathlib.h:229-232 */
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* mathlib.h:234-237 */
static inline v3 v3_cross(v3 a, v3 b) {                                              /* mathlib.h:239-246 */
    return v3_make(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline v3 v3_normalize(v3 a) {                                                /* mathlib.h:253-262 */
    float l2 = v3_dot(a, a);
    if (l2 == 0.0f) return a;
    float l = sqrtf(l2);
    return v3_make(a.x / l, a.y / l, a.z / l);
}
/* ---- mathlib.h:264-385 (Vector4) ------------------------------------------------------- */
static inline v4 v4_make(float x, float y, float z, float w) { v4 r = { x, y, z, w }; return r; }
static inline v4 v4_ld(const float *p) { return v4_make(p[0], p[1], p[2], p[3]); }
static inline v4 v4_add(v4 a, v4 b) { return v4_make(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline v4 v4_sub(v4 a, v4 b) { return v4_make(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline v4 v4_mul(v4 a, v4 b) { return v4_make(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline v4 v4_scale(v4 a, float s) { return v4_make(a.x * s, a.y * s, a.z * s, a.w * s); }
#define MAXF(a, b) ((a) > (b) ? (a) : (b))                                         /* mathlib.h:8 */
#define MINF(a, b) ((a) < (b) ? (a) : (b))                                         /* mathlib.h:7 */
#define CLAMPF(n, a, b) (MINF(MAXF(n, a), b))                                      /* mathlib.h:9 */

/* ======================================================================================== */
/* random.h:4-61                                                                            */
/* ======================================================================================== */
typedef struct { uint64_t s[16]; int32_t p; } orc_rng;

void orc_rng_seed(orc_rng *r, uint64_t seed) {                 /* random.h:9-27 (note >>25) */
    if (seed == 0) seed = 0x5555555555555555ULL;
    r->p = 0;
    uint64_t x = seed;
    for (int i = 0; i < 16; ++i) {
        x ^= x >> 12;
        x ^= x >> 25;
        x ^= x >> 27;
        r->s[i] = x * 2685821657736338717ULL;
    }
}

uint64_t orc_rng_next(orc_rng *r) {                            /* random.h:29-42 (note &=) */
    uint64_t s0 = r->s[r->p];
    r->p = (r->p + 1) & 15;
    uint64_t s1 = r->s[r->p];
    s1 ^= s1 << 31;
    s1 ^= s1 >> 11;
    s0 &= s0 >> 30;
    r->s[r->p] = s0 ^ s1;
    return r->s[r->p] * 1181783497276652981ULL;
}

float orc_rng_float01(orc_rng *r) {                            /* random.h:49-56 */
    const uint64_t max_value = 0xFFFFFFFFFFFFFFFFULL;
    float f = (float)orc_rng_next(r) / (float)max_value;
    return CLAMPF(f, 0.0f, 1.0f);
}

float orc_rng_float11(orc_rng *r) { return (orc_rng_float01(r) * 2.0f) - 1.0f; }  /* random.h:58-61 */

static inline uint64_t sample_seed(uint64_t base, uint32_t pixel, uint32_t sample) {
    return base ^ ((uint64_t)pixel * 0x9E3779B97F4A7C15ULL + (uint64_t)sample);    /* rt_b200.h contract */
}

/* ======================================================================================== */
/* scene handle                                                                             */
/* ======================================================================================== */
typedef struct orc_scene {
    rt_scene_desc d;          /* borrowed pointers: the caller keeps the arrays alive */
    int32_t *group_sphere;    /* group -> sphere index */
} orc_scene;

orc_scene *orc_scene_create(const rt_scene_desc *desc) {
    orc_scene *s = (orc_scene *)calloc(1, sizeof(*s));
    s->d = *desc;
    s->group_sphere = (int32_t *)malloc(sizeof(int32_t) * (desc->n_groups ? desc->n_groups : 1));
    for (uint32_t i = 0; i < desc->n_spheres; ++i)
        if (desc->sphere_group[i] >= 0) s->group_sphere[desc->sphere_group[i]] = (int32_t)i;
    return s;
}

void orc_scene_destroy(orc_scene *s) {
    if (!s) return;
    free(s->group_sphere);
    free(s);
}

static inline const rt_material *object_material(const orc_scene *s, int32_t sphere) {
    int32_t g = s->d.sphere_group[sphere];                      /* main.cpp:586-589 */
    int32_t m = g >= 0 ? s->d.group_material[g] : -1;
    return m >= 0 ? &s->d.materials[m] : &s->d.default_material;
}

/* ======================================================================================== */
/* main.cpp:164-177 MakeCameraRay                                                           */
/* ======================================================================================== */
rt_ray orc_camera_ray(const rt_camera *cam, float ox, float oy) {
    float nx = 2.0f * (ox + 0.5f) * cam->inv_width - 1.0f;
    float ny = 1.0f - 2.0f * (oy + 0.5f) * cam->inv_height;
    v3 fwd = v3_ld(cam->forward), right = v3_ld(cam->right), up = v3_ld(cam->up);
    v3 a = v3_scale(v3_scale(v3_scale(right, cam->tan_a2), cam->aspect), nx);
    v3 b = v3_scale(v3_scale(up, cam->tan_a2), ny);
    v3 dir = v3_normalize(v3_add(v3_add(fwd, a), b));
    rt_ray r;
    memcpy(r.origin, cam->position, 12);
    r.direction[0] = dir.x; r.direction[1] = dir.y; r.direction[2] = dir.z;
    return r;
}

/* ======================================================================================== */
/* raytracer.cpp:32-60 IntersectRaySphere                                                   */
/* ======================================================================================== */
int orc_intersect_sphere(v3 o, v3 d, const float *center, float radius, float *out_t) {
    v3 m = v3_sub(o, v3_ld(center));
    float b = v3_dot(m, d);
    float c = v3_dot(m, m) - radius * radius;
    if (c > 0.0f && b > 0.0f) return 0;
    float disc = b * b - c;
    if (disc < 0.0f) return 0;
    float t = -b - sqrtf(disc);
    if (t < 0.0f) t = 0.0f;
    *out_t = t;
    return 1;
}

/* ======================================================================================== */
/* raytracer.cpp:82-125 IntersectRayTriangle. io_t: in = the caller's current best t          */
/* ======================================================================================== */
typedef struct { float t; v3 bw; uint32_t vertex0; v3 position; v3 normal; int32_t object; } orc_hit;

int orc_intersect_triangle(v3 o, v3 d, v3 a, v3 b, v3 c, orc_hit *h) {
    v3 ab = v3_sub(b, a);
    v3 ac = v3_sub(c, a);
    v3 q = v3_add(o, d);
    v3 qp = v3_sub(o, q);
    v3 n = v3_cross(ab, ac);
    float dd = v3_dot(qp, n);
    if (dd <= 0.0f) return 0;
    v3 ap = v3_sub(o, a);
    float t = v3_dot(ap, n);
    if (t < 0.0f) return 0;
    if (t > h->t * dd) return 0;
    v3 e = v3_cross(qp, ap);
    float v = v3_dot(ac, e);
    if (v < 0.0f || v > dd) return 0;
    float w = -v3_dot(ab, e);
    if (w < 0.0f || (v + w) > dd) return 0;
    float ood = 1.0f / dd;
    h->t = t * ood;
    h->bw.y = v * ood;
    h->bw.z = w * ood;
    h->bw.x = 1.0f - h->bw.y - h->bw.z;
    h->position = v3_add(o, v3_scale(d, h->t));
    h->normal = v3_normalize(n);
    return 1;
}

/* ======================================================================================== */
/* raytracer.cpp:127-157 IntersectRayMesh (one mesh group = one hierarchy leaf)              */
/* ======================================================================================== */
static int intersect_group(const orc_scene *s, v3 o, v3 d, int32_t sphere, orc_hit *io) {
    const rt_scene_desc *sc = &s->d;
    int32_t g = sc->sphere_group[sphere];
    uint32_t first = sc->group_first[g], last = sc->group_first[g + 1];
    int hit = 0;
    orc_hit best;
    memset(&best, 0, sizeof(best));
    best.t = io->t;
    best.object = -1;
    for (uint32_t i = first; i < last; i += 3) {
        v3 pa = v3_ld(sc->positions + 3 * (size_t)sc->idx_positions[i + 0]);
        v3 pb = v3_ld(sc->positions + 3 * (size_t)sc->idx_positions[i + 1]);
        v3 pc = v3_ld(sc->positions + 3 * (size_t)sc->idx_positions[i + 2]);
        orc_hit cur;
        memset(&cur, 0, sizeof(cur));
        cur.t = best.t;
        if (orc_intersect_triangle(o, d, pa, pb, pc, &cur)) {
            cur.vertex0 = i - first;
            cur.object = sphere;
            if (cur.t < best.t) { best = cur; hit = 1; }
        }
    }
    *io = best;
    return hit;
}

/* ======================================================================================== */
/* raytracer.cpp:159-232 TraceRay                                                           */
/* ======================================================================================== */
int orc_trace_ray(const orc_scene *s, const rt_params *P, const rt_ray *ray, orc_hit *out, rt_counters *dbg) {
    dbg->ray_count++;
    v3 d = v3_ld(ray->direction);
    v3 o = v3_add(v3_ld(ray->origin), v3_scale(d, P->ray_bias));
    int hit = 0;
    orc_hit best;
    memset(&best, 0, sizeof(best));
    best.t = FLT_MAX;
    best.object = -1;
    uint32_t stack_small[128];
    uint32_t *stack = stack_small, cap = 128, sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        uint32_t i = stack[--sp];
        const rt_bsphere *bs = &s->d.spheres[i];
        float st;
        dbg->sphere_check_count++;
        if (orc_intersect_sphere(o, d, bs->center, bs->radius, &st)) {
            if (st > best.t) continue;
            if (bs->c0 && bs->c1) {
                if (sp + 2 > cap) {
                    uint32_t *n = (uint32_t *)malloc(sizeof(uint32_t) * cap * 2);
                    memcpy(n, stack, sizeof(uint32_t) * sp);
                    if (stack != stack_small) free(stack);
                    stack = n; cap *= 2;
                }
                stack[sp++] = bs->c0;
                stack[sp++] = bs->c1;          /* popped first: c1 subtree before c0 */
            } else {
                orc_hit cur;
                memset(&cur, 0, sizeof(cur));
                cur.t = best.t;
                dbg->mesh_check_count++;
                if (intersect_group(s, o, d, (int32_t)i, &cur)) {
                    if (cur.t < best.t) { best = cur; hit = 1; }
                }
            }
        }
    }
    if (stack != stack_small) free(stack);
    if (out) *out = best;
    return hit;
}

/* ======================================================================================== */
/* color.h:13-21, texture.cpp:5-83                                                          */
/* ======================================================================================== */
static inline float srgb_to_linear(float srgb) {
    if (srgb <= 0.04045f) return srgb / 12.92f;
    return powf((srgb + 0.055f) / 1.055f, 2.4f);
}

static inline float wrap_uv(float uv) {                         /* texture.cpp:5-13 */
    if (uv >= 0.0f) return fmodf(uv, 1.0f);
    return 1.0f + fmodf(uv, 1.0f);
}

static v4 get_texel(const rt_texture *t, uint32_t x, uint32_t y) {   /* texture.cpp:17-51 */
    uint32_t idx = y * t->size_x + x;
    uint8_t r = 0, g = 0, b = 0, a = 0;
    const uint8_t *p = t->texels + (size_t)idx * t->channels;
    if (t->channels >= 4) a = p[3];
    if (t->channels >= 3) b = p[2];
    if (t->channels >= 2) g = p[1];
    if (t->channels >= 1) r = p[0];
    if (t->channels < 4) a = 255;
    if (t->channels == 1) { g = r; b = r; }
    const float one_over_255 = 1.0f / 255.0f;
    v4 c = v4_scale(v4_make((float)r, (float)g, (float)b, (float)a), one_over_255);
    c.x = srgb_to_linear(c.x);
    c.y = srgb_to_linear(c.y);
    c.z = srgb_to_linear(c.z);
    c.w = srgb_to_linear(c.w);
    return c;
}

static inline v4 v4_lerp(v4 a, v4 b, float t) { return v4_add(a, v4_scale(v4_sub(b, a), t)); }  /* mathlib.h:10 */

v4 orc_texture_sample(const rt_texture *t, float u, float v) {  /* texture.cpp:53-83 */
    if (!t) return v4_make(0, 0, 0, 0);
    u = wrap_uv(u);
    v = 1.0f - wrap_uv(v);
    float sx = (float)(t->size_x - 2);
    float sy = (float)(t->size_y - 2);
    float tx = CLAMPF(u * sx, 0.0f, sx);
    float ty = CLAMPF(v * sy, 0.0f, sy);
    uint32_t tx0 = (uint32_t)floorf(tx);
    uint32_t ty0 = (uint32_t)floorf(ty);
    uint32_t tx1 = tx0 + 1, ty1 = ty0 + 1;
    float fx = tx - tx0;
    float fy = ty - ty0;
    v4 s00 = get_texel(t, tx0, ty0);
    v4 s01 = get_texel(t, tx0, ty1);
    v4 s10 = get_texel(t, tx1, ty0);
    v4 s11 = get_texel(t, tx1, ty1);
    return v4_lerp(v4_lerp(s00, s01, fy), v4_lerp(s10, s11, fy), fx);
}

/* ======================================================================================== */
/* raytracer.cpp:273-371 samplers, Reflect, Fresnel                                         */
/* ======================================================================================== */
#define ORC_PI32 (3.1415927f)                                   /* brt.h:23 */

float orc_radical_inverse(uint32_t bits) {                      /* raytracer.cpp:273-282 */
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)(bits * 2.3283064365386963e-10);              /* double product, narrowed on return */
}

void orc_hammersley(uint32_t i, uint32_t n, float *xi) {        /* raytracer.cpp:284-288 */
    xi[0] = (float)i / (float)n;
    xi[1] = orc_radical_inverse(i);
}

static v3 to_world(v3 normal, v3 local) {                       /* raytracer.cpp:306-312 / 330-336 */
    v3 up = fabsf(normal.z) < 0.9999f ? v3_make(0, 0, 1) : v3_make(1, 0, 0);
    v3 tangent = v3_normalize(v3_cross(up, normal));
    v3 bitangent = v3_normalize(v3_cross(normal, tangent));
    v3 w = v3_add(v3_add(v3_scale(tangent, local.x), v3_scale(bitangent, local.y)), v3_scale(normal, local.z));
    return v3_normalize(w);
}

v3 orc_diffuse_direction(v3 normal, const float *xi) {          /* raytracer.cpp:320-341 */
    float phi = xi[1] * 2.0f * ORC_PI32;
    float cp = cosf(phi);
    float sp = sinf(phi);
    float ct = sqrtf(1.0f - xi[0]);
    float st = sqrtf(1.0f - ct * ct);
    return to_world(normal, v3_make(cp * st, sp * st, ct));
}

v3 orc_specular_direction(v3 normal, float e, const float *xi) { /* raytracer.cpp:290-318 */
    float phi = 2.0f * ORC_PI32 * xi[0];
    float cp = cosf(phi);
    float sp = sinf(phi);
    float ct = powf(1.0f - xi[1], 1.0f / (e + 1.0f));
    float st = sqrtf(1.0f - (ct * ct));
    return to_world(normal, v3_make(cp * st, sp * st, ct));
}

static inline v3 reflect(v3 v, v3 n) {                          /* raytracer.cpp:343-346 */
    return v3_sub(v3_scale(v3_scale(n, 2.0f), v3_dot(v, n)), v);
}

float orc_fresnel(float ior_exit, float ior_enter, v3 normal, v3 incident) {   /* raytracer.cpp:348-371 */
    float r0 = (ior_exit - ior_enter) / (ior_exit + ior_enter);
    r0 *= r0;
    float ct = MAXF(0.0f, -v3_dot(normal, incident));
    if (ior_exit > ior_enter) {
        float n = ior_exit / ior_enter;
        float st_sq = n * n * (1.0f - ct * ct);
        if (st_sq > 1.0f) return 1.0f;
        ct = sqrtf(1.0f - st_sq);
    }
    float x = 1.0f - ct;
    float x2 = x * x;
    float x3 = x * x2;
    return r0 + (1.0f - r0) * x2 * x3;
}

/* ======================================================================================== */
/* raytracer.cpp:234-250, 378-411 ShadeLight                                                */
/* ======================================================================================== */
static void shade_light(const orc_scene *s, const rt_params *P, const rt_light *L, v3 view_dir, v3 normal, v3 position,
                        float spec_intensity, rt_counters *dbg, v4 *diffuse, v4 *specular) {
    *diffuse = v4_make(0, 0, 0, 0);
    *specular = v4_make(0, 0, 0, 0);
    rt_ray sray;
    sray.origin[0] = position.x; sray.origin[1] = position.y; sray.origin[2] = position.z;
    v4 color = v4_ld(L->color);
    if (L->type == RT_LIGHT_DIRECTIONAL) {
        v3 lv = v3_scale(v3_ld(L->facing), -1.0f);
        sray.direction[0] = lv.x; sray.direction[1] = lv.y; sray.direction[2] = lv.z;
        if (!orc_trace_ray(s, P, &sray, NULL, dbg)) {
            float spec_cos = v3_dot(v3_scale(view_dir, -1.0f), reflect(lv, normal));
            *diffuse = v4_scale(v4_scale(color, 2.0f), MAXF(0.0f, v3_dot(normal, lv)));
            *specular = v4_scale(color, powf(MAXF(0.0f, spec_cos), spec_intensity));
        }
    } else {
        v3 lp = v3_ld(L->position);
        v3 lv = v3_normalize(v3_sub(lp, position));
        sray.direction[0] = lv.x; sray.direction[1] = lv.y; sray.direction[2] = lv.z;
        v3 dv = v3_sub(lp, position);
        float dist_sq = v3_dot(dv, dv);
        orc_hit h;
        /* raytracer.cpp:395-396: lit when nothing is hit OR the nearest hit is NEARER than the light (sic) */
        if (!orc_trace_ray(s, P, &sray, &h, dbg) || h.t * h.t <= dist_sq) {
            float fd = (sqrtf(dist_sq) / L->falloff) + 1.0f;
            v4 lc = v4_scale(color, 1.0f / (fd * fd));
            float spec_cos = v3_dot(v3_scale(view_dir, -1.0f), reflect(lv, normal));
            *diffuse = v4_scale(v4_scale(lc, 2.0f), MAXF(0.0f, v3_dot(normal, lv)));
            *specular = v4_scale(lc, powf(MAXF(0.0f, spec_cos), spec_intensity));
        }
    }
}

/* ======================================================================================== */
/* raytracer.cpp:413-577 TraceRayColor                                                      */
/* ======================================================================================== */
static inline const rt_texture *tex_or_null(const orc_scene *s, int32_t idx) {
    return idx >= 0 ? &s->d.textures[idx] : NULL;
}

v4 orc_trace_color(const orc_scene *s, const rt_params *P, rt_ray ray, int32_t iters, rt_counters *dbg, orc_rng *rng) {
    v4 color = v4_make(0, 0, 0, 0);
    if (iters < 0 || (iters != (int32_t)P->bounce_depth && orc_rng_float01(rng) < 0.5f)) return color;

    orc_hit hit;
    if (!orc_trace_ray(s, P, &ray, &hit, dbg)) return v4_ld(P->background_color);

    const rt_scene_desc *sc = &s->d;
    v3 rd = v3_ld(ray.direction);
    v3 hit_p = v3_add(hit.position, v3_scale(hit.normal, P->ray_bias));
    v3 hit_normal = hit.normal;
    const rt_material *mat = object_material(s, hit.object);
    v4 ambient = v4_ld(mat->ambient_color);
    v4 diffuse = v4_ld(mat->diffuse_color);
    v4 specular = v4_ld(mat->specular_color);
    float alpha = mat->alpha;

    {
        int32_t g = sc->sphere_group[hit.object];
        uint32_t base = sc->group_first[g] + hit.vertex0;
        float bw[3] = { hit.bw.x, hit.bw.y, hit.bw.z };
        float u = 0.0f, v = 0.0f;
        for (int k = 0; k < 3; ++k) {
            const float *tc = sc->texcoords + 2 * (size_t)sc->idx_texcoords[base + k];
            u += tc[0] * bw[k];
            v += tc[1] * bw[k];
        }
        if (mat->alpha <= 1.0f || mat->alpha_texture >= 0) {
            if (mat->alpha_texture >= 0) alpha *= orc_texture_sample(tex_or_null(s, mat->alpha_texture), u, v).x;
            if (alpha <= 0.05f) {
                v3 no = v3_add(hit.position, v3_scale(v3_scale(rd, P->ray_bias), 2.0f));
                ray.origin[0] = no.x; ray.origin[1] = no.y; ray.origin[2] = no.z;
                return orc_trace_color(s, P, ray, iters, dbg, rng);
            }
        }
        if (mat->ambient_texture >= 0) ambient = v4_mul(ambient, orc_texture_sample(tex_or_null(s, mat->ambient_texture), u, v));
        if (mat->diffuse_texture >= 0) diffuse = v4_mul(diffuse, orc_texture_sample(tex_or_null(s, mat->diffuse_texture), u, v));
        if (mat->specular_texture >= 0) specular = orc_texture_sample(tex_or_null(s, mat->specular_texture), u, v);

        v3 n = v3_make(0, 0, 0);
        for (int k = 0; k < 3; ++k)
            n = v3_add(n, v3_scale(v3_ld(sc->normals + 3 * (size_t)sc->idx_normals[base + k]), bw[k]));
        hit_normal = v3_normalize(n);
        if (mat->bump_texture >= 0) {
            v3 tg = v3_make(0, 0, 0);
            for (int k = 0; k < 3; ++k)
                tg = v3_add(tg, v3_scale(v3_ld(sc->tangents + 3 * (size_t)sc->idx_normals[base + k]), bw[k]));
            tg = v3_normalize(tg);
            v3 bt = v3_normalize(v3_cross(hit_normal, tg));
            v4 smp = orc_texture_sample(tex_or_null(s, mat->bump_texture), u, v);
            v3 sn = v3_sub(v3_scale(v3_make(smp.x, smp.y, smp.z), 2.0f), v3_make(1.0f, 1.0f, 1.0f));
            v3 w;                                               /* Matrix33 * Vector3, mathlib.h:696-711 */
            w.x = tg.x * sn.x + bt.x * sn.y + hit_normal.x * sn.z;
            w.y = tg.y * sn.x + bt.y * sn.y + hit_normal.y * sn.z;
            w.z = tg.z * sn.x + bt.z * sn.y + hit_normal.z * sn.z;
            hit_normal = w;                                     /* not re-normalised (raytracer.cpp:494-495) */
        }
    }

    v4 direct = v4_make(0, 0, 0, 0), direct_spec = v4_make(0, 0, 0, 0);
    for (uint32_t i = 0; i < sc->n_lights; ++i) {
        v4 dd, ds;
        shade_light(s, P, &sc->lights[i], rd, hit_normal, hit_p, mat->specular_intensity, dbg, &dd, &ds);
        direct = v4_add(direct, dd);
        direct_spec = v4_add(direct_spec, ds);
    }

    v4 indirect = v4_make(0, 0, 0, 0), indirect_spec = v4_make(0, 0, 0, 0);
    if (iters > 0) {
        for (uint32_t samp = 0; samp < P->reflection_samples; ++samp) {
            const uint32_t series_n = 1024;
            uint32_t series_i = (uint32_t)(orc_rng_next(rng) % series_n);
            float xi[2];
            orc_hammersley(series_i, series_n, xi);
            v3 dir = orc_diffuse_direction(hit_normal, xi);
            rt_ray rr = { { hit_p.x, hit_p.y, hit_p.z }, { dir.x, dir.y, dir.z } };
            v4 rc = orc_trace_color(s, P, rr, iters - 1, dbg, rng);
            indirect = v4_add(indirect, v4_scale(rc, MAXF(0.0f, v3_dot(hit_normal, dir))));
        }
        for (uint32_t samp = 0; samp < P->spec_samples; ++samp) {
            float xi[2];
            orc_hammersley(samp, P->spec_samples, xi);
            v3 dir = orc_specular_direction(hit_normal, mat->specular_intensity, xi);
            rt_ray rr = { { hit_p.x, hit_p.y, hit_p.z }, { dir.x, dir.y, dir.z } };
            v4 sc4 = orc_trace_color(s, P, rr, iters - 1, dbg, rng);
            float weight = MAXF(0.0f, v3_dot(dir, v3_neg(rd)));
            indirect_spec = v4_add(indirect_spec, v4_scale(sc4, weight));
        }
    }

    float object_reflectivity = 0.04f;
    float fresnel = orc_fresnel(1.0f, mat->index_of_refraction, hit_normal, rd);
    float w_reflect = (object_reflectivity + (1.0f - object_reflectivity) * fresnel);
    float w_diffuse = 1.0f - w_reflect;

    color = v4_add(color, v4_scale(ambient, 0.1f));
    color = v4_add(color, v4_scale(v4_mul(v4_add(indirect, direct), diffuse), w_diffuse));
    color = v4_add(color, v4_mul(v4_add(indirect_spec, direct_spec), specular));

    if (alpha < 1.0f) {
        v3 no = v3_add(hit.position, v3_scale(v3_scale(rd, P->ray_bias), 2.0f));
        ray.origin[0] = no.x; ray.origin[1] = no.y; ray.origin[2] = no.z;
        v4 back = orc_trace_color(s, P, ray, iters - 1, dbg, rng);
        color = v4_add(v4_scale(color, alpha), v4_scale(back, 1.0f - alpha));
    }
    return color;
}

/* ======================================================================================== */
/* main.cpp:179-265 Color_Distance, CalculateVariance, RenderPixel (per-sample reseeding)    */
/* ======================================================================================== */
static float calc_variance(const v4 *vals, uint32_t count) {    /* main.cpp:206-222 */
    v4 mean = v4_make(0, 0, 0, 0);
    for (uint32_t i = 0; i < count; ++i) mean = v4_add(mean, vals[i]);
    float fc = (float)count;                                    /* Vector4 /= u32: scale converts to float */
    mean = v4_make(mean.x / fc, mean.y / fc, mean.z / fc, mean.w / fc);
    float variance = 0.0f;
    for (uint32_t i = 0; i < count; ++i) {
        float d = fabsf(vals[i].x - mean.x) + fabsf(vals[i].y - mean.y) + fabsf(vals[i].z - mean.z);
        variance += d * d;
    }
    variance /= (count - 1);
    return variance;
}

static v4 one_sample(const orc_scene *s, const rt_camera *cam, const rt_params *P, uint32_t x, uint32_t y, uint64_t seed,
                     float jitter_scale, rt_counters *dbg) {
    orc_rng rng;
    orc_rng_seed(&rng, seed);
    float jy = orc_rng_float11(&rng);                           /* g++ evaluates the 2nd ctor argument first */
    float jx = orc_rng_float11(&rng);                           /* (main.cpp:238; pinned by golden) */
    float px = (float)x + jx * jitter_scale;
    float py = (float)y + jy * jitter_scale;
    rt_ray ray = orc_camera_ray(cam, px, py);
    return orc_trace_color(s, P, ray, (int32_t)P->bounce_depth, dbg, &rng);
}

static v4 render_pixel(const orc_scene *s, const rt_camera *cam, const rt_params *P, uint32_t width, uint32_t pixel,
                       uint32_t sample_begin, int sum_only, v4 *scratch, rt_counters *dbg, uint32_t *out_ns) {
    uint32_t x = pixel % width, y = pixel / width;               /* main.cpp:274-275 */
    v4 color = v4_make(0, 0, 0, 0);
    uint32_t samp = 0;
    for (; samp < P->min_samples; ++samp) {                      /* main.cpp:237-243 */
        scratch[samp] = one_sample(s, cam, P, x, y, sample_seed(P->base_seed, pixel, sample_begin + samp), 0.5f, dbg);
        color = v4_add(color, scratch[samp]);
    }
    if (P->min_samples < P->max_samples) {                       /* main.cpp:245-258 */
        float var = calc_variance(scratch, samp);
        (void)var;
        for (; samp < P->max_samples; ++samp) {
            scratch[samp] = one_sample(s, cam, P, x, y, sample_seed(P->base_seed, pixel, sample_begin + samp), 1.0f, dbg);
            color = v4_add(color, scratch[samp]);
            var = calc_variance(scratch, samp);
            if (var <= 0.01f) break;
        }
    }
    if (out_ns) *out_ns = samp;
    if (!sum_only) {                                             /* main.cpp:262-263 */
        float fs = (float)samp;
        color = v4_make(color.x / fs, color.y / fs, color.z / fs, color.w / fs);
        color.w = 1.0f;
    }
    return color;
}

/* ======================================================================================== */
/* batch entry points (ctypes)                                                              */
/* ======================================================================================== */
static void hit_out(rt_hit *o, int hit, const orc_hit *h) {
    o->t = h->t;
    o->bw[0] = h->bw.x; o->bw[1] = h->bw.y; o->bw[2] = h->bw.z;
    o->vertex0 = h->vertex0;
    o->position[0] = h->position.x; o->position[1] = h->position.y; o->position[2] = h->position.z;
    o->normal[0] = h->normal.x; o->normal[1] = h->normal.y; o->normal[2] = h->normal.z;
    o->object = hit ? h->object : -1;
    o->hit = hit ? 1u : 0u;
}

void orc_trace_rays(const orc_scene *s, const rt_params *P, const rt_ray *rays, uint64_t n, rt_hit *out, rt_counters *cnt) {
    rt_counters dbg = { 0, 0, 0 };
    for (uint64_t i = 0; i < n; ++i) {
        orc_hit h;
        int hit = orc_trace_ray(s, P, &rays[i], &h, &dbg);
        hit_out(&out[i], hit, &h);
    }
    if (cnt) *cnt = dbg;
}

void orc_trace_colors(const orc_scene *s, const rt_params *P, const rt_ray *rays, const uint64_t *seeds, uint64_t n,
                      float *out_rgba, rt_counters *cnt) {
    rt_counters dbg = { 0, 0, 0 };
    for (uint64_t i = 0; i < n; ++i) {
        orc_rng rng;
        orc_rng_seed(&rng, seeds[i]);
        v4 c = orc_trace_color(s, P, rays[i], (int32_t)P->bounce_depth, &dbg, &rng);
        memcpy(out_rgba + 4 * i, &c, 16);
    }
    if (cnt) *cnt = dbg;
}

void orc_trace_primary(const orc_scene *s, const rt_camera *cam, const rt_params *P, uint32_t width, uint32_t height,
                       const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                       uint32_t sample_count, rt_ray *out_rays, rt_hit *out_hits) {
    rt_counters dbg = { 0, 0, 0 };
    (void)height;
    for (uint32_t k = 0; k < pixel_count; ++k) {
        uint32_t pixel = pixel_ids ? pixel_ids[k] : pixel_begin + k;
        uint32_t x = pixel % width, y = pixel / width;
        for (uint32_t sidx = 0; sidx < sample_count; ++sidx) {
            orc_rng rng;
            orc_rng_seed(&rng, sample_seed(P->base_seed, pixel, sample_begin + sidx));
            float jy = orc_rng_float11(&rng);
            float jx = orc_rng_float11(&rng);
            rt_ray ray = orc_camera_ray(cam, (float)x + jx * 0.5f, (float)y + jy * 0.5f);
            size_t o = (size_t)k * sample_count + sidx;
            if (out_rays) out_rays[o] = ray;
            if (out_hits) {
                orc_hit h;
                int hit = orc_trace_ray(s, P, &ray, &h, &dbg);
                hit_out(&out_hits[o], hit, &h);
            }
        }
    }
}

typedef struct {
    const orc_scene *s; const rt_camera *cam; const rt_params *P;
    uint32_t width; const uint32_t *pixel_ids; uint32_t pixel_begin, k0, k1, sample_begin; int sum_only;
    float *out; uint32_t *ns; rt_counters dbg;
} render_job;

static void *render_worker(void *arg) {
    render_job *j = (render_job *)arg;
    uint32_t cap = j->P->max_samples > j->P->min_samples ? j->P->max_samples : j->P->min_samples;
    v4 *scratch = (v4 *)calloc(cap ? cap : 1, sizeof(v4));     /* main.cpp:232 */
    for (uint32_t k = j->k0; k < j->k1; ++k) {
        uint32_t pixel = j->pixel_ids ? j->pixel_ids[k] : j->pixel_begin + k;
        uint32_t ns = 0;
        v4 c = render_pixel(j->s, j->cam, j->P, j->width, pixel, j->sample_begin, j->sum_only, scratch, &j->dbg, &ns);
        memcpy(j->out + 4 * (size_t)k, &c, 16);
        if (j->ns) j->ns[k] = ns;
    }
    free(scratch);
    return NULL;
}

/* Render (main.cpp:301-358) over `threads` host threads, each an equal contiguous chunk of the pixel list
 * (== MPI ranks, main.cpp:311-319). Returns wall seconds of the render region ("Render, sync", main.cpp:326-333). */
double orc_render(const orc_scene *s, const rt_camera *cam, const rt_params *P, uint32_t width, uint32_t height,
                  const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                  int sum_only, uint32_t threads, float *out_rgba, uint32_t *out_nsamples, rt_counters *cnt) {
    (void)height;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    render_job jobs[256];
    pthread_t tid[256];
    uint32_t per = (pixel_count + threads - 1) / threads;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t t = 0; t < threads; ++t) {
        render_job *j = &jobs[t];
        memset(j, 0, sizeof(*j));
        j->s = s; j->cam = cam; j->P = P; j->width = width; j->pixel_ids = pixel_ids; j->pixel_begin = pixel_begin;
        j->k0 = t * per; j->k1 = (t + 1) * per > pixel_count ? pixel_count : (t + 1) * per;
        if (j->k0 > pixel_count) j->k0 = pixel_count;
        j->sample_begin = sample_begin; j->sum_only = sum_only; j->out = out_rgba; j->ns = out_nsamples;
        if (threads == 1) render_worker(j); else pthread_create(&tid[t], NULL, render_worker, j);
    }
    if (threads > 1) for (uint32_t t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (cnt) {
        memset(cnt, 0, sizeof(*cnt));
        for (uint32_t t = 0; t < threads; ++t) {
            cnt->ray_count += jobs[t].dbg.ray_count;
            cnt->sphere_check_count += jobs[t].dbg.sphere_check_count;
            cnt->mesh_check_count += jobs[t].dbg.mesh_check_count;
        }
    }
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ======================================================================================== */
/* main.cpp:78-127 LogAverageLuma + WriteFramebufferImage's tone map, color.h:94-111 Color_Luma / Color_Pack  */
/* ======================================================================================== */
static inline float color_luma(v4 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }

float orc_log_average_luma(const float *rgba, uint32_t width, uint32_t height) {       /* main.cpp:78-99 */
    float lavg = 0.0f;
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            float cl = color_luma(v4_ld(rgba + 4 * ((size_t)y * width + x)));
            if (cl > 0.0f) lavg += logf(0.01f + cl);
        }
    return expf(lavg / (float)(width * height));
}

float orc_tonemap(const float *rgba, uint32_t width, uint32_t height, uint8_t *out_rgba8) {   /* main.cpp:107-127 */
    float scene_luma = orc_log_average_luma(rgba, width, height);
    for (size_t idx = 0; idx < (size_t)width * height; ++idx) {
        v4 c = v4_ld(rgba + 4 * idx);
        float key_alpha = 0.18f;
        float pixel_luma = color_luma(c);
        float l_xy = key_alpha * pixel_luma / scene_luma;
        float l_d = l_xy / (1.0f + l_xy);
        float scale = l_d / pixel_luma;
        c.x *= scale; c.y *= scale; c.z *= scale;
        uint8_t *p = out_rgba8 + 4 * idx;                                                 /* Color_Pack, color.h:105-111 */
        p[0] = (uint8_t)(CLAMPF(c.x, 0.0f, 1.0f) * 255.0f);
        p[1] = (uint8_t)(CLAMPF(c.y, 0.0f, 1.0f) * 255.0f);
        p[2] = (uint8_t)(CLAMPF(c.z, 0.0f, 1.0f) * 255.0f);
        p[3] = (uint8_t)(CLAMPF(c.w, 0.0f, 1.0f) * 255.0f);
    }
    return scene_luma;
}

/* ======================================================================================== */
/* bsphere.cpp:8-444 BuildHierarchy: leaf spheres (EigenSphere + Ritter_Iterative), greedy   */
/* min-radius agglomeration, pre-order flattening. Restated on flat arrays; float order kept. */
/* ======================================================================================== */
typedef struct { v3 c; float r; } sph;
typedef struct { float e[9]; } m33;                               /* mathlib.h:540-599, row-major e[i*3+j] */

static m33 m33_identity(void) { m33 m = { { 1, 0, 0, 0, 1, 0, 0, 0, 1 } }; return m; }
static m33 m33_mul(m33 a, m33 b) {                                /* mathlib.h:652-694: (a0*b0 + a1*b1) + a2*b2 */
    m33 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.e[i * 3 + j] = a.e[i * 3 + 0] * b.e[0 * 3 + j] + a.e[i * 3 + 1] * b.e[1 * 3 + j] + a.e[i * 3 + 2] * b.e[2 * 3 + j];
    return r;
}
static m33 m33_transpose(m33 m) {
    m33 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.e[i * 3 + j] = m.e[j * 3 + i];
    return r;
}

static void update_sphere_with_point(sph *s, v3 p) {             /* bsphere.cpp:14-26 */
    v3 pc = v3_sub(p, s->c);
    float sq = v3_dot(pc, pc);
    if (sq > (s->r * s->r)) {
        float dist = sqrtf(sq);
        float nr = (s->r + dist) * 0.5f + 1e-2;                   /* double constant: float + double -> float */
        float k = (nr - s->r) / dist;
        s->r = nr;
        s->c = v3_add(s->c, v3_scale(pc, k));
    }
}

static m33 covariance_matrix(const v3 *pts, uint32_t n) {         /* bsphere.cpp:45-81 (m(2,1) is never set: sic) */
    float inv = 1.0f / (float)n;
    v3 c = v3_make(0, 0, 0);
    float e00 = 0, e11 = 0, e22 = 0, e01 = 0, e02 = 0, e12 = 0;
    for (uint32_t i = 0; i < n; ++i) c = v3_add(c, pts[i]);
    c = v3_scale(c, inv);
    for (uint32_t i = 0; i < n; ++i) {
        v3 p = v3_sub(pts[i], c);
        e00 += p.x * p.x; e11 += p.y * p.y; e22 += p.z * p.z;
        e01 += p.x * p.y; e02 += p.x * p.z; e12 += p.y * p.z;
    }
    m33 m; memset(&m, 0, sizeof(m));
    m.e[0] = e00 * inv; m.e[4] = e11 * inv; m.e[8] = e22 * inv;
    m.e[1] = m.e[3] = e01 * inv;
    m.e[2] = m.e[6] = e02 * inv;
    m.e[5] = e12 * inv;
    return m;
}

static void sym_schur2(m33 m, uint32_t p, uint32_t q, float *c, float *s) {   /* bsphere.cpp:83-102 */
    const float epsilon = 0.0001f;
    if (fabsf(m.e[p * 3 + q]) > epsilon) {
        float r = (m.e[q * 3 + q] - m.e[p * 3 + p]) / (2.0f * m.e[p * 3 + q]);
        float t;
        if (r >= 0.0f) t = 1.0f / (r + sqrtf(1.0f + r * r));
        else t = -1.0f / (-r + sqrtf(1.0f + r * r));
        *c = 1.0f / sqrtf(1.0f + t * t);
        *s = (*c) * t;
    } else { *c = 1.0f; *s = 0.0f; }
}

static void jacobi(m33 *a, m33 *v) {                              /* bsphere.cpp:104-154 */
    float prevoff = 0.0f, c, s;
    *v = m33_identity();
    for (uint32_t n = 0; n < 50; ++n) {
        uint32_t p = 0, q = 1;
        for (uint32_t i = 0; i < 3; ++i)
            for (uint32_t j = 0; j < 3; ++j)
                if (i != j && fabsf(a->e[i * 3 + j]) > fabsf(a->e[p * 3 + q])) { p = i; q = j; }
        sym_schur2(*a, p, q, &c, &s);
        m33 J = m33_identity();
        J.e[p * 3 + p] = c; J.e[p * 3 + q] = s; J.e[q * 3 + p] = -s; J.e[q * 3 + q] = c;
        *v = m33_mul(*v, J);
        *a = m33_mul(m33_mul(m33_transpose(J), *a), J);
        float off = 0.0f;
        for (uint32_t i = 0; i < 3; ++i)
            for (uint32_t j = 0; j < 3; ++j)
                if (i != j) off += a->e[i * 3 + j] * a->e[i * 3 + j];
        if (n > 2 && off >= prevoff) return;
        prevoff = off;
    }
}

static sph eigen_sphere(const v3 *pts, uint32_t n) {              /* bsphere.cpp:156-195 */
    m33 m = covariance_matrix(pts, n), v;
    jacobi(&m, &v);
    uint32_t max_c = 0;
    float max_e = fabsf(m.e[0]);
    if (fabsf(m.e[4]) > max_e) { max_c = 1; max_e = fabsf(m.e[4]); }
    if (fabsf(m.e[8]) > max_e) { max_c = 2; max_e = fabsf(m.e[8]); }
    v3 e = v3_make(v.e[0 * 3 + max_c], v.e[1 * 3 + max_c], v.e[2 * 3 + max_c]);
    uint32_t imin = 0, imax = 0;                                  /* bsphere.cpp:28-43 */
    float minp = FLT_MAX, maxp = -FLT_MAX;
    for (uint32_t i = 0; i < n; ++i) {
        float proj = v3_dot(pts[i], e);
        if (proj < minp) { imin = i; minp = proj; }
        if (proj > maxp) { imax = i; maxp = proj; }
    }
    v3 a = pts[imin], b = pts[imax];
    sph r;
    r.c = v3_scale(v3_add(a, b), 0.5f);
    v3 d = v3_sub(a, b);
    r.r = sqrtf(v3_dot(d, d)) * 0.5f;
    for (uint32_t i = 0; i < n; ++i) update_sphere_with_point(&r, pts[i]);
    return r;
}

static sph ritter_iterative(sph s, v3 *pts, uint32_t n) {         /* bsphere.cpp:197-230 (shuffles pts in place) */
    orc_rng rng;
    orc_rng_seed(&rng, 0x201701260526ull);
    sph s2 = s;
    for (uint32_t k = 0; k < 16; ++k) {
        s2.r *= 0.9f;
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t remaining = n - i - 1;
            if (remaining) {
                uint32_t j = (uint32_t)orc_rng_next(&rng) % remaining;
                j += i + 1;
                v3 tmp = pts[i]; pts[i] = pts[j]; pts[j] = tmp;
            }
            update_sphere_with_point(&s2, pts[i]);
        }
        if (s2.r < s.r) s = s2;
    }
    for (uint32_t i = 0; i < n; ++i) update_sphere_with_point(&s, pts[i]);
    return s;
}

static sph sphere_from_children(sph s0, sph s1) {                 /* bsphere.cpp:248-279 */
    sph r;
    v3 v = v3_sub(s1.c, s0.c);
    float sq = v3_dot(v, v);
    float dr = s1.r - s0.r;
    if ((dr * dr) >= sq) {
        r = (s1.r >= s0.r) ? s1 : s0;
    } else {
        float dist = sqrtf(sq);
        r.r = (dist + s0.r + s1.r) * 0.5f;
        r.c = s0.c;
        if (dist > 0.001f) {
            v = v3_make(v.x / dist, v.y / dist, v.z / dist);
            r.c = v3_add(r.c, v3_scale(v, r.r - s0.r));
        }
    }
    r.r *= 1.0001f;
    return r;
}

/* out_spheres / out_sphere_group: 2 * n_groups - 1 entries. Returns the number of spheres written. */
uint32_t orc_build_hierarchy(const float *positions, uint32_t n_groups, const uint32_t *group_first, const uint32_t *idx_positions,
                             rt_bsphere *out_spheres, int32_t *out_sphere_group) {
    if (n_groups == 0) return 0;
    uint32_t total = 2 * n_groups - 1;
    sph *S = (sph *)malloc(sizeof(sph) * total);
    int32_t *c0 = (int32_t *)malloc(sizeof(int32_t) * total), *c1 = (int32_t *)malloc(sizeof(int32_t) * total);
    for (uint32_t g = 0; g < n_groups; ++g) {                     /* bsphere.cpp:232-246, 384-393 */
        uint32_t n = group_first[g + 1] - group_first[g];
        v3 *pts = (v3 *)calloc(n ? n : 1, sizeof(v3));
        for (uint32_t i = 0; i < n; ++i) pts[i] = v3_ld(positions + 3 * (size_t)idx_positions[group_first[g] + i]);
        sph r = eigen_sphere(pts, n);
        S[g] = ritter_iterative(r, pts, n);
        free(pts);
        c0[g] = -1; c1[g] = -1;
    }
    uint32_t *list = (uint32_t *)malloc(sizeof(uint32_t) * n_groups), m = n_groups, created = n_groups;
    for (uint32_t g = 0; g < n_groups; ++g) list[g] = g;
    while (m >= 2) {                                              /* bsphere.cpp:281-314, 405-427 */
        float best = FLT_MAX; uint32_t bi = m, bj = m; sph merged = S[list[0]];
        for (uint32_t i = 0; i < m; ++i)
            for (uint32_t j = i + 1; j < m; ++j) {
                sph parent = sphere_from_children(S[list[i]], S[list[j]]);
                if (parent.r < best) { bi = i; bj = j; merged = parent; best = parent.r; }
            }
        uint32_t a = list[bi], b = list[bj];
        memmove(list + bj, list + bj + 1, sizeof(uint32_t) * (m - bj - 1));      /* erase the later index first */
        memmove(list + bi, list + bi + 1, sizeof(uint32_t) * (m - 1 - bi - 1));
        m -= 2;
        S[created] = merged; c0[created] = (int32_t)a; c1[created] = (int32_t)b;
        list[m++] = created++;
    }
    /* FlattenHierarchyTree (bsphere.cpp:328-350): pre-order, c0 subtree before c1, child index 0 = leaf sentinel */
    uint32_t *stack = (uint32_t *)malloc(sizeof(uint32_t) * (total + 1)), *slot_of = (uint32_t *)malloc(sizeof(uint32_t) * total), sp = 0, next = 0;
    stack[sp++] = list[0];
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * total);
    while (sp) {
        uint32_t n = stack[--sp];
        slot_of[n] = next; order[next++] = n;
        if (c0[n] >= 0) { stack[sp++] = (uint32_t)c1[n]; stack[sp++] = (uint32_t)c0[n]; }
    }
    for (uint32_t k = 0; k < next; ++k) {
        uint32_t n = order[k];
        out_spheres[k].center[0] = S[n].c.x; out_spheres[k].center[1] = S[n].c.y; out_spheres[k].center[2] = S[n].c.z;
        out_spheres[k].radius = S[n].r;
        out_spheres[k].c0 = c0[n] >= 0 ? slot_of[c0[n]] : 0;
        out_spheres[k].c1 = c1[n] >= 0 ? slot_of[c1[n]] : 0;
        out_sphere_group[k] = c0[n] >= 0 ? -1 : (int32_t)n;
    }
    free(S); free(c0); free(c1); free(list); free(stack); free(slot_of); free(order);
    return next;
}

/* ======================================================================================== */
/* mesh.h:59-129 CalculateTangents; texture.cpp:85-144 WriteNormal / ConvertHeightMapToNormalMap */
/* ======================================================================================== */
void orc_calculate_tangents(const float *positions, const float *texcoords, uint32_t n_normals, uint32_t n_groups, const uint32_t *group_first,
                            const uint32_t *idx_p, const uint32_t *idx_t, const uint32_t *idx_n, const uint8_t *group_has_bump, float *tangents) {
    memset(tangents, 0, sizeof(float) * 3 * (size_t)n_normals);
    for (uint32_t g = 0; g < n_groups; ++g) {
        if (!group_has_bump[g]) continue;                                      /* mesh.h:70-74 */
        for (uint32_t i = group_first[g]; i < group_first[g + 1]; i += 3) {
            v3 p0 = v3_ld(positions + 3 * (size_t)idx_p[i]), p1 = v3_ld(positions + 3 * (size_t)idx_p[i + 1]), p2 = v3_ld(positions + 3 * (size_t)idx_p[i + 2]);
            const float *uv0 = texcoords + 2 * (size_t)idx_t[i], *uv1 = texcoords + 2 * (size_t)idx_t[i + 1], *uv2 = texcoords + 2 * (size_t)idx_t[i + 2];
            v3 dp0 = v3_sub(p1, p0), dp1 = v3_sub(p2, p0);
            float d0x = uv1[0] - uv0[0], d0y = uv1[1] - uv0[1], d1x = uv2[0] - uv0[0], d1y = uv2[1] - uv0[1];
            float f = (d0x * d1y - d1x * d0y);
            if (f <= 1e-7) continue;                                           /* float compared with a double literal */
            f = 1.0f / f;
            v3 t;
            t.x = f * (d1y * dp0.x - d0y * dp1.x);
            t.y = f * (d1y * dp0.y - d0y * dp1.y);
            t.z = f * (d1y * dp0.z - d0y * dp1.z);
            for (int k = 0; k < 3; ++k) {
                float *dst = tangents + 3 * (size_t)idx_n[i + k];
                dst[0] += t.x; dst[1] += t.y; dst[2] += t.z;
            }
        }
    }
    for (uint32_t i = 0; i < n_normals; ++i) {
        v3 t = v3_normalize(v3_ld(tangents + 3 * (size_t)i));
        tangents[3 * (size_t)i] = t.x; tangents[3 * (size_t)i + 1] = t.y; tangents[3 * (size_t)i + 2] = t.z;
    }
}

static inline float linear_to_srgb(float linear) {                             /* color.h:3-11 */
    if (linear <= 0.0031308f) return 12.92f * linear;
    return 1.055f * powf(linear, 1.0f / 2.4f) - 0.055f;
}

void orc_height_to_normal_map(uint32_t sx, uint32_t sy, const uint8_t *height, uint8_t *out_rgb) {   /* texture.cpp:102-125, 85-100 */
    const float one_over_255 = 1.0f / 255.0f;
    for (uint32_t y = 0; y < sy; ++y)
        for (uint32_t x = 0; x < sx; ++x) {
            uint32_t x1 = (x + 1) % sx, y1 = (y + 1) % sy;
            float h00 = srgb_to_linear((float)height[y * sx + x] * one_over_255);
            float h10 = srgb_to_linear((float)height[y * sx + x1] * one_over_255);
            float h01 = srgb_to_linear((float)height[y1 * sx + x] * one_over_255);
            float a = 2.5f;
            v3 n = v3_normalize(v3_make((h01 - h00) * a, (h10 - h00) * a, 1.0f));
            n = v3_scale(v3_add(n, v3_make(1.0f, 1.0f, 1.0f)), 0.5f);
            uint8_t *o = out_rgb + 3 * ((size_t)y * sx + x);
            o[0] = (uint8_t)(linear_to_srgb(n.x) * 255.0f);
            o[1] = (uint8_t)(linear_to_srgb(n.y) * 255.0f);
            o[2] = (uint8_t)(linear_to_srgb(n.z) * 255.0f);
        }
}

/* ---- function-level probes --------------------------------------------------------------- */
void orc_rng_next_n(uint64_t seed, uint32_t n, uint64_t *out) {
    orc_rng r; orc_rng_seed(&r, seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = orc_rng_next(&r);
}
void orc_rng_float_n(uint64_t seed, uint32_t n, int signed11, float *out) {
    orc_rng r; orc_rng_seed(&r, seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = signed11 ? orc_rng_float11(&r) : orc_rng_float01(&r);
}
void orc_camera_rays(const rt_camera *cam, uint32_t n, const float *xy, rt_ray *out) {
    for (uint32_t i = 0; i < n; ++i) out[i] = orc_camera_ray(cam, xy[2 * i], xy[2 * i + 1]);
}
void orc_intersect_triangle_n(uint32_t n, const rt_ray *rays, const float *tri, const float *best_t, uint32_t *out_hit,
                              float *out10) {
    for (uint32_t i = 0; i < n; ++i) {
        const float *p = tri + 9 * (size_t)i;
        orc_hit h; memset(&h, 0, sizeof(h)); h.t = best_t[i];
        out_hit[i] = (uint32_t)orc_intersect_triangle(v3_ld(rays[i].origin), v3_ld(rays[i].direction), v3_ld(p), v3_ld(p + 3),
                                                      v3_ld(p + 6), &h);
        float *o = out10 + 10 * (size_t)i;
        o[0] = h.t; o[1] = h.bw.x; o[2] = h.bw.y; o[3] = h.bw.z; o[4] = h.normal.x; o[5] = h.normal.y; o[6] = h.normal.z;
        o[7] = h.position.x; o[8] = h.position.y; o[9] = h.position.z;
    }
}
void orc_intersect_sphere_n(uint32_t n, const rt_ray *rays, const float *sph, uint32_t *out_hit, float *out_t) {
    for (uint32_t i = 0; i < n; ++i) {
        float t = 0.0f;
        out_hit[i] = (uint32_t)orc_intersect_sphere(v3_ld(rays[i].origin), v3_ld(rays[i].direction), sph + 4 * (size_t)i,
                                                    sph[4 * (size_t)i + 3], &t);
        out_t[i] = t;
    }
}
void orc_hammersley_n(uint32_t n, const uint32_t *i, const uint32_t *N, float *out2) {
    for (uint32_t k = 0; k < n; ++k) orc_hammersley(i[k], N[k], out2 + 2 * (size_t)k);
}
void orc_diffuse_rays(uint32_t n, const float *origin, const float *normal, const float *xi, rt_ray *out) {
    for (uint32_t k = 0; k < n; ++k) {
        v3 d = orc_diffuse_direction(v3_ld(normal + 3 * (size_t)k), xi + 2 * (size_t)k);
        memcpy(out[k].origin, origin + 3 * (size_t)k, 12);
        out[k].direction[0] = d.x; out[k].direction[1] = d.y; out[k].direction[2] = d.z;
    }
}
void orc_specular_rays(uint32_t n, const float *origin, const float *normal, const float *spec, const float *xi, rt_ray *out) {
    for (uint32_t k = 0; k < n; ++k) {
        v3 d = orc_specular_direction(v3_ld(normal + 3 * (size_t)k), spec[k], xi + 2 * (size_t)k);
        memcpy(out[k].origin, origin + 3 * (size_t)k, 12);
        out[k].direction[0] = d.x; out[k].direction[1] = d.y; out[k].direction[2] = d.z;
    }
}
void orc_fresnel_n(uint32_t n, const float *ior_exit, const float *ior_enter, const float *normal, const float *incident,
                   float *out) {
    for (uint32_t k = 0; k < n; ++k)
        out[k] = orc_fresnel(ior_exit[k], ior_enter[k], v3_ld(normal + 3 * (size_t)k), v3_ld(incident + 3 * (size_t)k));
}
void orc_texture_sample_n(const rt_texture *t, uint32_t n, const float *uv, float *out4) {
    for (uint32_t k = 0; k < n; ++k) {
        v4 c = orc_texture_sample(t, uv[2 * (size_t)k], uv[2 * (size_t)k + 1]);
        memcpy(out4 + 4 * (size_t)k, &c, 16);
    }
}
void orc_srgb_lut(float *out256) {
    const float one_over_255 = 1.0f / 255.0f;
    for (int i = 0; i < 256; ++i) out256[i] = srgb_to_linear((float)i * one_over_255);
}
