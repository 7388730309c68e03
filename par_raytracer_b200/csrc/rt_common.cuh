// rt_common.cuh -- device-side math and data layout shared by every kernel of librt_b200.so.
//
// Arithmetic contract: everything that decides a hit, a ray or a branch uses IEEE
// binary32 +,-,*,/ and sqrt in the reference's operation ORDER with no FMA contraction. The whole
// translation unit is compiled with -fmad=false (and the default -prec-div=true -prec-sqrt=true
// -ftz=false); code that is allowed to be approximate (bounding-sphere culling) uses explicit
// __fmaf_rn / approx intrinsics and says so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define RT_DEVICE __device__ __forceinline__

struct f3 { float x, y, z; };

RT_DEVICE f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_DEVICE f3 mk3(float4 v) { return mk3(v.x, v.y, v.z); }
// mathlib.h:199-232
RT_DEVICE f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEVICE f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEVICE f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_DEVICE f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_DEVICE f3 neg3(f3 a) { return a * -1.0f; }                                    // mathlib.h:229-232
RT_DEVICE float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // mathlib.h:234-237
RT_DEVICE f3 cross3(f3 a, f3 b) {                                                // mathlib.h:239-246
    return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
// IEEE x / l for l > 0. A zero numerator sends CUDA's div.rn into its out-of-line slow path (~35 instructions + call): ncu showed 13-18 % of
// k_logic's executed instructions there, because the tangent frames of raytracer.cpp:306-312 always carry an exact 0 component. A branch
// around the division does not help -- the compiler if-converts it and the FCHK of the (now speculative) division still calls the slow
// path -- so the division is given a numerator it can handle (l / l) and the exact answer is selected afterwards: 0 / l is the numerator
// itself (sign kept) for every l that is not NaN. Same bits, no slow path.
RT_DEVICE float div_pos(float x, float l) {
    const bool zero = x == 0.0f;
    const float q = (zero ? l : x) / l;
    return (zero && l == l) ? x : q;
}
RT_DEVICE f3 normalize3(f3 a) {                                                  // mathlib.h:253-262
    float l2 = dot3(a, a);
    if (l2 == 0.0f) return a;
    float l = sqrtf(l2);
    return mk3(div_pos(a.x, l), div_pos(a.y, l), div_pos(a.z, l));
}
RT_DEVICE float max0(float x) { return 0.0f > x ? 0.0f : x; }                    // Max(0.0f, x), mathlib.h:8
RT_DEVICE float clampf(float n, float a, float b) {                              // Clamp, mathlib.h:9
    float m = n > a ? n : a;
    return m < b ? m : b;
}

RT_DEVICE f3 ld3(const float *p, uint32_t i) { return mk3(p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]); }
RT_DEVICE uint32_t find_group(const uint32_t *group_first, uint32_t n_groups, uint32_t index) {
    uint32_t lo = 0, hi = n_groups;     // last g with group_first[g] <= index
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (group_first[mid] <= index) lo = mid; else hi = mid; }
    return lo;
}
RT_DEVICE float4 mk4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
RT_DEVICE float4 mk4u(f3 v, uint32_t w) { return make_float4(v.x, v.y, v.z, __uint_as_float(w)); }

// ---------------------------------------------------------------------------------------------
// device-resident scene
// ---------------------------------------------------------------------------------------------

// One node of the GPU-built bounding-sphere hierarchy: the bounds of BOTH children live in the parent (one 80-byte
// fetch decides both descents). Each child is bounded by a sphere and, inside it, by a slab: the sphere says where the
// subtree is, the slab (unit normal n = the subtree's mean triangle normal, dmin <= n.v <= dmax for every vertex v) says how
// thin it is -- surface patches are thin shells, and a sphere alone admits every ray that grazes them.
// child >= 0: node index; child < 0: cluster (leaf) encoded as -(1 + first_triangle * 8 + triangle_count), count <= 7.
struct __align__(16) HNode {
    float4 s0;          // child 0: center.xyz, radius
    float4 s1;          // child 1
    float4 p0;          // child 0 slab: n.xyz, dmin        (n = 0, dmin = -inf: no slab)
    float4 p1;          // child 1 slab
    float dmax0, dmax1;
    int32_t c0, c1;
};
static_assert(sizeof(HNode) == 80, "HNode");

// Float-box form of the same node (RT_B200_BOUNDS=box): both children's axis-aligned bounds in the parent as centre + half-extent,
// 64 bytes = half a cache line, four 16-byte loads. Same tree, same child refs as HNode.
//   a = (c0.x c0.y c0.z h0.x)  b = (h0.y h0.z c1.x c1.y)  c = (c1.z h1.x h1.y h1.z)
struct __align__(16) BNode {
    float4 a, b, c;
    int32_t c0, c1;
    uint32_t pad0, pad1;
};
static_assert(sizeof(BNode) == 64, "BNode");

// Quantised form (default): both child boxes on a 15-bit grid spanning the scene bounds, 32 bytes = TWO 16-byte loads per node
// visit. The trace kernel is bound by L1 wavefronts of scattered node loads (ncu: l1tex data-pipe 73 % busy), so bytes per visit
// are what count. Grid: world = qbase + q * qstep per axis, q in [0, 32767]; lo rounded down, hi rounded up (conservative).
//   w[0..2] = child 0: per axis (q_lo | q_hi << 16);  w[3] = child 0 ref;  w[4..6] = child 1;  w[7] = child 1 ref
struct __align__(16) QNode { uint32_t w[8]; };
static_assert(sizeof(QNode) == 32, "QNode");

// 4-wide quantised form (RT_BOUNDS_QBOX4): up to four child boxes on the same 15-bit grid in 64 bytes = FOUR 16-byte loads per node visit,
// from the binary tree by collapsing (a node's children are replaced by their own children, largest box first, until four slots are
// full). Half the dependent node fetches per ray of the binary form -- the latency-bound case once the scene has left L1 / L2.
//   c[k] = (x lo | hi << 16, y lo | hi << 16, z lo | hi << 16, child ref);  empty slot: ref == RT_EMPTY_REF
struct __align__(16) Q4Node { uint4 c[4]; };
static_assert(sizeof(Q4Node) == 64, "Q4Node");
#define RT_EMPTY_REF 0x80000000u      // == RT_DONE of the traversal: never a node index, never a leaf ref

#ifndef RT_LEAF_MAX
#define RT_LEAF_MAX 2            // triangles per cluster (leaf): subtrees of <= RT_LEAF_MAX triangles collapse into one cluster (measured 1 / 2 / 3 / 4 / 6: 2 is fastest)
#endif

RT_DEVICE int leaf_ref(uint32_t first, uint32_t count) { return -(int)(1u + first * 8u + count); }
RT_DEVICE uint32_t leaf_first(int ref) { return ((uint32_t)(-ref) - 1u) >> 3; }
RT_DEVICE uint32_t leaf_count(int ref) { return ((uint32_t)(-ref) - 1u) & 7u; }

// Triangle record for the intersection test, in cluster order: exactly the intermediates the
// reference computes first (raytracer.cpp:85-91), produced with the same unfused arithmetic:
//   r0 = (n.x, n.y, n.z, a.x)  r1 = (a.y, a.z, ab.x, ab.y)  r2 = (ab.z, ac.x, ac.y, ac.z)
// with ab = b - a, ac = c - a, n = Cross(ab, ac).
struct __align__(16) TriRec { float4 r0, r1, r2; };

struct DevMaterial {          // 64 B
    float specular_intensity, index_of_refraction, alpha;
    int32_t tex_ambient;
    float ambient[3];  int32_t tex_diffuse;
    float diffuse[3];  int32_t tex_specular;
    float specular[3]; int32_t tex_alpha;
    int32_t tex_bump; int32_t pad[3];
};
static_assert(sizeof(DevMaterial) == 80, "DevMaterial");

struct DevTexture { uint32_t size_x, size_y, channels, offset; };

struct DevLight {             // scene.h:9-15
    int32_t type;
    float color[3];
    float position[3];
    float facing[3];
    float falloff;
    float pad;
};

struct DevScene {
    const HNode *nodes;            // sphere + slab child bounds (RT_B200_BOUNDS=sphere)
    const BNode *bnodes;           // axis-aligned child bounds, full floats (RT_B200_BOUNDS=box)
    const QNode *qnodes;           // axis-aligned child bounds on the 15-bit scene grid (default)
    const Q4Node *q4nodes;         // same grid, four children per node (RT_B200_BOUNDS=qbox4)
    float qmid[3];                 // qbase - 32768 * qstep per axis (the decode offset, see qbox_ray_setup)
    float qstep[3];
    const TriRec *tris;
    const uint32_t *tri_rank;      // tie-break rank: position in the reference's leaf visit order
    const float4 *tri_uv;          // 2 per triangle: (u0 v0 u1 v1) (u2 v2 material_bits -)
    const float4 *tri_nrm;         // 3 per triangle: vertex normals; the three w's hold Normalize(Cross(ab, ac)) (raytracer.cpp:122)
    const float4 *tri_tan;         // 3 per triangle: vertex tangents (NULL when no bump map)
    const uint8_t *tri_mat;        // material index & 255 per triangle: the sort key of the shading step (grouping only, never a result)
    const uint32_t *tri_vertex0;   // RaycastHit::vertex0 (raytracer.cpp:147)
    const int32_t *tri_object;     // RaycastHit::object as sphere index (raytracer.cpp:148)
    const DevMaterial *materials;  // [n_materials] + default at index n_materials
    const DevTexture *textures;
    const uint8_t *texels;
    const DevLight *lights;
    const float *srgb_lut;         // 256 entries: Color_SRGBToLinear(i / 255) computed by the host libm
    const float4 *hamm_dir;        // 1024 tangent-space cosine-hemisphere directions (raytracer.cpp:322-328)
    const float4 *spec_dir;        // [(n_materials + 1) * spec_samples] Phong-lobe directions (raytracer.cpp:290-300)
    uint32_t n_tris, n_nodes, n_materials, n_lights;
    int32_t root;                  // 0, or a leaf ref when the scene has <= one cluster
    float cull_bound;              // >= |c|_1 + r for every sphere of the hierarchy (see cull_sphere)
};

struct DevParams {
    float ray_bias;
    uint32_t reflection_samples, spec_samples, bounce_depth;
    float bg[3];
    uint32_t pad;
    uint64_t base_seed;
};

#define RT_BOUNDS_SPHERE 0        // child bound of the traversal (rt_trace.cuh): sphere + slab, float boxes, boxes on the 15-bit scene grid
#define RT_BOUNDS_BOX 1
#define RT_BOUNDS_QBOX 2
#define RT_BOUNDS_QBOX4 3
#define RT_STACK_MAX 64           // traversal stack entries per ray; the build guarantees depth + 2 <= RT_STACK_MAX
#define RT_STACK4_MAX 96          // 4-wide form: at most 3 pushes per level; the build guarantees 3 * levels + 2 <= RT_STACK4_MAX

#define RT_SEED_MULT 0x9E3779B97F4A7C15ULL
