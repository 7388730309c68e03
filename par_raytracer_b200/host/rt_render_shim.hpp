// rt_render_shim.hpp -- the host side of the drop-in, in the reference's own language (C++11).
//
// Include this AFTER the reference's translation unit (main.cpp / raytracer.cpp) so that its types are
// visible: Scene, SceneObject, Mesh, MeshGroup, Material, Texture, BoundingHierarchy, Camera, Framebuffer,
// gParams, gMPI_CommRank / gMPI_CommSize. It provides
//
//     Framebuffer RenderB200(Camera * cam, Scene * scene, u32 width, u32 height);      // <-> Render, main.cpp:301-358
//
// with the signature and the MPI behaviour of the reference's Render: every rank renders its contiguous pixel
// range [count_per_proc * rank, count_per_proc * (rank + 1)) (main.cpp:311-317) -- on its GPU instead of its CPU
// core -- and the ranges are gathered on rank 0 with the same MPI_Gather (main.cpp:345-347). Everything else of the
// reference (InitParams, ParseOBJ, CalculateTangents, BuildHierarchy, InitScene, WriteFramebufferImage) is untouched.
//
// The shim only flattens pointers into the POD arrays of include/rt_b200.h; all rendering happens behind the C ABI.
#pragma once
#include <map>
#include <vector>
#include <cstring>
#include <cstdio>

#include "rt_b200.h"

namespace rt_b200 {

struct FlatScene {
    std::vector<u32> group_first, idx_p, idx_t, idx_n;
    std::vector<int32_t> group_material, sphere_group;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    std::map<Material *, int> material_index;
    std::map<Texture *, int> texture_index;
    rt_scene_desc desc;
};

inline int TextureId(FlatScene * f, Texture * t) {
    if (!t) return -1;
    std::map<Texture *, int>::iterator it = f->texture_index.find(t);
    if (it != f->texture_index.end()) return it->second;
    rt_texture rt;
    rt.size_x = t->size_x; rt.size_y = t->size_y; rt.channels = t->channels; rt.texels = t->texels;
    int id = (int)f->textures.size();
    f->textures.push_back(rt);
    f->texture_index[t] = id;
    return id;
}

inline void CopyMaterial(FlatScene * f, rt_material * o, Material * m) {
    o->specular_intensity = m->specular_intensity;
    o->index_of_refraction = m->index_of_refraction;
    o->alpha = m->alpha;
    memcpy(o->ambient_color, &m->ambient_color, 16);
    memcpy(o->diffuse_color, &m->diffuse_color, 16);
    memcpy(o->specular_color, &m->specular_color, 16);
    memcpy(o->emissive_color, &m->emissive_color, 16);
    o->ambient_texture = TextureId(f, m->ambient_texture);
    o->diffuse_texture = TextureId(f, m->diffuse_texture);
    o->specular_texture = TextureId(f, m->specular_texture);
    o->alpha_texture = TextureId(f, m->alpha_texture);
    o->bump_texture = TextureId(f, m->bump_texture);
}

inline int MaterialId(FlatScene * f, Material * m) {
    if (!m) return -1;
    std::map<Material *, int>::iterator it = f->material_index.find(m);
    if (it != f->material_index.end()) return it->second;
    rt_material rm;
    CopyMaterial(f, &rm, m);
    int id = (int)f->materials.size();
    f->materials.push_back(rm);
    f->material_index[m] = id;
    return id;
}

// Scene (scene.h:29-36) + BoundingHierarchy (bsphere.cpp:322-326) + Mesh (mesh.h:46-55) -> rt_scene_desc
inline void Flatten(Scene * scene, FlatScene * f) {
    BoundingHierarchy * h = scene->hierarchy;
    Mesh * mesh = h->mesh;
    static_assert(sizeof(Vector3) == 12 && sizeof(Vector2) == 8, "vector layout");
    static_assert(sizeof(BoundingSphere) == sizeof(rt_bsphere), "BoundingSphere layout");
    static_assert(sizeof(LightSource) == sizeof(rt_light), "LightSource layout");
    size_t G = mesh->groups.size();
    f->group_first.assign(G + 1, 0);
    f->group_material.assign(G, -1);
    for (size_t g = 0; g < G; ++g) {
        MeshGroup & mg = mesh->groups[g];
        f->group_first[g + 1] = f->group_first[g] + (u32)mg.idx_positions.size();
        f->idx_p.insert(f->idx_p.end(), mg.idx_positions.begin(), mg.idx_positions.end());
        f->idx_t.insert(f->idx_t.end(), mg.idx_texcoords.begin(), mg.idx_texcoords.end());
        f->idx_n.insert(f->idx_n.end(), mg.idx_normals.begin(), mg.idx_normals.end());
    }
    f->sphere_group.assign(h->spheres.size(), -1);
    for (size_t i = 0; i < h->mesh_groups.size(); ++i) {
        MeshGroup * mg = h->mesh_groups[i];
        if (!mg) continue;
        int32_t g = (int32_t)(mg - &mesh->groups[0]);
        f->sphere_group[i] = g;
        // the material the integrator reads is SceneObject::material (main.cpp:586-589), default_mat when the group has none
        Material * m = scene->objects[i]->material;
        f->group_material[g] = (m == scene->default_mat) ? -1 : MaterialId(f, m);
    }
    rt_scene_desc & d = f->desc;
    memset(&d, 0, sizeof(d));
    d.n_positions = (u32)mesh->positions.size(); d.positions = (const float *)mesh->positions.data();
    d.n_texcoords = (u32)mesh->texcoords.size(); d.texcoords = (const float *)mesh->texcoords.data();
    d.n_normals = (u32)mesh->normals.size();     d.normals = (const float *)mesh->normals.data();
    d.tangents = mesh->tangents.size() == mesh->normals.size() ? (const float *)mesh->tangents.data() : NULL;
    d.n_groups = (u32)G;
    d.group_first = f->group_first.data();
    d.idx_positions = f->idx_p.data(); d.idx_texcoords = f->idx_t.data(); d.idx_normals = f->idx_n.data();
    d.group_material = f->group_material.data();
    d.n_spheres = (u32)h->spheres.size();
    d.spheres = (const rt_bsphere *)h->spheres.data();
    d.sphere_group = f->sphere_group.data();
    CopyMaterial(f, &d.default_material, scene->default_mat);
    d.n_materials = (u32)f->materials.size(); d.materials = f->materials.data();
    d.n_textures = (u32)f->textures.size();   d.textures = f->textures.data();
    d.n_lights = scene->light_count;          d.lights = (const rt_light *)scene->lights;
}

// Defaults = Render's hard-coded adaptive 10..50 samples (main.cpp:308-309); min_samples == max_samples gives the fixed-spp
// mean; base_seed: see rt_params in rt_b200.h.
inline Framebuffer RenderB200(Camera * cam, Scene * scene, u32 width, u32 height, u32 min_samples = 10, u32 max_samples = 50,
                              u64 base_seed = 0x835fdd9143716fe3ULL, int device = -1, rt_counters * out_counters = NULL) {
    Framebuffer result;
    result.width = width; result.height = height; result.pixels = NULL;
    static_assert(sizeof(Camera) == sizeof(rt_camera), "Camera layout");

    FlatScene flat;
    Flatten(scene, &flat);
    rt_scene * handle = NULL;
    if (device < 0) device = gMPI_CommRank;             // one GPU per rank; wraps below if the node has fewer
    int rc = rt_scene_create(&flat.desc, device, &handle);
    if (rc == RT_ERR_ARG && device > 0) rc = rt_scene_create(&flat.desc, 0, &handle);
    if (rc != RT_OK) {
        fprintf(stderr, "rt_scene_create failed: %s\n", rt_last_error());     // no CPU fallback: report and return an empty frame
        return result;
    }
    rt_params params;
    params.ray_bias = gParams.ray_bias;
    params.reflection_samples = gParams.reflection_samples;
    params.spec_samples = gParams.spec_samples;
    params.bounce_depth = gParams.bounce_depth;
    memcpy(params.background_color, &gParams.background_color, 16);
    params.min_samples = min_samples; params.max_samples = max_samples;
    params.base_seed = base_seed;

    u32 total_pixel_count = width * height;                                      // main.cpp:311-317
    u32 count_per_proc = (total_pixel_count + gMPI_CommSize - 1) / gMPI_CommSize;
    u32 start_idx = count_per_proc * gMPI_CommRank;
    u32 count = start_idx < total_pixel_count ? (start_idx + count_per_proc <= total_pixel_count ? count_per_proc : total_pixel_count - start_idx) : 0;
    Vector4 * buffer = (Vector4 *)calloc(sizeof(Vector4), count_per_proc);       // main.cpp:318
    rt_counters counters;
    memset(&counters, 0, sizeof(counters));
    {
        MPI_Barrier(MPI_COMM_WORLD);                                             // main.cpp:326-333
        TIME_BLOCK("Render, sync");
        rc = rt_render(handle, (const rt_camera *)cam, &params, width, height, NULL, start_idx, count, 0, min_samples,
                       RT_OUT_MEAN | (min_samples < max_samples ? RT_FLAG_ADAPTIVE : 0u), (float *)buffer, &counters);
        if (rc != RT_OK) fprintf(stderr, "rt_render failed: %s\n", rt_last_error());
        MPI_Barrier(MPI_COMM_WORLD);
    }
    if (gMPI_CommRank == 0) result.pixels = (Vector4 *)calloc(sizeof(Vector4), (size_t)count_per_proc * gMPI_CommSize);   // main.cpp:338-340
    {
        TIME_BLOCK("Reduce");
        MPI_Gather(buffer, count_per_proc * 4, MPI_FLOAT, result.pixels, count_per_proc * 4, MPI_FLOAT, 0, MPI_COMM_WORLD);   // main.cpp:345-347
        MPI_Barrier(MPI_COMM_WORLD);
    }
    printf("Process %d\n", gMPI_CommRank);                                       // main.cpp:351-356
    printf("Rays cast:          %llu\n", (unsigned long long)counters.ray_count);
    if (out_counters) *out_counters = counters;
    free(buffer);
    rt_scene_destroy(handle);
    return result;
}

} // namespace rt_b200
