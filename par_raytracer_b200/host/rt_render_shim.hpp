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

// ---------------------------------------------------------------------------------------------------------------------
// Device-resident scenes are kept between RenderB200 calls (creating one uploads the mesh and builds the GPU hierarchy: up to
// 0.24 s at 10 M triangles, and each owns a multi-GB path pool), keyed by the reference's Scene pointer. RenderB200Shutdown()
// frees them; the reference itself never frees anything (ownership).
// ---------------------------------------------------------------------------------------------------------------------
struct DeviceSet {
    Scene * key;
    std::vector<rt_scene *> scenes;      // one per GPU this rank drives
    std::vector<rt_comm *> comms;        // local group (CommSize == 1) or one NCCL rank (CommSize > 1)
    bool nccl_ranks;                     // comms[0] spans the MPI ranks
};
inline std::vector<DeviceSet> & DeviceSets() { static std::vector<DeviceSet> v; return v; }

inline void RenderB200Shutdown() {
    std::vector<DeviceSet> & v = DeviceSets();
    for (size_t i = 0; i < v.size(); ++i) {
        for (size_t k = 0; k < v[i].comms.size(); ++k) rt_comm_destroy(v[i].comms[k]);
        for (size_t k = 0; k < v[i].scenes.size(); ++k) rt_scene_destroy(v[i].scenes[k]);
    }
    v.clear();
}

// every rank agrees on failure BEFORE entering a collective (a rank that returned early would hang the others' MPI_Gather /
// ncclReduce): min over ranks of the status codes
inline int AgreeStatus(int rc) {
    int all = rc;
    if (gMPI_CommSize > 1) MPI_Allreduce(&rc, &all, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
    return all;
}

inline DeviceSet * GetDeviceSet(Scene * scene, int device_override) {
    std::vector<DeviceSet> & v = DeviceSets();
    for (size_t i = 0; i < v.size(); ++i) if (v[i].key == scene) return &v[i];
    DeviceSet ds;
    ds.key = scene; ds.nccl_ranks = false;
    int ndev = rt_device_count();
    int rc = ndev > 0 ? RT_OK : RT_ERR_CUDA;
    if (rc != RT_OK) fprintf(stderr, "RenderB200: no CUDA device (there is no CPU fallback)\n");
    FlatScene flat;
    if (rc == RT_OK) Flatten(scene, &flat);
    std::vector<int> devices;
    if (rc == RT_OK) {
        if (gMPI_CommSize > 1 || device_override >= 0) devices.push_back(device_override >= 0 ? device_override : gMPI_CommRank % ndev);   // one GPU per rank, wrapping
        else for (int d = 0; d < ndev; ++d) devices.push_back(d);                // a single rank drives every GPU of the node
        for (size_t k = 0; k < devices.size() && rc == RT_OK; ++k) {
            rt_scene * h = NULL;
            rc = rt_scene_create(&flat.desc, devices[k], &h);
            if (rc != RT_OK) fprintf(stderr, "rt_scene_create (GPU %d) failed: %s\n", devices[k], rt_last_error());
            else ds.scenes.push_back(h);
        }
    }
    if (rc == RT_OK && gMPI_CommSize == 1) {
        ds.comms.resize(devices.size(), NULL);
        rc = rt_comm_create_local((int)devices.size(), devices.data(), ds.comms.data());
        if (rc != RT_OK) { fprintf(stderr, "rt_comm_create_local failed: %s\n", rt_last_error()); ds.comms.clear(); }
    }
    rc = AgreeStatus(rc);
    if (rc == RT_OK && gMPI_CommSize > 1) {
        // NCCL across the MPI ranks: rank 0 draws the id, one MPI_Bcast hands it out. If NCCL is not available the ranks agree to keep
        // the reference's host-side MPI_Gather (a transport choice; the rendering is on the GPU either way).
        uint8_t id[RT_COMM_ID_BYTES];
        memset(id, 0, sizeof(id));
        int have = gMPI_CommRank == 0 ? (rt_comm_unique_id(id) == RT_OK ? 1 : 0) : 1;
        MPI_Bcast(&have, 1, MPI_INT, 0, MPI_COMM_WORLD);
        if (have) {
            MPI_Bcast(id, RT_COMM_ID_BYTES, MPI_BYTE, 0, MPI_COMM_WORLD);
            rt_comm * c = NULL;
            int rcc = rt_comm_create(gMPI_CommSize, gMPI_CommRank, id, devices[0], &c);
            if (rcc != RT_OK) fprintf(stderr, "rt_comm_create failed: %s\n", rt_last_error());
            rc = AgreeStatus(rcc);
            if (c) ds.comms.push_back(c);
            ds.nccl_ranks = rc == RT_OK;
        }
    }
    if (rc != RT_OK) {
        for (size_t k = 0; k < ds.comms.size(); ++k) if (ds.comms[k]) rt_comm_destroy(ds.comms[k]);
        for (size_t k = 0; k < ds.scenes.size(); ++k) rt_scene_destroy(ds.scenes[k]);
        return NULL;
    }
    v.push_back(ds);
    return &v.back();
}

// Defaults = Render's hard-coded adaptive 10..50 samples (main.cpp:308-309); min_samples == max_samples gives the fixed-spp
// mean; base_seed: see rt_params in rt_b200.h. device < 0: GPU = rank % GPUs of the node when there are several MPI ranks, ALL GPUs
// of the node (rt_render_multi) when there is one.
inline Framebuffer RenderB200(Camera * cam, Scene * scene, u32 width, u32 height, u32 min_samples = 10, u32 max_samples = 50,
                              u64 base_seed = 0x835fdd9143716fe3ULL, int device = -1, rt_counters * out_counters = NULL) {
    Framebuffer result;
    result.width = width; result.height = height; result.pixels = NULL;
    static_assert(sizeof(Camera) == sizeof(rt_camera), "Camera layout");

    DeviceSet * ds = GetDeviceSet(scene, device);          // collective; NULL on every rank if any rank failed
    if (!ds) return result;                                // no CPU fallback: an empty frame, reported above
    rt_params params;
    params.ray_bias = gParams.ray_bias;
    params.reflection_samples = gParams.reflection_samples;
    params.spec_samples = gParams.spec_samples;
    params.bounce_depth = gParams.bounce_depth;
    memcpy(params.background_color, &gParams.background_color, 16);
    params.min_samples = min_samples; params.max_samples = max_samples;
    params.base_seed = base_seed;
    const u32 flags = min_samples < max_samples ? RT_FLAG_ADAPTIVE : 0u;
    const u32 tile = 32;

    u32 total_pixel_count = width * height;                                      // main.cpp:311-317
    u32 count_per_proc = (total_pixel_count + gMPI_CommSize - 1) / gMPI_CommSize;
    rt_counters counters;
    memset(&counters, 0, sizeof(counters));
    int rc = RT_OK;
    if (gMPI_CommRank == 0) result.pixels = (Vector4 *)calloc(sizeof(Vector4), (size_t)count_per_proc * gMPI_CommSize);   // main.cpp:338-340

    if (gMPI_CommSize == 1) {
        // one rank: every GPU of the node renders interleaved tiles, gathered through peer memory (or NCCL) inside the library
        MPI_Barrier(MPI_COMM_WORLD);                                             // main.cpp:326-333
        TIME_BLOCK("Render, sync");
        rc = rt_render_multi(ds->scenes.data(), ds->comms.data(), (int)ds->scenes.size(), (const rt_camera *)cam, &params, width, height,
                             RT_PART_TILES, tile, flags, (float *)result.pixels, NULL, NULL, &counters);
        if (rc != RT_OK) fprintf(stderr, "rt_render_multi failed: %s\n", rt_last_error());
        MPI_Barrier(MPI_COMM_WORLD);
    } else if (ds->nccl_ranks) {
        // several ranks, NCCL: interleaved tiles (contiguous ranges load-balance badly, NOTES.txt:25), ncclReduce to rank 0 on the render
        // stream instead of staging every rank's pixels through host memory for MPI_Gather (main.cpp:345-347)
        MPI_Barrier(MPI_COMM_WORLD);
        TIME_BLOCK("Render + reduce, sync");
        rc = rt_render_combined(ds->scenes[0], ds->comms[0], (const rt_camera *)cam, &params, width, height, RT_PART_TILES, tile, flags, 0,
                                (float *)result.pixels, NULL, NULL, &counters);
        if (rc != RT_OK) fprintf(stderr, "rt_render_combined failed: %s\n", rt_last_error());
        MPI_Barrier(MPI_COMM_WORLD);
    } else {
        // several ranks without NCCL: the reference's own contiguous ranges and its MPI_Gather
        u32 start_idx = count_per_proc * gMPI_CommRank;
        u32 count = start_idx < total_pixel_count ? (start_idx + count_per_proc <= total_pixel_count ? count_per_proc : total_pixel_count - start_idx) : 0;
        Vector4 * buffer = (Vector4 *)calloc(sizeof(Vector4), count_per_proc);   // main.cpp:318
        {
            MPI_Barrier(MPI_COMM_WORLD);
            TIME_BLOCK("Render, sync");
            rc = rt_render(ds->scenes[0], (const rt_camera *)cam, &params, width, height, NULL, start_idx, count, 0, min_samples,
                           RT_OUT_MEAN | flags, (float *)buffer, &counters);
            if (rc != RT_OK) fprintf(stderr, "rt_render failed: %s\n", rt_last_error());
            MPI_Barrier(MPI_COMM_WORLD);
        }
        {
            TIME_BLOCK("Reduce");
            MPI_Gather(buffer, count_per_proc * 4, MPI_FLOAT, result.pixels, count_per_proc * 4, MPI_FLOAT, 0, MPI_COMM_WORLD);   // main.cpp:345-347
            MPI_Barrier(MPI_COMM_WORLD);
        }
        free(buffer);
    }
    rc = AgreeStatus(rc);                                                        // a failed rank's zeros are not a frame
    if (rc != RT_OK) { free(result.pixels); result.pixels = NULL; return result; }
    printf("Process %d\n", gMPI_CommRank);                                       // main.cpp:351-356
    printf("Rays cast:          %llu\n", (unsigned long long)counters.ray_count);
    if (out_counters) *out_counters = counters;
    return result;
}

} // namespace rt_b200
