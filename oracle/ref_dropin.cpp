// ref_dropin.cpp -- TEST INFRASTRUCTURE ONLY: the drop-in demonstrated on the UNMODIFIED reference.
//
// One translation unit = the reference's main.cpp (included by path, `main` renamed) + the product's host shim
// (par_raytracer_b200/host/rt_render_shim.hpp), linked against librt_b200.so. dropin_render() does what the
// reference's main() does (main.cpp:544-602) with the single change INTEGRATION.md describes: the call to
// Render(...) becomes RenderB200(...). Built by oracle/Makefile into oracle/_ref/libref_dropin.so.
#include <string>
#include <cstring>

#define main reference_main
#include "main.cpp"
#undef main

#include "rt_render_shim.hpp"

// min_spp < max_spp: RenderB200's default mode, the reference's adaptive sampling (main.cpp:308-309: 10 .. 50).
static int g_dropin_device = 0;      // -1: RenderB200's default -- a single rank drives EVERY GPU of the node (rt_render_multi)
extern "C" void dropin_set_device(int device) { g_dropin_device = device; }

extern "C" int dropin_render_adaptive(const char *dir, u32 width, u32 height, float fov, const float *cam_pos, const float *cam_facing,
                                      u32 min_spp, u32 max_spp, u64 base_seed, float *out_rgba, unsigned long long *out_rays) {
    char *argv0[] = { (char *)"ref", nullptr };
    InitParams(1, argv0);                                                        // main.cpp:544
    gParams.image_width = width; gParams.image_height = height; gParams.camera_fov = fov;
    gParams.camera_position = Vector3(cam_pos[0], cam_pos[1], cam_pos[2]);
    gParams.camera_facing = Vector3(cam_facing[0], cam_facing[1], cam_facing[2]);
    gMPI_CommSize = 1; gMPI_CommRank = 0;
    Camera cam = MakeCamera(gParams.camera_fov, gParams.image_width, gParams.image_height);   // main.cpp:546
    Matrix33 transform; transform.SetIdentity();
    char *d = strdup(dir); char fname[] = "sponza.obj";
    Mesh *mesh = ParseOBJ(d, fname, transform);                                  // main.cpp:553
    free(d);
    if (!mesh) return -1;
    CalculateTangents(mesh);                                                     // main.cpp:557
    BoundingHierarchy hierarchy;
    BuildHierarchy(&hierarchy, mesh);                                            // main.cpp:573
    Scene scene = InitScene();                                                   // main.cpp:577-599
    scene.hierarchy = &hierarchy;
    scene.default_mat = MakeMaterial(Vector4(0.75f, 0.5f, 0.75f, 1.0f));
    for (u32 i = 0; i < hierarchy.mesh_groups.size(); ++i) {
        MeshGroup *mg = hierarchy.mesh_groups[i];
        SceneObject *obj = (SceneObject *)calloc(1, sizeof(SceneObject));
        obj->mesh_group = mg; obj->mesh = mesh; obj->type = ObjectType_MeshGroup;
        obj->material = scene.default_mat;
        if (mg && mg->material) obj->material = mg->material;
        scene.objects.push_back(obj);
    }
    rt_counters counters;
    Framebuffer fb = rt_b200::RenderB200(&cam, &scene, gParams.image_width, gParams.image_height, min_spp, max_spp, base_seed, g_dropin_device, &counters);   // <-> main.cpp:602
    rt_b200::RenderB200Shutdown();                                               // the Scene above lives on this stack frame: drop the cached device copy
    if (!fb.pixels) return -2;
    memcpy(out_rgba, fb.pixels, (size_t)width * height * 16);
    if (out_rays) *out_rays = counters.ray_count;
    free(fb.pixels);
    return 0;
}

extern "C" int dropin_render(const char *dir, u32 width, u32 height, float fov, const float *cam_pos, const float *cam_facing,
                             u32 spp, u64 base_seed, float *out_rgba, unsigned long long *out_rays) {
    return dropin_render_adaptive(dir, width, height, fov, cam_pos, cam_facing, spp, spp, base_seed, out_rgba, out_rays);
}
