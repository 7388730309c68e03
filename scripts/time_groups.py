import sys, time; sys.path.insert(0,'.')
from par_raytracer_b200 import api, scenes
for cells, block in ((724, 32), (2236, 32)):
    sd = scenes.heightfield_scene(cells, cells, block=block, size=400.0, amp=20.0, textured=False)
    t = time.time(); sp, sg = api.build_group_hierarchy(sd); dt = time.time() - t
    t = time.time(); sp, sg = api.build_group_hierarchy(sd); dt2 = time.time() - t
    print(f"{sd.n_groups} groups x {sd.n_triangles // sd.n_groups} tris: GPU BuildHierarchy {dt:.2f}s (2nd call {dt2:.2f}s), {len(sp)} spheres")
