#!/bin/bash
# builds a variant of librt_b200.so with extra -D flags for rt_render.cu (the unit holding the trace / shading kernels):
#   scripts/build_variant.sh <name> -DFOO=1 ...   ->  par_raytracer_b200/librt_b200_<name>.so   (select at run time with RT_B200_LIB)
set -e
name=$1; shift
cd "$(dirname "$0")/../par_raytracer_b200"
python -m build >/dev/null 2>&1 || (cd .. && python -m par_raytracer_b200.build >/dev/null)
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-strict-aliasing"
nvcc $F "$@" -c -o build/rt_render_$name.o csrc/rt_render.cu
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o librt_b200_$name.so build/rt_scene.o build/rt_render_$name.o build/rt_loadtime.o build/rt_comm.o -ldl
echo librt_b200_$name.so
