#!/bin/bash
# Builds one library per setting of the fused loop's tuning knobs (kernels_quad.cuh) next to the product library:
#   bash profiles/exp/quad_variants.sh name1:"-DKNOB=1 ..." name2:"..."   ->  groan_rs_b200/libquad_<name>.so
# Only groan_gpu.cu includes kernels_quad.cuh; the other objects are taken from build/obj (run groan_rs_b200/build.py first).
set -e
cd "$(dirname "$0")/../.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-pthread -rdc=true"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  ( nvcc $FLAGS $defs -c groan_rs_b200/csrc/groan_gpu.cu -o build/obj/quad_$name.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -rdc=true -lcudadevrt -o groan_rs_b200/libquad_$name.so \
      build/obj/quad_$name.o $(ls build/obj/groan_*.o | grep -v groan_gpu.o) && echo built $name ) &
done
wait
