#!/bin/bash
# Builds mitsuba-path-guiding_b200/_variants/libb200pg_<name>.so with extra -D flags for kernels.cu / volpath.cu
# (A/B experiments on the GPU box: B200PG_LIB=<path> python bench.py ...). usage: tools/build_variant.sh name -DPG_X=1 ...
set -e
name=$1; shift
cd "$(dirname "$0")/../mitsuba-path-guiding_b200/csrc"
make -s >/dev/null
out=../_variants; obj=_obj/var_$name
mkdir -p $out $obj
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp,-O3 --expt-relaxed-constexpr -Xptxas -v"
for f in kernels volpath; do $NV "$@" -c $f.cu -o $obj/$f.o 2> $obj/$f.ptxas.log & done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libb200pg_$name.so $obj/kernels.o $obj/volpath.o _obj/integrator.o _obj/guiding.o _obj/host_scene.o _obj/xml_scene.o -Xcompiler -fopenmp -lgomp -ldl -lz -cudart shared
grep -A2 "k_shadeENS" $obj/kernels.ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " -> $out/libb200pg_$name.so"
