#!/bin/bash
# Builds a development variant of the library with extra ray-tracer flags:
#   scripts/build_variant.sh t128 "-DC2B_RT_THREADS=128 -DC2B_CTA_PER_SM=4"   ->  gpurun_variants/libc2ray_b200_t128.so
# Use it with C2B_LIB=gpurun_variants/libc2ray_b200_t128.so python bench.py ...
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/gpurun_variants; mkdir -p $out/obj_$name
cd $root/c2ray3dm_b200/csrc
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC"
$NV $flags -c raytrace.cu -o $out/obj_$name/raytrace.o
for f in chemistry tables c2b_api; do [ -f $f.o ] || make -s $f.o; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $out/libc2ray_b200_$name.so $out/obj_$name/raytrace.o chemistry.o tables.o c2b_api.o -ldl
echo built $out/libc2ray_b200_$name.so
