#!/bin/bash
# Developer A/B builds: tools/build_variant.sh NAME "-DARVC_NBQ=2 ..."  ->  lidar_slam_arvc_b200/csrc/build/variants/NAME.so
# (select at run time with ARVC_LIB_VARIANT=<path>)
set -e
cd "$(dirname "$0")/../lidar_slam_arvc_b200/csrc"
name=$1; shift
out=build/variants/$name
mkdir -p $out
for f in api preprocess normals normals_blk icp map plane; do
  extra=""
  case $f in normals|normals_blk|plane) extra="-fmad=false";; esac
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $extra "$@" -c $f.cu -o $out/$f.o 2> $out/$f.ptxas.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $out/*.o -lcudart
grep -h "k_normals_blk" -A3 $out/normals_blk.ptxas.log | grep Used
