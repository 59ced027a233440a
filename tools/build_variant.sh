#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..." : libgi2d variant with extra defines for gi2d_fit.cu / gi2d_raster.cu
# (A/B timing on the GPU box: GI2D_LIB=<path> python bench.py ...).  Output: gaussianimage_plus_b200/csrc/build/libgi2d_NAME.so
set -e
cd "$(dirname "$0")/../gaussianimage_plus_b200/csrc"
NAME=$1; FLAGS=$2
mkdir -p build/v_$NAME
for f in gi2d_fit gi2d_raster; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $FLAGS -c -o build/v_$NAME/$f.o $f.cu
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libgi2d_$NAME.so build/gi2d_common.o build/gi2d_project.o build/gi2d_binning.o build/gi2d_loss.o build/gi2d_quant.o build/v_$NAME/gi2d_fit.o build/v_$NAME/gi2d_raster.o
echo built build/libgi2d_$NAME.so
