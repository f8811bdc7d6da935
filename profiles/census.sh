#!/bin/bash
# Builds the census variant of the library (-DTELR_CENSUS_BUILD=1: cycle counters inside k_al_fused) into telr_b200/_variants/,
# swaps it in for one run of prof_run.py and prints where the alignment kernel's warp cycles go.  Run on a gpurun snapshot.
set -e
N=${1:-296}
CFG=${2:-ont_3k_50x}
mkdir -p telr_b200/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DTELR_CENSUS_BUILD=1 $CENSUS_EXTRA \
     -o telr_b200/_variants/census.so telr_b200/csrc/telr_af.cu
cp telr_b200/_telr_af.so telr_b200/_variants/product.so
cp telr_b200/_variants/census.so telr_b200/_telr_af.so
TELR_CENSUS=1 python profiles/prof_run.py $N $CFG 2>&1 | grep census | awk '!seen[$0]++'
cp telr_b200/_variants/product.so telr_b200/_telr_af.so
