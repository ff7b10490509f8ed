#!/bin/bash
# builds engine variants with different -D settings into vp8oclenc_b200/_variants/ (kernel tuning only)
# usage: tools/build_variants.sh name "-DA=1 -DB=2" [name flags ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p vp8oclenc_b200/_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  objs=""
  for s in me_kernels transform_kernels loopfilter_kernels capi_misc engine mb_fused_kernel; do
    extra=""; case $s in transform_kernels|mb_fused_kernel) extra="-fmad=false";; esac
    if [ $s = me_kernels ] || [ $s = mb_fused_kernel ] || [ ! -f vp8oclenc_b200/_obj/$s.cu.o ]; then
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I include -I vp8oclenc_b200/csrc $extra $flags -Xptxas -v -c vp8oclenc_b200/csrc/$s.cu -o /tmp/var_$s.o 2>&1 | grep -A2 "k_luma_search\|k_mb_fused" | grep -E "Used|spill" | sed "s/^/  [$name] /"
      objs="$objs /tmp/var_$s.o"
    else
      objs="$objs vp8oclenc_b200/_obj/$s.cu.o"
    fi
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o vp8oclenc_b200/_variants/lib_$name.so $objs -lcudart
done
ls vp8oclenc_b200/_variants/
