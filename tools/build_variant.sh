#!/bin/bash
# usage: tools/build_variant.sh <name> [nvcc flags...]  ->  gpurun_out/variants/libgpr_b200_<name>.so  (select with GPRB_LIB=...)
# Kernel A/B experiments: same sources, extra -D flags; objects are built in a scratch directory.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/gpr_calculator_b200/csrc
out=$root/variants; mkdir -p $out /tmp/gprb_var_$name
for f in pack cov_mma cov_ee gp_linalg so3 peer; do
  if [ $f = cov_mma ] || [ ! -f /tmp/gprb_var_base/$f.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o /tmp/gprb_var_$name/$f.o $src/$f.cu &
  fi
done
wait
mkdir -p /tmp/gprb_var_base
if [ "$name" != base ]; then for f in pack cov_ee gp_linalg so3 peer; do if [ -f /tmp/gprb_var_$name/$f.o ]; then cp /tmp/gprb_var_$name/$f.o /tmp/gprb_var_base/$f.o; fi; done; fi
objs=""; for f in pack cov_mma cov_ee gp_linalg so3 peer; do if [ -f /tmp/gprb_var_$name/$f.o ]; then objs="$objs /tmp/gprb_var_$name/$f.o"; else objs="$objs /tmp/gprb_var_base/$f.o"; fi; done
nvcc -shared -o $out/libgpr_b200_$name.so $objs -L/usr/local/cuda/lib64 -lcusolver -lcublas -Xlinker -rpath -Xlinker /usr/local/cuda/lib64 2>/dev/null
ls -la $out/libgpr_b200_$name.so
