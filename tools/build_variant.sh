#!/bin/bash
# Tuning build: tools/build_variant.sh <name> <source.cu> [-DFLAG=...]  ->  gpurun_out/variants/lib_<name>.so
# (objects of the other sources come from imagine_b200/csrc/build/; run `python -m imagine_b200._build` first).
# Use with IMB_LIB_PATH=gpurun_out/variants/lib_<name>.so python tools/los_probe.py
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
out=variants; mkdir -p $out
cs=imagine_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $cs/$src -o $out/${name}_${src%.cu}.o
objs=""
for f in api fields los los_rl likelihood gram_batch dense comm; do
  if [ "$f.cu" == "$src" ]; then objs="$objs $out/${name}_${src%.cu}.o"; else objs="$objs $cs/build/$f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out/lib_$name.so $objs
echo $out/lib_$name.so
