#!/bin/bash
# tools/build_variant.sh <name> <file.cu> [-D...]: builds optical_flow_b200/lib/libofb200_<name>.so with ONE translation
# unit recompiled under extra flags (A/B experiments on a GPU box: OFB_LIB_PATH=.../libofb200_<name>.so python bench.py).
set -e
cd "$(dirname "$0")/../optical_flow_b200/csrc"
name=$1; unit=$2; shift 2
make -s
objs=""
for f in pyramid polyexp matrices blur_solve iter viz preprocess jpeg engine; do
  if [ "$f.cu" = "$unit" ]; then
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC \
         -Xcompiler -fvisibility=hidden "$@" -c $f.cu -o ../../build/obj/${f}_$name.o
    objs="$objs ../../build/obj/${f}_$name.o"
  else
    objs="$objs ../../build/obj/$f.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libofb200_$name.so $objs -Xcompiler -fPIC
echo built ../lib/libofb200_$name.so
