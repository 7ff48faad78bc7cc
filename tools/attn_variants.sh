#!/usr/bin/env bash
# Build variants of one translation unit with -D switches and link each into build/variants/_C_<name>.so; then
# (on a GPU box) time them back to back on the same device:
#   bash tools/attn_variants.sh build  csrc/attn/flash_fwd.cu  base:""  poly4:"-DTB_F2_POLY=4" ...
#   bash tools/attn_variants.sh run    "python benchmarks/attn_bench.py --only native"
set -eu
mode=$1; shift
VD=build/variants; mkdir -p $VD
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -I csrc"
if [ "$mode" = build ]; then
  src=$1; shift
  objname=$(echo "${src#csrc/}" | sed 's#/#__#g; s#\.cu$#.o#')
  python -c "from torchacc_b200.build_native import build; build()" >/dev/null
  others=$(ls build/obj/*.o | grep -v "/$objname$")
  for spec in "$@"; do
    name=${spec%%:*}; defs=${spec#*:}
    ( nvcc $FLAGS $defs -c $src -o $VD/$name.o && nvcc -shared -o $VD/_C_$name.so $VD/$name.o $others -lcudart -gencode arch=compute_100a,code=sm_100a && echo "built $name ($defs)" ) &
  done
  wait
else
  cmd=$1
  for so in $VD/_C_*.so; do
    name=$(basename $so .so); name=${name#_C_}
    printf "%-14s " "$name"
    TORCHACC_B200_NATIVE_LIB=$PWD/$so timeout 120 $cmd 2>&1 | tail -1 | cut -c1-170
  done
fi
