#!/bin/bash
# A-B library: tools/build_ab.sh <out.so> <file.cu>[,<file.cu>...] [-DFLAG ...]
# recompiles the named csrc files with the extra flags and links them with the in-tree objects of everything else
# (run the normal build first).  Load with TVIT_LIB_PATH=<out.so>.
set -e
out=$1; files=$2; shift 2
cd "$(dirname "$0")/../neural_vit_b200"
tmp=$(mktemp -d)
objs=""
for o in build/*.o; do
  b=$(basename $o .o)
  if [[ ",$files," == *",$b.cu,"* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$b.cu -o $tmp/$b.o &
    objs="$objs $tmp/$b.o"
  else
    objs="$objs $o"
  fi
done
wait
nvcc -shared -o "$OLDPWD/$out" $objs -lcudart
rm -rf $tmp
echo "built $out"
