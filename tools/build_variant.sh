#!/bin/bash
# tools/build_variant.sh NAME "EXTRA nvcc flags" [log2n ...] — developer aid: a copy of libtsb200.so whose chain
# kernels of the given sizes (default: 10) are compiled with extra flags, as build/var/NAME/libtsb200.so (select it
# with TSB200_LIB=...).  Everything else is linked from the regular build.
set -e
name=$1; extra=$2; shift 2 || true
sizes=${@:-10}
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/tinyspec-cling_b200/csrc
out=$root/build/var/$name
mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
FL="$ARCH -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -ffp-contract=off"
objs=""
for o in $src/build/*.o; do
  b=$(basename $o)
  keep=1
  for s in $sizes; do [ "$b" = "chain_$s.o" ] && keep=0; done
  [ $keep = 1 ] && objs="$objs $o"
done
for s in $sizes; do
  nvcc $FL $extra -DTSB_LOG2N=$s -Xptxas -v -c $src/tsb_chain_inst.cu -o $out/chain_$s.o 2> $out/ptxas_$s.log &
done
wait
for s in $sizes; do objs="$objs $out/chain_$s.o"; done
nvcc $ARCH -shared -o $out/libtsb200.so $objs -cudart static
echo built $out/libtsb200.so
