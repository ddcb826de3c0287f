#!/bin/bash
# tools/build_spec_variant.sh NAME "EXTRA nvcc flags" [source.cu] — developer aid: libtsb200.so with tsb_spectral.cu (or
# another source put in its place) compiled with extra flags, as build/var/NAME/libtsb200.so (TSB200_LIB=...)
set -e
name=$1; extra=$2; srcf=${3:-tinyspec-cling_b200/csrc/tsb_spectral.cu}
root=$(cd "$(dirname "$0")/.." && pwd); cd $root
out=build/var/$name; mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
cp $srcf tinyspec-cling_b200/csrc/_var_${name}_spectral.cu
nvcc $ARCH -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -ffp-contract=off $extra -c tinyspec-cling_b200/csrc/_var_${name}_spectral.cu -o $out/spectral.o
rm tinyspec-cling_b200/csrc/_var_${name}_spectral.cu
objs=$(ls tinyspec-cling_b200/csrc/build/*.o | grep -v spectral.o)
nvcc $ARCH -shared -o $out/libtsb200.so $objs $out/spectral.o -cudart static
echo built $out/libtsb200.so
