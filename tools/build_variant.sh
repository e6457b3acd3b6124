#!/bin/bash
# Builds an A/B variant of the CUDA library with extra nvcc -D flags:
#   tools/build_variant.sh <name> "<extra flags>"   ->  snark-setup_b200/csrc/variants/libss_<name>.so
# Select it at run time with SS_LIB=<path>.
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d /tmp/ssvar_XXXX)
mkdir -p $tmp/snark-setup_b200 $tmp/include
cp -r $root/snark-setup_b200/csrc $tmp/snark-setup_b200/
cp $root/include/*.h $tmp/include/
cd $tmp/snark-setup_b200/csrc && rm -f *.o *.so
make -j8 NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -diag-suppress 20091 $extra" > build.log 2>&1 || { tail -20 build.log; exit 1; }
mkdir -p $root/snark-setup_b200/csrc/variants
cp libsnarksetup_b200.so $root/snark-setup_b200/csrc/variants/libss_$name.so
grep -h -A2 "k_scalar_mul" kern_bls377_g1.ptxas.log kern_bls377_g2.ptxas.log | grep -E "registers|spill" | paste - - | sed "s/^/$name: /" | cut -c1-220
rm -rf $tmp
