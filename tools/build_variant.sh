#!/bin/bash
# build an A/B variant of the CUDA library: tools/build_variant.sh NAME "<extra nvcc flags for exact.cu>" "<extra flags for gather.cu>"
# -> many-lights-gi_b200/lib/variants/NAME.so  (timed by tools/ab.sh through MLGI_LIB_PATH)
set -e
NAME=$1; EX=$2; GA=$3
cd "$(dirname "$0")/../many-lights-gi_b200"
mkdir -p build/var_$NAME lib/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off"
$NV -fmad=false $EX -c csrc/exact.cu -o build/var_$NAME/exact.o
$NV -ftz=true $GA -c csrc/gather.cu -o build/var_$NAME/gather.o
$NV -c csrc/api.cu -o build/var_$NAME/api.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o lib/variants/$NAME.so build/var_$NAME/exact.o build/var_$NAME/gather.o build/var_$NAME/api.o
echo built lib/variants/$NAME.so
