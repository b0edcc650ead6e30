#!/bin/bash
# Development helper: link a variant of the library whose FFT length N uses another plan.
#   scripts/mkvariant.sh NAME N "R1, R2, R3, T, W, XR1, XR2, XR3, XT, XL" [extra nvcc flags]
# -> variants/libmvdecon_NAME.so (A/B with scripts/ab_libs.sh on the GPU box)
set -e
NAME=$1; N=$2; PLAN=$3; shift 3
D=build/variants/$NAME; mkdir -p $D
printf '#include "len_ops_impl.cuh"\nMVD_DEFINE_LEN(%s, %s)\n' "$N" "$PLAN" > $D/len_$N.cu
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v -diag-suppress 177,550 \
     -Imultiview-reconstruction_b200/csrc "$@" -c $D/len_$N.cu -o $D/len_$N.o 2> $D/ptxas.log
OBJS=$(ls build/obj/*.o | grep -v "/len_$N.o")
nvcc -shared -o variants/libmvdecon_$NAME.so $OBJS $D/len_$N.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart -ldl
grep -E "registers|spill" $D/ptxas.log | paste - - | awk '{print $5,$9,"|",$16,$17}' | tr '\n' ';'; echo
