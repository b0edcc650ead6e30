#!/bin/bash
# Development helper: link a variant of the library whose FFT length N uses another plan.
#   scripts/mkvariant.sh NAME N "R1, R2, R3, T, W, XR1, XR2, XR3, XT, XL" [extra nvcc flags]
# -> variants/libmvdecon_NAME.so (A/B with scripts/ab_libs.sh on the GPU box).  The variant only carries the small lengths the PSF
# derivation uses plus N (the snapshot that travels to the GPU box is limited to 512 MiB).
set -e
NAME=$1; N=$2; PLAN=$3; shift 3
SUB="32 36 40 48 50 54 60 64 72 80 90 96 100 128"
D=build/variants/$NAME; mkdir -p $D/gen variants
MVD_LENGTHS="$SUB $N" python3 multiview-reconstruction_b200/csrc/gen_lengths.py $D/gen > /dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v -diag-suppress 177,550 -Imultiview-reconstruction_b200/csrc"
printf '#include "len_ops_impl.cuh"\nMVD_DEFINE_LEN(%s, %s)\n' "$N" "$PLAN" > $D/len_$N.cu
nvcc $FLAGS "$@" -c $D/len_$N.cu -o $D/len_$N.o 2> $D/ptxas.log
nvcc $FLAGS -x cu -c $D/gen/registry.cpp -o $D/registry.o 2> /dev/null
OBJS=""
for n in $SUB; do [ "$n" != "$N" ] && OBJS="$OBJS build/obj/len_$n.o"; done
for o in engine pointwise comm psf_prep tiff_io n5_io capi; do OBJS="$OBJS build/obj/$o.o"; done
nvcc -shared -o variants/libmvdecon_$NAME.so $OBJS $D/registry.o $D/len_$N.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart -ldl -lz
grep -E "registers|spill" $D/ptxas.log | paste - - | awk '{print $5,$9,"|",$16,$17}' | tr '\n' ';'; echo
