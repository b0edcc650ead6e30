#!/bin/bash
mkdir -p gpurun_out
L=multiview-reconstruction_b200/libmvdecon.so
bash scripts/ab_libs.sh c3 2 $L variants/libmvdecon_c3018.so variants/libmvdecon_c2720.so variants/libmvdecon_c1830.so variants/libmvdecon_xl4.so variants/libmvdecon_xl2.so $L 2>&1 | tee gpurun_out/ab_b.txt
