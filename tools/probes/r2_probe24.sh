#!/bin/bash
# round-2 probe 24: where do the 21.7 us of the 3x3 256->256 @40x40 layer go?  (tile fixed at BN=128 MT=2; DBG surgery)
O=gpurun_out/probe24; mkdir -p $O
export ADAS_B200_AUTOTUNE=0 ADAS_B200_BN=128 ADAS_B200_MT=2
for d in 0 64 48 18 16 34 4 8; do ADAS_B200_DBG=$d timeout 300 python tools/layer_bench.py "P4 3x3 256->256 40x40" 2>&1 | grep "dbg="; done | tee $O/decomp_40x40.txt
export ADAS_B200_BN=128 ADAS_B200_MT=1
for d in 0 48 18 16; do ADAS_B200_DBG=$d timeout 300 python tools/layer_bench.py "P4 3x3 256->256 40x40" 2>&1 | grep "dbg="; done | tee $O/decomp_40x40_mt1.txt
export ADAS_B200_BN=64 ADAS_B200_MT=1
for d in 0 48 18 16; do ADAS_B200_DBG=$d timeout 300 python tools/layer_bench.py "P5 3x3 256->256 20x20" 2>&1 | grep "dbg="; done | tee $O/decomp_20x20.txt
