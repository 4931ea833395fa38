#!/bin/bash
# round-2 probe 31: what the 6.5 us epilogue of a 256x128 tile is made of (DBG 34 = epilogue only; + 256 no TMEM read, + 512 no staging
# writes, + 1024 no chunk barrier, + 8 no activation, + 4 no TMA store)
O=gpurun_out/probe31; mkdir -p $O
export ADAS_B200_AUTOTUNE=0 ADAS_B200_BN=128 ADAS_B200_MT=2
for d in 34 290 546 1058 42 38 1834 1838; do ADAS_B200_DBG=$d timeout 300 python tools/layer_bench.py "P4 3x3 256->256 40x40" 2>&1 | grep "dbg=" | grep -v B32; done | tee $O/epilogue_decomp.txt
