#!/bin/bash
# round-2 probe 3: why are cta_group::2 pairs slower?  v2 kernel, forced 256-wide tiles, single vs pair, full / loads-only / MMA-only
O=gpurun_out/probe3; mkdir -p $O
for mc in 0 2; do for dbg in 0 18 48 16; do for bn in 256 128; do
  echo "== MC=$mc DBG=$dbg BN=$bn" >> $O/pair.txt
  ADAS_B200_GEMM=v2 ADAS_B200_MC=$mc ADAS_B200_DBG=$dbg ADAS_B200_AUTOTUNE=0 ADAS_B200_BN=$bn ADAS_B200_MT=1 python tools/layer_bench.py >> $O/pair.txt 2>&1
done; done; done
for mc in 0 2; do for dbg in 0 18 48; do
  echo "== MC=$mc DBG=$dbg BN=128 MT=2" >> $O/pair.txt
  ADAS_B200_GEMM=v2 ADAS_B200_MC=$mc ADAS_B200_DBG=$dbg ADAS_B200_AUTOTUNE=0 ADAS_B200_BN=128 ADAS_B200_MT=2 python tools/layer_bench.py >> $O/pair.txt 2>&1
done; done
cat $O/pair.txt | grep -E "==|P4 3x3 256->256 40x40  |P3 1x1|P3 3x3 128->128 80x80  "
