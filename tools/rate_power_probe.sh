#!/bin/bash
# Sustained L2-hot rate of the production GEMM pipeline (sb_selftest_mma_rate, ~150 ms per run) with constant operands
# and with operands that have the statistics of a real null: the tensor pipe's power draw depends on the data, and the
# chip runs at its power cap.  A background nvidia-smi samples the clock and the power while the runs go.
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 100 > gpurun_out/rate_power_smi.txt &
SMI=$!
for rep in 1 2; do
  echo "== constant operands (A = 0x55.., B = 1)"; python tools/gpu_rate.py 65536 148 192 | grep -E "production|MMAs only"
  for fill in 33 50 100; do
    echo "== random digits, masks ${fill} % filled"; SB_RATE_RANDOM=$fill python tools/gpu_rate.py 65536 148 192 | grep -E "production|MMAs only"
  done
done
kill $SMI
sort gpurun_out/rate_power_smi.txt | uniq -c | sort -k1,1nr | head -12
