#!/bin/bash
# repeat the C2 parity test: a forward that is not bit-reproducible shows up as an occasional
# failure (ReLU / LeakyReLU masks of elements at rounding-noise distance from 0)
n=${1:-6}; f=0
for i in $(seq $n); do
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "c2_mosei" >/dev/null 2>&1
  c=$(grep -c "^FAIL" gpurun_out/parity_c2_mosei_b256.txt)
  [ "$c" != "0" ] && f=$((f+1))
done
echo "c2 parity: $f / $n runs failed"
