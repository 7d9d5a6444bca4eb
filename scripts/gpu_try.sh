#!/bin/bash
# quick GPU experiment: parity tests + phase probe for the kernel variants named on the command line (FEMB200_ROWS values)
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== FEMB200_ROWS=$v"
  FEMB200_ROWS=$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not full_size" 2>&1 | tail -3
  for cfg in tetra4 quad4 hexa8 tetra10; do
    FEMB200_ROWS=$v timeout 300 python scripts/phase_probe.py $cfg 2>&1 | tail -1
  done
done
