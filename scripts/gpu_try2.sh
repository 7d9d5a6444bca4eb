#!/bin/bash
# usage: gpu_try2.sh "<cfgs>" "<env settings 1>" "<env settings 2>" ...   each env string like "FEMB200_ROWS=fused FEMB200_SROW_TPB=64"
cfgs=$1; shift
for v in "$@"; do
  echo "=== $v"
  for cfg in $cfgs; do
    env $v timeout 300 python scripts/phase_probe.py $cfg 2>&1 | tail -1 | sed 's/upload_validate=[0-9.]* //; s/merge=[0-9.]* //'
  done
done
