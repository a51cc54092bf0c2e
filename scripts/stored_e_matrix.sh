#!/bin/bash
# fused backward (stored-E and recompute) under different L2 footprints (block shape x super-tile x buffers); -> gpurun_out/$1
out=gpurun_out/${1:-stored_e_matrix.log}
mkdir -p gpurun_out
: > $out
run() { mode=$1; shift; echo "== $mode $*" >> $out; env "$@" PROBE_ONLY=$mode PROBE_REPS=5 timeout 100 python tests/gpu_stored_e_probe.py time 2>&1 | tail -1 >> $out; }
for mode in stored recompute; do
  run $mode MMG_FUSED_RB=2048 MMG_FUSED_CB=2048 MMG_FUSED_NBUF=3 MMG_FUSED_SR=2 MMG_FUSED_SC=2
  run $mode MMG_FUSED_RB=2048 MMG_FUSED_CB=1024 MMG_FUSED_NBUF=4 MMG_FUSED_SR=2 MMG_FUSED_SC=4
  run $mode MMG_FUSED_RB=1024 MMG_FUSED_CB=2048 MMG_FUSED_NBUF=4 MMG_FUSED_SR=4 MMG_FUSED_SC=2
  run $mode MMG_FUSED_RB=4096 MMG_FUSED_CB=2048 MMG_FUSED_NBUF=3 MMG_FUSED_SR=1 MMG_FUSED_SC=2
done
cat $out
