#!/bin/bash
# stored-E backward under different fused-plan shapes (tests/gpu_stored_e_probe.py time); output -> gpurun_out/$1
out=gpurun_out/${1:-stored_e_matrix.log}
mkdir -p gpurun_out
: > $out
run() { echo "== $*" >> $out; env "$@" PROBE_ONLY=stored timeout 100 python tests/gpu_stored_e_probe.py time 2>&1 | tail -1 >> $out; }
run MMG_NOP=1
run MMG_FUSED_KSL_T=64
run MMG_FUSED_KSL_T=64 MMG_FUSED_CB=4096 MMG_FUSED_KSL=64
run MMG_FUSED_NBUF=6
run MMG_FUSED_RB=2048 MMG_FUSED_KSL_T=32
echo "== recompute" >> $out; PROBE_ONLY=recompute timeout 100 python tests/gpu_stored_e_probe.py time 2>&1 | tail -1 >> $out
cat $out
