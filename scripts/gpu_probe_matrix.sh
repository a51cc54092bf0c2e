#!/bin/bash
# Runs tests/gpu_epi_probe.py under the epilogue-variant matrix (on the GPU box, via gpurun); output -> gpurun_out/$1
out=gpurun_out/${1:-probe.log}
P="timeout 120 python tests/gpu_epi_probe.py"
( $P; MMG_EPI_WARPS=8 $P; MMG_EPI_DBG=1 $P; MMG_EPI_DBG=2 $P; MMG_EPI_DBG=3 $P; MMG_EPI_WARPS=8 MMG_EPI_DBG=3 $P; MMG_TC_DUAL_SPLIT=0 $P ) > $out 2>&1
