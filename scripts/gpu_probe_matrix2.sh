#!/bin/bash
# ring-depth and block-shape matrix for tests/gpu_epi_probe.py; output -> gpurun_out/$1
out=gpurun_out/${1:-probe2.log}
export MMG_PROBE_QUICK=1
P="timeout 120 python tests/gpu_epi_probe.py"
V=$PWD/mmgclip_b200/variants
( $P
  MMGCLIP_B200_LIB=$V/libmmgclip_b200_s5.so $P
  MMGCLIP_B200_LIB=$V/libmmgclip_b200_s4.so $P
  MMGCLIP_B200_BLOCK_ROWS=16384 MMGCLIP_B200_BLOCK_COLS=8192 $P
  MMGCLIP_B200_BLOCK_ROWS=16384 MMGCLIP_B200_BLOCK_COLS=16384 $P
  MMGCLIP_B200_BLOCK_ROWS=32768 MMGCLIP_B200_BLOCK_COLS=8192 $P
  MMGCLIP_B200_BLOCK_ROWS=4096 MMGCLIP_B200_BLOCK_COLS=8192 $P
  MMGCLIP_B200_BLOCK_ROWS=4096 MMGCLIP_B200_BLOCK_COLS=4096 $P
) > $out 2>&1
