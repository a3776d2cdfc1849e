#!/bin/bash
# parameter sweep of the diagonal engine on cfg2 (run under gpurun)
for boot in 1 2 4 8; do for slabs in 4 8 16; do for rows in 2048 4096 8192; do
  echo -n "boot=$boot slabs=$slabs rows=$rows "
  K4B_BOOT_TILES=$boot K4B_DIAG_SLABS=$slabs K4B_DIAG_ROWS=$rows python tools/diag_probe.py cfg2 2>&1 | grep '"rep": 1' 
done; done; done
