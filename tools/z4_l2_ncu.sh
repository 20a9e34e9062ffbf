#!/bin/bash
# DRAM bytes of the headline launch shape (256^3, 32 matrices) with and without L2 eviction hints (VT_Z4_L2_HINTS).
set -u
for H in ${HINTS:-0 1 3}; do
  rm -f voltools_b200/csrc/vt_resample_z4.o
  VT_NVCC_EXTRA="-DVT_Z4_L2_HINTS=$H" python voltools_b200/csrc/build.py > /dev/null 2>&1
  echo "== VT_Z4_L2_HINTS=$H"
  python tools/z4_sweep_ncu.py > /dev/null 2>&1 || echo "plain run failed"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:vt_z4_kernel python tools/z4_sweep_ncu.py 2>&1 | grep -E "vt_z4_kernel|dram__|gpu__time|lts__" | sed 's/  */ /g'
done
rm -f voltools_b200/csrc/vt_resample_z4.o
python voltools_b200/csrc/build.py > /dev/null
