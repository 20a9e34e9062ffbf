#!/bin/bash
# A/B of L2 eviction hints in the slice4 kernels: rebuilds vt_resample_z4.cu on the GPU box with -DVT_Z4_L2_HINTS=N
# (bit 0: TMA loads evict_last, bit 1: streaming output stores) and times the probe.  usage (under gpurun): bash tools/z4_l2_ab.sh
set -u
for H in ${HINTS:-0 1 2 3}; do
  rm -f voltools_b200/csrc/vt_resample_z4.o
  VT_NVCC_EXTRA="-DVT_Z4_L2_HINTS=$H" python voltools_b200/csrc/build.py > /dev/null 2>&1
  echo "== VT_Z4_L2_HINTS=$H"
  python tools/z4_probe.py 256 512 --quick --no-parity 2>&1 | grep "axis 0 \|sweep of 30 angles"
done
rm -f voltools_b200/csrc/vt_resample_z4.o
python voltools_b200/csrc/build.py > /dev/null
