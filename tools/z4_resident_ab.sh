#!/bin/bash
# A/B of the cubic_tex slice4 kernel's register cap (resident 128-thread CTAs per SM): rebuilds vt_resample_z4.cu on the
# GPU box with -DVT_Z4_CT8_RESIDENT=N and times the cubic_tex launches.   usage (under gpurun): bash tools/z4_resident_ab.sh
set -u
for R in ${RESIDENTS:-4 5 6}; do
  rm -f voltools_b200/csrc/vt_resample_z4.o
  VT_NVCC_EXTRA="-DVT_Z4_CT8_RESIDENT=$R" python voltools_b200/csrc/build.py -v 2>&1 | grep -A2 "vt_z4_kernelILi1ELi0ELb1ELi4ELi8" | grep -E "registers|spill"
  echo "== VT_Z4_CT8_RESIDENT=$R"
  python tools/z4_probe.py 256 512 --quick --no-parity 2>&1 | grep "axis 0 cubic_tex\|sweep of 30 angles cubic_tex"
done
rm -f voltools_b200/csrc/vt_resample_z4.o
python voltools_b200/csrc/build.py > /dev/null
