#!/bin/bash
# ncu launch lists (durations only) of a single IPA proof and of one multiproof over 2^12 openings.
#   gpurun -- bash tools/profile_single.sh
set -e
mkdir -p gpurun_out
for t in ipa_single multiproof; do
  python tools/${t}_trace.py
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${t}_launches.csv \
      python tools/${t}_trace.py > gpurun_out/${t}_ncu.log 2>&1
done
python tools/launch_summary.py gpurun_out/ipa_single_launches.csv k_barycentric | tee gpurun_out/ipa_single_summary.txt
python tools/launch_summary.py gpurun_out/multiproof_launches.csv k_powers | tee gpurun_out/multiproof_summary.txt
