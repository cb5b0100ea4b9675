#!/bin/bash
# DRAM traffic of one config-5 K1 launch under different L2 plans (ncu, two metrics only)
for cfg in "24 2" "40 2" "56 2" "80 2"; do
  set -- $cfg
  JEGAL_CHUNK_MB=$1 JEGAL_C_POLICY=$2 PROFILE_REPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:simpool_kernel --csv python scripts/profile_targets.py k1 2>/dev/null | grep -E "simpool" | awk -F'","' -v c="$1/$2" '{print c, $(NF-2), $(NF-1), $NF}'
done
