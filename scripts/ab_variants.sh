#!/bin/bash
# Time every library under _variants/ (plus the in-tree default) on the same box: bash scripts/ab_variants.sh [orders] [sigmas]
for L in elasticdeform_b200/libedf_b200.so _variants/*.so; do
  EDF_B200_LIB=$PWD/$L python scripts/ab_time.py "${1:-3}" "${2:-8}" ${3:-4}
done
