#!/bin/bash
# usage: tools/ablate.sh CIN COUT H W  -- prints the autotune table of one layer under each SB_ABLATE mask
for ab in 0 1 2 4 8 15; do
  echo "ABLATE $ab shapes=${SB_HALO_SHAPES:-default}"
  SB_ABLATE=$ab SB_DEBUG=1 timeout 120 python tools/prof_layer.py $1 $2 3 $3 $4 8 1 2>&1 | grep "sb_conv_tc\]" | tail -1 | sed 's/.*persist/persist/'
done
