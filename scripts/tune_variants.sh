#!/bin/bash
# Build kernel tuning variants (CPU side) -- each becomes adseismic.jl_b200/libadseis_b200_<tag>.so.
# usage: scripts/tune_variants.sh "tag:-DAC_UA=4 -DAC_MINB_ADJ=1" ...
set -e
cd "$(dirname "$0")/.."
for v in "$@"; do
  tag="${v%%:*}"; flags="${v#*:}"
  ADSEIS_LIB_SUFFIX="_$tag" ADSEIS_NVCC_EXTRA="$flags" python "adseismic.jl_b200/_build.py" --force 2>&1 | grep -A2 "ac_adj_kernel\|ac_fwd_kernel\|el_" | grep -E "spill|Used" | tr '\n' ' '
  echo " <- $tag"
done
