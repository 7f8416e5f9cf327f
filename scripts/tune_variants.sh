#!/bin/bash
# Build kernel tuning variants (CPU side) -- each becomes adseismic.jl_b200/libadseis_b200_<tag>.so.
# usage: scripts/tune_variants.sh "tag:-DAC_UA=4 -DAC_MINB_ADJ=1" ...
set -e
cd "$(dirname "$0")/.."
for v in "$@"; do
  tag="${v%%:*}"; flags="${v#*:}"
  ADSEIS_LIB_SUFFIX="_$tag" ADSEIS_NVCC_EXTRA="$flags" python "adseismic.jl_b200/_build.py" --force >/dev/null 2>&1
  echo "== $tag ($flags)"
  grep -A2 "Compiling entry function '_Z1[0-9]*\(ac_\|el_\)" adseismic.jl_b200/build.log | grep -E "Compiling|spill|Used" | sed -e "s/.*function '_Z[0-9]*\([a-z_0-9]*kernel\|el_[a-z_]*\).*/\1/" | paste - - - | awk '{print "   ", $0}' | sed -e 's/ptxas info    ://' -e 's/bytes stack frame/B stack/' | cut -c1-200
done
