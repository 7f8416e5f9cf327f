#!/bin/bash
# A/B experiment for the forward-ring non-determinism seen in round 1 at ring depth 8.
# Variants are built on the CPU side first (they travel with the snapshot):
#   scripts/tune_variants.sh "nst8:-DAC_NST_FWD=8 -DADSEIS_NO_SMEM_PAD" "nst8pad:-DAC_NST_FWD=8" "nst4nopad:-DADSEIS_NO_SMEM_PAD"
# then on the GPU box:   gpurun --timeout 1500 -- scripts/ring_race_ab.sh
# A = depth 8 with the packed round-1 stage layout (two bulk copies share a 128-byte shared-memory line),
# B = depth 8 with every bulk-copy destination padded to whole 128-byte lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REPS=${REPS:-40}
for v in _nst8 _nst8pad _nst4nopad ""; do
  if [ -f "adseismic.jl_b200/libadseis_b200$v.so" ]; then
    for pdl in 1 0; do
      echo "== variant '$v' PDL=$pdl" | tee -a gpurun_out/ring_ab.log
      ADSEIS_PDL=$pdl ADSEIS_LIB_SUFFIX=$v REPS=$REPS timeout 600 python scripts/determinism_probe.py 2>&1 | tail -6 | tee -a gpurun_out/ring_ab.log
    done
  fi
done
# shared-memory hazards / barrier misuse of the TMA ring on a small grid (cheap under the sanitizer)
for tool in racecheck synccheck; do
  echo "== compute-sanitizer $tool (variant _nst8)" | tee -a gpurun_out/ring_ab.log
  ADSEIS_LIB_SUFFIX=_nst8 PNX=1024 PNY=1100 PT=10 REPS=2 timeout 500 compute-sanitizer --tool $tool \
      python scripts/determinism_probe.py 2>&1 | tail -25 | tee -a gpurun_out/ring_ab.log
done
