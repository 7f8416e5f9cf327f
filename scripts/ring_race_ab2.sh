#!/bin/bash
# second A/B round for the depth-8 forward ring (see ring_race_ab.sh): instrumented build + proxy-fence variants +
# depth 4 at depth 8's occupancy
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REPS=${REPS:-40}
for v in _nst8dbg _nst8pf1 _nst8pf2 _nst4big; do
  echo "== variant '$v'" | tee -a gpurun_out/ring_ab2.log
  ADSEIS_LIB_SUFFIX=$v REPS=$REPS timeout 900 python scripts/determinism_probe.py 2>&1 | tail -40 | tee -a gpurun_out/ring_ab2.log
done
