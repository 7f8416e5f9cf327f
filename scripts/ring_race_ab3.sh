#!/bin/bash
# confirmation round for the ring fix (fence.proxy.async between the producer's acquire of `empty` and the refill):
# fenced vs unfenced builds in the SAME session, at the two configurations that failed (ring depth 8; depth 4 with the
# shared-memory footprint of depth 8, i.e. 2 CTAs per SM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in _nst8nf _nst4bignf; do
  echo "== unfenced control '$v'" | tee -a gpurun_out/ring_ab3.log
  ADSEIS_LIB_SUFFIX=$v REPS=${CREPS:-80} timeout 900 python scripts/determinism_probe.py 2>&1 | grep -v "^run " | tail -3 | tee -a gpurun_out/ring_ab3.log
done
for v in _nst8f _nst4bigf ""; do
  echo "== fenced '$v'" | tee -a gpurun_out/ring_ab3.log
  ADSEIS_LIB_SUFFIX=$v REPS=${REPS:-250} timeout 1200 python scripts/determinism_probe.py 2>&1 | grep -v "^run " | tail -3 | tee -a gpurun_out/ring_ab3.log
done
echo "== fenced, gradient (adjoint ring)" | tee -a gpurun_out/ring_ab3.log
PGRAD=1 PT=60 REPS=60 timeout 900 python scripts/determinism_probe.py 2>&1 | grep -v "^run " | tail -3 | tee -a gpurun_out/ring_ab3.log
