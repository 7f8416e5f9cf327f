#!/bin/bash
# Next-round recipe (needs a B200): build the forward kernel with ring depth 8 as a side-by-side variant and look for
# the hazard behind the non-deterministic repeat runs.
#   gpurun --timeout 900 -- scripts/racecheck_ring8.sh
set -e
cd "$(dirname "$0")/.."
ADSEIS_LIB_SUFFIX=_nst8 ADSEIS_NVCC_EXTRA="-DAC_NST_FWD=8" python adseismic.jl_b200/_build.py --force > /dev/null
export ADSEIS_LIB_SUFFIX=_nst8
mkdir -p gpurun_out
# 1. how often do repeat runs differ (small grid first: cheap under the sanitizer)
REPS=20 python scripts/determinism_probe.py | tail -3 | tee gpurun_out/ring8_probe.log
# 2. shared-memory hazards of the TMA ring (racecheck tracks cp.async.bulk / mbarrier since CUDA 12.3)
PNX=1024 PNY=1024 PT=12 REPS=2 timeout 600 compute-sanitizer --tool racecheck --racecheck-report all \
    python scripts/determinism_probe.py 2>&1 | tail -40 | tee gpurun_out/ring8_racecheck.log
# 3. barrier misuse
PNX=1024 PNY=1024 PT=12 REPS=2 timeout 600 compute-sanitizer --tool synccheck \
    python scripts/determinism_probe.py 2>&1 | tail -20 | tee gpurun_out/ring8_synccheck.log
