#!/bin/bash
# usage: scripts/ac_dd_sweep.sh NGPU "PDL:RB PDL:RB ..."   -> gpurun_out/ac_dd_sweep.log
N=$1; shift
mkdir -p gpurun_out
for v in $1; do
  pdl=${v%%:*}; rb=${v##*:}
  export ADSEIS_PDL=$pdl
  if [ "$rb" = "auto" ]; then unset ADSEIS_AC_RB; else export ADSEIS_AC_RB=$rb; fi
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/ac_dd_probe.py 2>&1 | grep "acoustic DD" | tee -a gpurun_out/ac_dd_sweep.log
done
