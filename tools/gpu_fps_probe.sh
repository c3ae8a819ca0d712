#!/bin/bash
# FPS rounds: list-length band sweep with the probe (event timing + round statistics)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=${TAG:-fps}
: > gpurun_out/${T}_fps_band.txt
for band in "6 20" "10 28" "8 24" "12 30" "4 14"; do
  set -- $band
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 549 -DPDAB_FPS_TIMERS -DPDAB_FPS_LIST_LO=$1 -DPDAB_FPS_LIST_HI=$2 -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe_t > /tmp/p.log 2>&1 || { tail -5 /tmp/p.log; continue; }
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 549 -DPDAB_FPS_LIST_LO=$1 -DPDAB_FPS_LIST_HI=$2 -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe > /tmp/p.log 2>&1
  echo "== band $1 $2" >> gpurun_out/${T}_fps_band.txt
  timeout 60 /tmp/fps_probe 16 16384 4096 1 0 | tail -1 >> gpurun_out/${T}_fps_band.txt
  timeout 60 /tmp/fps_probe 16 4096 1024 1 0 | tail -1 >> gpurun_out/${T}_fps_band.txt
  timeout 60 /tmp/fps_probe 32 65536 16384 4 0 | tail -1 >> gpurun_out/${T}_fps_band.txt
  timeout 60 /tmp/fps_probe_t 16 16384 4096 1 0 | tail -1 | sed 's/.*| rounds/   rounds/' >> gpurun_out/${T}_fps_band.txt
done
cat gpurun_out/${T}_fps_band.txt
