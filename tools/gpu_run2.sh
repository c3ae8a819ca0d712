#!/bin/bash
# ncu capture of the FPS probe (one launch) + setup-time measurement
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/ncu
T=${TAG:-r4d}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -diag-suppress 549 -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe > /tmp/p.log 2>&1 || { tail -20 /tmp/p.log; exit 1; }
(
timeout 60 /tmp/fps_probe 16 16384 2 1 0 | tail -1
timeout 60 /tmp/fps_probe 16 16384 64 1 0 | tail -1
timeout 60 /tmp/fps_probe 16 16384 256 1 0 | tail -1
timeout 60 /tmp/fps_probe 16 16384 4096 1 0 | tail -1
timeout 60 /tmp/fps_probe 16 4096 2 1 0 | tail -1
timeout 60 /tmp/fps_probe 16 4096 1024 1 0 | tail -1
) > gpurun_out/${T}_fps_setup.txt 2>&1
cat gpurun_out/${T}_fps_setup.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:fps_pruned -s 2 -c 1 -o /tmp/ncu/fps /tmp/fps_probe 16 16384 4096 1 0 > gpurun_out/${T}_ncu.log 2>&1
ncu -i /tmp/ncu/fps.ncu-rep --page raw --csv > gpurun_out/${T}_fps_raw.csv 2>/dev/null
ncu -i /tmp/ncu/fps.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_fps_source.csv 2>/dev/null
ncu -i /tmp/ncu/fps.ncu-rep --page details > gpurun_out/${T}_fps_details.txt 2>/dev/null
ls -la gpurun_out/${T}_*
