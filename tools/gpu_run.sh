#!/bin/bash
# One gpurun call of the FPS work: the FPS parity tests, then the probe (plain timing and phase timers).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=${TAG:-r4a}
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "fps" 2>&1 | tail -8 > gpurun_out/${T}_fps_tests.log
cat gpurun_out/${T}_fps_tests.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 549 -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe > /tmp/p.log 2>&1 || { tail -20 /tmp/p.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 549 -DPDAB_FPS_TIMERS -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe_t > /tmp/p.log 2>&1 || { tail -20 /tmp/p.log; exit 1; }
(
for v in 0 1 2 3 4; do timeout 60 /tmp/fps_probe 16 16384 4096 1 $v | tail -1; done
for v in 0 11 12 14 15; do timeout 60 /tmp/fps_probe 16 4096 1024 1 $v | tail -1; done
timeout 60 /tmp/fps_probe 16 16384 1024 1 0 | tail -1
timeout 60 /tmp/fps_probe 16 16384 512 1 0 | tail -1
timeout 60 /tmp/fps_probe 4 65536 4096 4 0 | tail -1
echo "-- phases"
for v in 1 2 3; do timeout 60 /tmp/fps_probe_t 16 16384 4096 1 $v | tail -1; done
for v in 12 14; do timeout 60 /tmp/fps_probe_t 16 4096 1024 1 $v | tail -1; done
) > gpurun_out/${T}_fps_probe.txt 2>&1
cat gpurun_out/${T}_fps_probe.txt
