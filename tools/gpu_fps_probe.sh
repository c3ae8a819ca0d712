#!/bin/bash
# One gpurun call of the FPS work: the FPS parity tests, then the probe (plain timing and phase timers).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=${TAG:-r4a}
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "fps" 2>&1 | tail -8 > gpurun_out/${T}_fps_tests.log
cat gpurun_out/${T}_fps_tests.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 549 -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe > /tmp/p.log 2>&1 || { tail -20 /tmp/p.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 549 -DPDAB_FPS_TIMERS -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe_t > /tmp/p.log 2>&1 || { tail -20 /tmp/p.log; exit 1; }
(
timeout 60 /tmp/fps_probe 16 16384 4096 1 0 | tail -1
timeout 60 /tmp/fps_probe 16 4096 1024 1 0 | tail -1
echo "-- clusters: 4 x 65536 -> 4096 on CL 4, steps (variant 20) then rounds"
timeout 60 /tmp/fps_probe 4 65536 4096 4 20 | tail -1
timeout 60 /tmp/fps_probe 4 65536 4096 4 0 | tail -1
echo "-- 32 x 65536 -> 16384 on CL 4 (ONCE L0), steps then rounds"
timeout 120 /tmp/fps_probe 32 65536 16384 4 20 | tail -1
timeout 120 /tmp/fps_probe 32 65536 16384 4 0 | tail -1
echo "-- 8 x 65536 -> 16384 on CL 8; 4 x 262144 -> 512 on CL 16"
timeout 120 /tmp/fps_probe 8 65536 16384 8 0 | tail -1
timeout 120 /tmp/fps_probe 4 262144 512 16 0 | tail -1
echo "-- phases"
timeout 60 /tmp/fps_probe_t 4 65536 4096 4 0 | tail -1
timeout 60 /tmp/fps_probe_t 32 65536 16384 4 0 | tail -1
) > gpurun_out/${T}_fps_probe.txt 2>&1
cat gpurun_out/${T}_fps_probe.txt
