cd $GRAFT_REPO_ROOT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -diag-suppress 549 -I pdanet_b200/csrc -I include tools/fps_phase_probe.cu -o /tmp/fps_probe > /tmp/p.log 2>&1 || exit 1
timeout 60 /tmp/fps_probe 32 65536 16384 4 0 | tail -1
timeout 300 ncu --set full --clock-control none -k regex:fps_pruned -s 1 -c 1 -o /tmp/fpsc /tmp/fps_probe 32 65536 16384 4 0 > gpurun_out/r5g_ncu.log 2>&1
ncu -i /tmp/fpsc.ncu-rep --page details > gpurun_out/r5g_fps_cluster_details.txt 2>/dev/null
grep -E "Duration|Issue Slots Busy|Executed Ipc Active|Registers Per|Dynamic Shared|Achieved Occupancy|Cluster|Grid Size|DRAM Throughput" gpurun_out/r5g_fps_cluster_details.txt | head -12
