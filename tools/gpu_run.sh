cd $GRAFT_REPO_ROOT
mkdir -p /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2z_tests_gpu.log
timeout 900 python bench.py --kernels 80 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
timeout 300 python tools/run_step.py --steps 2 > gpurun_out/r2z_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -k regex:"tc_gemm_kernel|ffn_fused|sa_fused_pair|pda_encode|group_attention_h|fps_pruned|grid_build|nms_|post_|topk" -s 60 -c 60 -o /tmp/ncu/step python tools/run_step.py --steps 2 > gpurun_out/r2z_ncu.log 2>&1
ncu -i /tmp/ncu/step.ncu-rep --page raw --csv > gpurun_out/r2z_step_raw.csv 2> gpurun_out/r2z_export.err
ls -la /tmp/ncu >> gpurun_out/r2z_ncu.log
