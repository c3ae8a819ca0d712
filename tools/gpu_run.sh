cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "ffn_fused" 2>&1 | tail -30 > gpurun_out/r2q_tests_ffn.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -k "pda_fast_path or teacher or golden" 2>&1 | tail -8 > gpurun_out/r2q_tests_model.log
timeout 600 python bench.py --steps 100 --no-cpu-baseline --kernels 60 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
