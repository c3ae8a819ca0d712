cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_model.py -q 2>&1 | tail -12 > gpurun_out/r2w_tests_model.log
timeout 600 python bench.py --steps 100 --no-cpu-baseline --kernels 70 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
