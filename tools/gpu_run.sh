cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2l_tests_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --kernels 60 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
