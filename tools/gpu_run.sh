cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -k "deterministic_backward or train_mode or custom_ops or h16_add_layernorm" 2>&1 | tail -25 > gpurun_out/r2p_tests.log
timeout 600 python tools/bench_train.py --steps 5 --warmup 2 > gpurun_out/r2p_train_1gpu.json 2> gpurun_out/r2p_train.err
timeout 900 python bench.py --steps 100 --no-cpu-baseline --kernels 60 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
