cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2m_tests_gpu.log
timeout 900 python bench.py --kernels 60 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 900 python tools/bench_ops.py > gpurun_out/r2m_ops_sweep.jsonl 2> gpurun_out/r2m_ops_sweep.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err
