cd $GRAFT_REPO_ROOT
for args in "ln_f32res 1500 256 1" "ln_split 1500 256 1" "ln_split 1500 512 1" "ln_split 150000 256 0" "pool_split 1600 256 1" "pool_split 160000 256 0"; do
  timeout 120 python tools/dbg_ln.py $args 2>&1 | tail -1 >> gpurun_out/r2f_dbg.log
done
timeout 900 python -m pytest tests/test_gpu_tc.py -q -k "h16" 2>&1 | tail -40 > gpurun_out/r2f_tests_h16.log
timeout 600 python -m pytest tests/test_gpu_model.py -q 2>&1 | tail -15 > gpurun_out/r2f_tests_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tc-passes 4 --kernels 60 > gpurun_out/r2f_bench_p4.json 2> gpurun_out/r2f_bench_p4.err
