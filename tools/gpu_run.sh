cd $GRAFT_REPO_ROOT
# ONCE config at BASELINE configs[2] size: batch 32 x 65536 points, with the CPU arm
timeout 1200 python bench.py --config once --kernels 40 > gpurun_out/r2u_bench_once_b32.json 2> gpurun_out/r2u_bench_once.err
# training step, 1 GPU
timeout 600 python tools/bench_train.py --steps 10 --warmup 3 > gpurun_out/r2u_train_1gpu.json 2> gpurun_out/r2u_train.err
# launch list of the default bench command (short)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_ncu.log 2>&1
tail -2 gpurun_out/r2u_bench_once.err
