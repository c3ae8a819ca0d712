# usage: bash tools/gpu_scale.sh N   (inside a gpurun --gpus N call)
cd $GRAFT_REPO_ROOT
N=$1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" = "1" ]; then RUN="python"; fi
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_scale_kitti_${N}gpu.json 2> gpurun_out/r2v_scale_kitti_${N}gpu.err
timeout 900 $RUN bench.py --gpus $N --config once --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_scale_once_${N}gpu.json 2> gpurun_out/r2v_scale_once_${N}gpu.err
timeout 600 $RUN tools/bench_train.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2v_train_${N}gpu.json 2> gpurun_out/r2v_train_${N}gpu.err
true
