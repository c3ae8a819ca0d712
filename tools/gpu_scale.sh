# usage: bash tools/gpu_scale.sh N [TAG]   (inside a gpurun --gpus N call): KITTI and ONCE bench lines at N GPUs
cd $GRAFT_REPO_ROOT
N=$1
T=${2:-r4y}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" = "1" ]; then RUN="python"; fi
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_scale_kitti_${N}gpu.json 2> gpurun_out/${T}_scale_kitti_${N}gpu.err
timeout 900 $RUN bench.py --gpus $N --config once --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_scale_once_${N}gpu.json 2> gpurun_out/${T}_scale_once_${N}gpu.err
python - <<PY
import json
for c in ("kitti","once"):
    try:
        d=json.loads(open("gpurun_out/${T}_scale_%s_${N}gpu.json"%c).read().strip().split("\n")[-1])
        print(c, "N=$N", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2))
    except Exception as e: print(c, "failed", e)
PY
true
