cd $GRAFT_REPO_ROOT
for cfg in "32 3" "16 6" "16 4" "8 8"; do
  set -- $cfg
  timeout 600 python bench.py --config once --batch $1 --depth $2 --steps 12 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('once batch $1 depth $2', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3))" >> gpurun_out/r2y_once_tune.log
done
