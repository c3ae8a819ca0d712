cd $GRAFT_REPO_ROOT
for cfg in "8 16" "12 16" "6 16" "8 8" "8 0" "12 8" "16 16"; do
  set -- $cfg
  timeout 300 python bench.py --steps 100 --no-cpu-baseline --depth $1 --reserve-sms $2 2> /dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('depth $1 reserve $2', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3))" >> gpurun_out/r2x_tune.log
done
