#!/bin/bash
# pipeline depth x reserved SMs sweep of the default bench (device-resident and e2e scenes/s)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=${TAG:-sweep}
: > gpurun_out/${T}_depth_reserve.txt
for cfg in "8 16" "8 0" "8 8" "6 0" "12 0" "12 16" "16 0" "8 24"; do
  set -- $cfg
  timeout 300 python bench.py --no-cpu-baseline --steps 120 --depth $1 --reserve-sms $2 > /tmp/b.json 2>/tmp/b.err || { echo "depth $1 reserve $2 FAILED" >> gpurun_out/${T}_depth_reserve.txt; continue; }
  python - "$1" "$2" >> gpurun_out/${T}_depth_reserve.txt <<'PY'
import json,sys
d=json.loads(open('/tmp/b.json').read().strip().split('\n')[-1])
print('depth',sys.argv[1],'reserve',sys.argv[2],round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'clk',d['clocks']['sm_mhz'])
PY
done
cat gpurun_out/${T}_depth_reserve.txt
