#!/bin/bash
# Full GPU check of the tree: the -m gpu suite, then the default bench line (and optionally the ONCE line).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=${TAG:-full}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_tests_gpu.log
cat gpurun_out/${T}_tests_gpu.log
timeout 900 python bench.py --kernels 80 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python tools/show_bench.py gpurun_out/${T}_bench.json 2>/dev/null | head -40 || head -c 1500 gpurun_out/${T}_bench.json
if [ -n "$ONCE" ]; then
  timeout 900 python bench.py --config once --no-cpu-baseline > gpurun_out/${T}_once.json 2> gpurun_out/${T}_once.err
  head -c 700 gpurun_out/${T}_once.json
fi
