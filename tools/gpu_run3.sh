#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=${TAG:-r4l}
timeout 300 python tools/run_attn_proj.py 524288 32 4 f16 > gpurun_out/${T}_attn_plain.txt 2>&1
timeout 300 python tools/run_attn_proj.py 262144 16 4 f16 >> gpurun_out/${T}_attn_plain.txt 2>&1
cat gpurun_out/${T}_attn_plain.txt
