cd $GRAFT_REPO_ROOT
timeout 300 python tools/dbg_ffn.py > gpurun_out/r2t_dbg.log 2>&1
