cd $GRAFT_REPO_ROOT
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/t_full.log; cat gpurun_out/t_full.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
