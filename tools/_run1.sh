cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_conv3x3_tc$' -c 3 -o gpurun_out/out_conv_v2 python bench.py --batch 2 --chunk 2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_out2.log 2>&1
tail -2 gpurun_out/ncu_out2.log | cut -c1-300
