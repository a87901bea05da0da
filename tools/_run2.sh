cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1e_scale_n2.json 2> gpurun_out/r1e_scale_n2.err
tail -c 400 gpurun_out/r1e_scale_n2.json; tail -3 gpurun_out/r1e_scale_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multi_gpu.py > gpurun_out/r1e_multi_gpu_parity.txt 2>&1
tail -5 gpurun_out/r1e_multi_gpu_parity.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 | tail -c 300
