cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "fused or raw_images or ensembles_e5 or batch_json" 2>&1 | tail -40 > gpurun_out/r2d_pytest.txt
tail -8 gpurun_out/r2d_pytest.txt
run() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', round(j['value'],2), round(j['ms_per_step'],1), j['roofline']['ms_per_step_by_kernel'], j['clocks']['sm_mhz'])" ; }
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_a.json 2>gpurun_out/r2d_a.err; run gpurun_out/r2d_a.json fused
POD_TC_DEBUG_NO_RMW=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_b.json 2>gpurun_out/r2d_b.err; run gpurun_out/r2d_b.json fused_no_rmw
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fuse-q1 > gpurun_out/r2d_c.json 2>gpurun_out/r2d_c.err; run gpurun_out/r2d_c.json unfused
