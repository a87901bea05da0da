cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py -q -x 2>&1 | tail -40 > gpurun_out/r2f_pytest_bb.txt
tail -25 gpurun_out/r2f_pytest_bb.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "raw_images" 2>&1 | tail -15
run() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', round(j['value'],2), round(j['e2e']['value'],2), round(j['ms_per_step'],1), j['roofline']['ms_per_step_by_kernel'], j['clocks']['sm_mhz'])" ; }
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2f_a.json 2>gpurun_out/r2f_a.err; run gpurun_out/r2f_a.json fused
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-layers --no-fuse-q1 > gpurun_out/r2f_c.json 2>gpurun_out/r2f_c.err; run gpurun_out/r2f_c.json unfused
