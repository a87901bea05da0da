cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -70 > gpurun_out/r2e_pytest.txt
tail -12 gpurun_out/r2e_pytest.txt
run() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', round(j['value'],2), round(j['e2e']['value'],2), round(j['ms_per_step'],1), j['roofline']['ms_per_step_by_kernel'], j['clocks']['sm_mhz'])" ; }
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_a.json 2>gpurun_out/r2e_a.err; run gpurun_out/r2e_a.json fused
POD_TC_DEBUG_NO_RMW=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_b.json 2>gpurun_out/r2e_b.err; run gpurun_out/r2e_b.json fused_no_rmw
