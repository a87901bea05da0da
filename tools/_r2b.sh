cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2b_pytest.txt
tail -15 gpurun_out/r2b_pytest.txt
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python -c "
import json
j=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['ms_per_step_by_kernel'], j['clocks'], j.get('cpu_baseline'))"
