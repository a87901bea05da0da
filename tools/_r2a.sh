cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nproc > gpurun_out/r2a_nproc.txt; lscpu | head -20 >> gpurun_out/r2a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
timeout 300 python tools/bayesod_envelope.py > gpurun_out/r2a_bayesod_envelope.txt 2> gpurun_out/r2a_bayesod_envelope.err
cat gpurun_out/r2a_bayesod_envelope.txt | tail -20
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python -c "
import json
j=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['ms_per_step_by_kernel'], j['clocks'])"
