cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "fused or raw_images or ensembles_e5 or batch_json or unread or chunked or batched_equals" 2>&1 | tail -80 > gpurun_out/r2c_pytest.txt
tail -30 gpurun_out/r2c_pytest.txt
for extra in "" "--no-fuse-q1"; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $extra > gpurun_out/r2c_bench$extra.json 2> gpurun_out/r2c_bench$extra.err
python -c "
import json
j=json.loads(open('gpurun_out/r2c_bench$extra.json').read().strip().splitlines()[-1])
print('[$extra]', j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['ms_per_step_by_kernel'], j['roofline']['achieved'], j['clocks'])" || tail -5 gpurun_out/r2c_bench$extra.err
done
