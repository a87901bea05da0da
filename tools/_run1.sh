cd $GRAFT_REPO_ROOT
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/t_full.log; cat gpurun_out/t_full.log
for k in "" "--keep-unread"; do
timeout 300 python bench.py --no-cpu-baseline $k > gpurun_out/b1.json 2> gpurun_out/b1.err; tail -3 gpurun_out/b1.err
python - <<PY
import json
j=json.loads(open("gpurun_out/b1.json").read().strip().splitlines()[-1])
print("$k", j["value"], j["e2e"]["value"], j["roofline"]["mma_frac"], j["roofline"]["ms_per_step_by_kernel"], j["clocks"], j["gpu_launches"])
PY
done
timeout 300 python bench.py --no-cpu-baseline --workload ensembles5 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ens5', j['value'], j['e2e']['value'])"
