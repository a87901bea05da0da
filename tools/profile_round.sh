#!/bin/bash
# One gpurun call that regenerates the judged evidence of a round:  bash tools/profile_round.sh <tag>
#   gpurun_out/<tag>_bench_n1.json / _reference_arm.json   default bench lines (both arms)
#   gpurun_out/<tag>_clocks.csv                            nvidia-smi clocks sampled during the bench
#   gpurun_out/<tag>_launches.csv                          ncu launch list (batch 4, one step)
#   gpurun_out/<tag>_side_workloads.jsonl                  bench lines of the other BASELINE configs (side workloads)
#   gpurun_out/<tag>_tower.ncu-rep, <tag>_out.ncu-rep      ncu --set full of the tower and output convolutions
# Summaries for profiles/ are produced here afterwards by tools/summarize_profiles.py.
TAG=${1:-rX}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --batch 4 --chunk 4 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_conv3x3_tc2$' -s 1 -c 1 -o gpurun_out/${TAG}_tower \
    python bench.py --batch 2 --chunk 2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_conv3x3_wt$' -c 3 -o gpurun_out/${TAG}_out \
    python bench.py --batch 2 --chunk 2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
for w in loss_att baseline ensembles5 bayes_od_mc mc_post; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/${TAG}_side_workloads.jsonl
done
tail -c 600 gpurun_out/${TAG}_bench_n1.json
ls -la gpurun_out | tail -12
