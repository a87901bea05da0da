#!/bin/bash
# One gpurun call that regenerates the judged evidence of a round:  bash tools/profile_round.sh <tag>
#   gpurun_out/<tag>_bench_n1.json / _reference_arm.json   default bench lines (both arms)
#   gpurun_out/<tag>_clocks.csv                            nvidia-smi clocks sampled during the bench
#   gpurun_out/<tag>_launches.csv                          ncu launch list (batch 4, one step)
#   gpurun_out/<tag>_{tower,out,hbm,backbone,post}.ncu-rep ncu --set full captures
#   gpurun_out/<tag>_side_workloads.jsonl                  bench lines of the other BASELINE configs (side workloads)
#   gpurun_out/<tag>_sanitizer_*.txt                       compute-sanitizer memcheck / racecheck / synccheck
#   gpurun_out/<tag>_{bayesod_envelope,mma_shape,q1_experiments}.txt
# Summaries for profiles/ are produced here afterwards by tools/summarize_profiles.py.
TAG=${1:-rX}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
G=gpurun_out/${TAG}
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > ${G}_pytest_gpu.txt
tail -3 ${G}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${G}_smoke.txt 2>&1; tail -1 ${G}_smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > ${G}_clocks.csv &
SMI=$!
timeout 900 python bench.py > ${G}_bench_n1.json 2> ${G}_bench_n1.err
kill $SMI
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > ${G}_reference_arm.json 2> ${G}_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${G}_launches.csv \
    python bench.py --steps 1 --warmup 0 --batch 4 --chunk 4 --no-cpu-baseline > /dev/null 2>&1
NCU="ncu --set full --clock-control none"
timeout 600 $NCU -k 'regex:^k_conv3x3_tc2$' -s 1 -c 1 -o ${G}_tower python bench.py --batch 2 --chunk 2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 $NCU -k 'regex:^k_conv3x3_wt$' -c 3 -o ${G}_out python bench.py --batch 2 --chunk 2 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 $NCU -k 'regex:^(k_q1_mean_act|k_mask_expand|k_sample_mean_q1_v4)$' -c 3 -o ${G}_hbm python bench.py --batch 4 --chunk 4 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 $NCU -k 'regex:^(k_conv3x3_tc|k_stem_conv7|k_maxpool3s2)$' -s 2 -c 8 -o ${G}_backbone python bench.py --workload loss_att --from-images --batch 4 --chunk 4 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 $NCU -k 'regex:^(k_decode|k_nms_fuse|k_cluster_merge|k_scores|k_topk)$' -c 10 -o ${G}_post python bench.py --workload mc_post --n-mc 30 --batch 4 --chunk 4 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
rm -f ${G}_side_workloads.jsonl
for w in "loss_att --batch 8 --chunk 8" loss_att baseline ensembles5 bayes_od_mc mc_post \
         "loss_att --from-images --batch 8 --chunk 8" "loss_att --from-images" "loss_att --from-images --backbone torch" \
         "baseline --from-images" "ensembles5 --from-images" \
         "loss_att --from-images --batch 8 --chunk 8 --cuda-graph" "loss_att --from-images --chunk 32 --cuda-graph" \
         "baseline --from-images --chunk 32 --cuda-graph" "loss_att --batch 8 --chunk 8 --cuda-graph"; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | tail -1 >> ${G}_side_workloads.jsonl
done
S="compute-sanitizer --print-limit 20"
( echo "# ${TAG}: $S --tool memcheck python -m pytest tests/test_gpu_kernels.py -q -m gpu -k 'conv or mean or mask or truncation'"
  timeout 900 $S --tool memcheck python -m pytest tests/test_gpu_kernels.py -q -m gpu -k 'conv or mean or mask or truncation' 2>&1 | tail -8
  echo "# ${TAG}: $S --tool memcheck python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_backbone.py -q -m gpu -k 'unread or chunked or end_to_end_matches_oracle or fused or wire or general_convolution or stem_pool'"
  timeout 1200 $S --tool memcheck python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_backbone.py -q -m gpu -k 'unread or chunked or end_to_end_matches_oracle or fused or wire or general_convolution or stem_pool' 2>&1 | tail -8 ) > ${G}_sanitizer_memcheck.txt
( echo "# ${TAG}: $S --tool racecheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -m gpu -k 'weights_as_a or hidden_dropout or more_samples_than_one_group'"
  timeout 1200 $S --tool racecheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -m gpu -k 'weights_as_a or hidden_dropout or more_samples_than_one_group' 2>&1 | tail -25
  echo "# ${TAG}: $S --tool racecheck ./tools/_racecheck_tmem_alloc2.bin   (minimal CTA-pair TMEM allocation, tools/racecheck_tmem_alloc2.cu)"
  timeout 300 $S --tool racecheck ./tools/_racecheck_tmem_alloc2.bin 2>&1 | tail -25 ) > ${G}_sanitizer_racecheck.txt
( echo "# ${TAG}: $S --tool synccheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -m gpu -k 'weights_as_a or hidden_dropout or more_samples_than_one_group or nms_edge'"
  timeout 1200 $S --tool synccheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -m gpu -k 'weights_as_a or hidden_dropout or more_samples_than_one_group or nms_edge' 2>&1 | tail -25 ) > ${G}_sanitizer_synccheck.txt
timeout 300 python tools/bayesod_envelope.py > ${G}_bayesod_envelope.txt 2>/dev/null
timeout 120 ./tools/_mma_shape_bench.bin > ${G}_mma_shape.txt 2>&1
timeout 300 python tools/q1_experiments.py 16 > ${G}_q1_experiments.txt 2>&1
timeout 300 python tools/tower_clock_probe.py 40 > ${G}_tower_clock.txt 2>&1
# summarise the ncu reports HERE (only gpurun_out/ travels back and it is capped at 64 MiB), keep the tower report only
POD_PROFILE_OUT=gpurun_out/${TAG}_profiles python tools/summarize_profiles.py ${TAG} > ${G}_summarize.log 2>&1
rm -f ${G}_out.ncu-rep ${G}_hbm.ncu-rep ${G}_backbone.ncu-rep ${G}_post.ncu-rep
tail -c 400 ${G}_bench_n1.json
for f in ${G}_sanitizer_memcheck.txt ${G}_sanitizer_racecheck.txt ${G}_sanitizer_synccheck.txt ${G}_summarize.log; do tail -n 3 $f; done
du -sh gpurun_out; ls -la gpurun_out | tail -30
