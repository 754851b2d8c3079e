#!/bin/bash
# round-2 evidence pass on one B200: sanitizer over the MSM / prover kernels, ncu launch lists, ncu --set full of the top kernels
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool python scripts/sanitize_workload.py msm prove > gpurun_out/sanitizer_r02b_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|msm ok|prove ok" gpurun_out/sanitizer_r02b_$tool.log | tail -4
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-prove --no-cpu > gpurun_out/b.log 2>&1; echo "bench launch list rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_prove_22.csv python scripts/tiny_prove.py 22 2 > gpurun_out/p.log 2>&1; echo "prove launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_msm_accumulate_24 -f python scripts/prof.py msm 24 3 > gpurun_out/n1.log 2>&1; echo "ncu accumulate rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_pass4_kernel --launch-skip 3 --launch-count 3 -o gpurun_out/r02_ntt_pass4_24 -f python scripts/prof.py ntt 24 2 > gpurun_out/n2.log 2>&1; echo "ncu ntt rc=$?"
ls -la gpurun_out | tail -12
