# round 1, last build (r01i + 32-byte BatchNorm apply passes, four-pixel head gradient prep, block-wise branch-free
# max-pool, parallel block sums in the reductions): full GPU test suite, bench line, ncu launch list of a train
# step, step breakdown.  The conv kernels are those of r01i (profiles/r01i_* still describe them).
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01j.json 2> gpurun_out/bench_r01j.err
tail -c 2400 gpurun_out/bench_r01j.json
MMR_NO_GRAPH=1 python scripts/profile_step.py 16 4 > gpurun_out/plain_r01j.log 2>&1 && \
MMR_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 760 --csv --log-file gpurun_out/launches_r01j.csv python scripts/profile_step.py 16 4 > gpurun_out/ncu_r01j.log 2>&1
tail -n 2 gpurun_out/ncu_r01j.log
python scripts/step_breakdown.py > gpurun_out/breakdown_r01j_final.txt 2>&1; head -12 gpurun_out/breakdown_r01j_final.txt
