# GPU tests + bench line + step breakdown + ncu launch list + full capture of the halo conv kernel (round 1, third kernel generation)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01c.log 2>&1; tail -3 gpurun_out/pytest_r01c.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
tail -c 3000 gpurun_out/bench_r01c.json
python scripts/step_breakdown.py > gpurun_out/breakdown_r01c.log 2>&1; tail -45 gpurun_out/breakdown_r01c.log
python scripts/profile_step.py > gpurun_out/plain_r01c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 420 --csv --log-file gpurun_out/launches_r01c.csv python scripts/profile_step.py > gpurun_out/ncu_r01c.log 2>&1
tail -3 gpurun_out/ncu_r01c.log
ONLY=x_1_3.conv1 python scripts/bench_halo.py p1 > gpurun_out/plain2_r01c.log 2>&1 && \
ONLY=x_1_3.conv1 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 9 -c 3 -o gpurun_out/halo_x13_r01c python scripts/bench_halo.py p1 > gpurun_out/ncu2_r01c.log 2>&1
tail -3 gpurun_out/ncu2_r01c.log
