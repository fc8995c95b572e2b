# bench line + ncu launch list + full capture of the halo conv kernels (round 1, second kernel generation)
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err
tail -c 3000 gpurun_out/bench_r01b.json
python scripts/profile_step.py > gpurun_out/plain_r01b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 420 --csv --log-file gpurun_out/launches_r01b.csv python scripts/profile_step.py > gpurun_out/ncu_r01b.log 2>&1
tail -3 gpurun_out/ncu_r01b.log
ONLY=x_1_3.conv1 python scripts/bench_halo.py p1 > gpurun_out/plain2_r01b.log 2>&1 && \
ONLY=x_1_3.conv1 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 9 -c 3 -o gpurun_out/halo_x13_r01b python scripts/bench_halo.py p1 > gpurun_out/ncu2_r01b.log 2>&1
tail -3 gpurun_out/ncu2_r01b.log
