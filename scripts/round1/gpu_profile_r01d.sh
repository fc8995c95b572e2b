# round 1, fourth kernel generation (fast epilogue, CUDA graphs): bench line, per-layer table, ncu launch list of a train step,
# full captures of the dominant conv kernel (x_1_3.conv1 fprop with statistics) and of the BN backward reduce
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err
tail -c 1800 gpurun_out/bench_r01d.json
python scripts/layer_table.py r01d > gpurun_out/layer_table_r01d.log 2>&1; tail -30 gpurun_out/layer_table_r01d.log | head -5
MMR_NO_GRAPH=1 python scripts/profile_step.py 16 4 > gpurun_out/plain_r01d.log 2>&1 && \
MMR_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 760 --csv --log-file gpurun_out/launches_r01d.csv python scripts/profile_step.py 16 4 > gpurun_out/ncu_r01d.log 2>&1
tail -n 2 gpurun_out/ncu_r01d.log
STATS=1 ONLY=x_1_3.conv1 KINDS=fprop python scripts/bench_halo.py p1 > gpurun_out/plain2_r01d.log 2>&1 && \
STATS=1 ONLY=x_1_3.conv1 KINDS=fprop ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 1 -o gpurun_out/halo_x13_r01d python scripts/bench_halo.py p1 > gpurun_out/ncu2_r01d.log 2>&1
tail -n 2 gpurun_out/ncu2_r01d.log
python scripts/bench_elementwise.py > gpurun_out/elementwise_r01d.log 2>&1; cat gpurun_out/elementwise_r01d.log
