# round 1, fifth kernel generation (row-phase stacking, z-mask BN backward, side-stream wgrad, graphs): bench line, per-layer
# table, ncu launch list of a train step, DRAM traffic of the conv launches, full capture of the dominant conv kernel
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err
tail -c 2200 gpurun_out/bench_r01e.json
python scripts/layer_table.py r01e > gpurun_out/layer_table_r01e.log 2>&1; grep -E "^(fprop|dgrad|wgrad|all)" gpurun_out/layer_table_r01e.log
MMR_NO_GRAPH=1 python scripts/profile_step.py 16 4 > gpurun_out/plain_r01e.log 2>&1 && \
MMR_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 760 --csv --log-file gpurun_out/launches_r01e.csv python scripts/profile_step.py 16 4 > gpurun_out/ncu_r01e.log 2>&1
tail -n 2 gpurun_out/ncu_r01e.log
MMR_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_halo_kernel|conv_gemm_tc' -s 170 -c 85 --csv --log-file gpurun_out/conv_traffic.csv python scripts/profile_step.py 16 4 > gpurun_out/ncu3_r01e.log 2>&1
tail -n 2 gpurun_out/ncu3_r01e.log
STATS=1 ONLY=x_1_3.conv1 KINDS=fprop python scripts/bench_halo.py p1 > gpurun_out/plain2_r01e.log 2>&1 && \
STATS=1 ONLY=x_1_3.conv1 KINDS=fprop ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 1 -o gpurun_out/halo_x13_r01e python scripts/bench_halo.py p1 > gpurun_out/ncu2_r01e.log 2>&1
tail -n 2 gpurun_out/ncu2_r01e.log
python scripts/infer_bench.py 2>&1 | tail -1
