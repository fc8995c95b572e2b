# round 1, final round-1 build (row-phase stacking, z-mask and no-g BN backward, side-stream wgrad, graphs, PDL, head epilogue): bench line, per-layer
# table, ncu launch list of a train step, DRAM traffic of the conv launches, full capture of the dominant conv kernel
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err
tail -c 2200 gpurun_out/bench_r01i.json
python scripts/layer_table.py r01i > gpurun_out/layer_table_r01i.log 2>&1; grep -E "^(fprop|dgrad|wgrad|all)" gpurun_out/layer_table_r01i.log
MMR_NO_GRAPH=1 python scripts/profile_step.py 16 4 > gpurun_out/plain_r01i.log 2>&1 && \
MMR_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 760 --csv --log-file gpurun_out/launches_r01i.csv python scripts/profile_step.py 16 4 > gpurun_out/ncu_r01i.log 2>&1
tail -n 2 gpurun_out/ncu_r01i.log
MMR_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_halo_kernel|conv_gemm_tc' -s 170 -c 85 --csv --log-file gpurun_out/conv_traffic.csv python scripts/profile_step.py 16 4 > gpurun_out/ncu3_r01i.log 2>&1
tail -n 2 gpurun_out/ncu3_r01i.log
STATS=1 ONLY=x_1_3.conv1 KINDS=fprop python scripts/bench_halo.py p1 > gpurun_out/plain2_r01i.log 2>&1 && \
STATS=1 ONLY=x_1_3.conv1 KINDS=fprop ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 1 -o gpurun_out/halo_x13_r01i python scripts/bench_halo.py p1 > gpurun_out/ncu2_r01i.log 2>&1
tail -n 2 gpurun_out/ncu2_r01i.log
python scripts/infer_bench.py 2>&1 | tail -1
