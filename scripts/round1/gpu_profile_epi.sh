# source-level captures of the epilogue-bound layers: 64->64 @256^2 (fprop with statistics, dgrad) and the 16-channel 512^2 layer
set -x
export STATS=1
ONLY=x_3_3.conv2 KINDS=fprop,dgrad python scripts/bench_halo.py epi
ONLY=x_0_4.conv2 KINDS=fprop,dgrad python scripts/bench_halo.py epi
ONLY=x_3_3.conv2 KINDS=fprop ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 1 -o gpurun_out/epi_c64_fprop python scripts/bench_halo.py epi > gpurun_out/ncu_epi1.log 2>&1
ONLY=x_3_3.conv2 KINDS=dgrad ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 1 -o gpurun_out/epi_c64_dgrad python scripts/bench_halo.py epi > gpurun_out/ncu_epi2.log 2>&1
ONLY=x_0_4.conv2 KINDS=fprop ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 1 -o gpurun_out/epi_c16_fprop python scripts/bench_halo.py epi > gpurun_out/ncu_epi3.log 2>&1
tail -2 gpurun_out/ncu_epi*.log
