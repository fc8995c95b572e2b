set -x
export ONLY=x_0_4.conv2 KINDS=fprop,wgrad
python scripts/bench_halo.py thin > gpurun_out/plain_thin.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 6 -c 1 -o gpurun_out/thin_fprop python scripts/bench_halo.py thin > gpurun_out/ncu_thin.log 2>&1
tail -3 gpurun_out/ncu_thin.log
cat gpurun_out/plain_thin.log | tail -4
