# c3 (ResNetUNet-34, batch 32 @ 512x512): step breakdown per C-ABI entry point and ncu launch list of one step
tag=${1:-r02v}
set -x
python scripts/step_breakdown.py 0 c3 > gpurun_out/${tag}_c3_breakdown.txt 2>&1; head -40 gpurun_out/${tag}_c3_breakdown.txt
MMR_NO_GRAPH=1 python scripts/profile_step.py 0 3 c3 > gpurun_out/${tag}_c3_plain.log 2>&1 && \
MMR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_c3_launches.csv python scripts/profile_step.py 0 3 c3 > gpurun_out/${tag}_c3_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_c3_launches.csv > gpurun_out/${tag}_c3_launch_summary.txt 2>&1; head -60 gpurun_out/${tag}_c3_launch_summary.txt
