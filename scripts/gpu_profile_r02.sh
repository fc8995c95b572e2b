# round 2 profile pass: ncu launch list of two train steps (after the same command exited 0 without ncu),
# step breakdown by C-ABI entry point, per-layer table.  Usage: bash scripts/gpu_profile_r02.sh <tag>
tag=${1:-r02}
set -x
MMR_NO_GRAPH=1 python scripts/profile_step.py 16 4 > gpurun_out/plain_$tag.log 2>&1 && \
MMR_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 760 --csv --log-file gpurun_out/launches_$tag.csv python scripts/profile_step.py 16 4 > gpurun_out/ncu_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_$tag.log
python scripts/launch_summary.py gpurun_out/launches_$tag.csv > gpurun_out/launch_summary_$tag.txt 2>&1; head -45 gpurun_out/launch_summary_$tag.txt
python scripts/step_breakdown.py > gpurun_out/breakdown_$tag.txt 2>&1; head -30 gpurun_out/breakdown_$tag.txt
