# usage (under gpurun): bash tools/profile.sh <tag> [bench args...]
# 1) plain run (must exit 0), 2) launch list with per-launch device time, 3) one --set full capture of the segment kernel
TAG=$1; shift
ARGS="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e $@"
python bench.py $ARGS > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
python bench.py $ARGS > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:ude_(seg|mma2?|tf32)_kernel" -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-160
