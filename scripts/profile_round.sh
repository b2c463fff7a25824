#!/bin/bash
# Profiling pass of one round (run on the GPU box through gpurun):  scripts/profile_round.sh r01i
#   1. the workloads WITHOUT ncu (must exit 0 first)
#   2. launch list of the default bench command (ncu --metrics gpu__time_duration.sum)
#   3. one `ncu --set full` capture of every kernel at the bench sizes (scripts/profile_target.py)
# Outputs land in gpurun_out/; scripts/ncu_summary.py and scripts/sass_by_line.py turn them into profiles/.
set -u
tag=${1:-rXX}
out=gpurun_out
export GK_PROFILE_ALL=1 GK_PROFILE_BOARDS=1048576 GK_PROFILE_ROLLOUTS=4096
python scripts/profile_target.py > $out/plain_$tag.log 2>&1 || { echo "profile_target failed"; tail -5 $out/plain_$tag.log; exit 1; }
python bench.py --steps 2 --warmup 3 --no-cpu > $out/bench_plain_$tag.json 2> $out/bench_plain_$tag.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_launch_$tag.log 2>&1
# the report itself stays on the box (gpurun_out/ is capped at 64 MiB): bring back its raw and per-kernel source pages
rep=/tmp/prof_$tag
GK_PROFILE_ONCE=1 ncu --set full --clock-control none --import-source on -f -o $rep python scripts/profile_target.py > $out/ncu_$tag.log 2>&1
tail -2 $out/ncu_$tag.log
ncu -i $rep.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --csv --kernel-name regex:ac_eval_kernel --launch-count 1 > $out/src_eval_$tag.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --csv --kernel-name regex:rollout_kernel --launch-count 1 > $out/src_roll_$tag.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --csv --kernel-name regex:guided_kernel --launch-count 1 > $out/src_guided_$tag.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --csv --kernel-name regex:guided_pair_kernel --launch-count 1 > $out/src_guided_pair_$tag.csv 2>/dev/null
ls -la $rep.ncu-rep $out/*_$tag*
