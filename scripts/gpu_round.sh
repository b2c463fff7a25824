#!/bin/bash
# One GPU-box pass of a round (run through gpurun):  scripts/gpu_round.sh r02a [tests|bench|profile ...]
#   tests    python -m pytest tests -m gpu  (rebuilds the library on the box first, see tests/conftest.py)
#   bench    the default bench command, 1 GPU
#   profile  scripts/profile_round.sh <tag>: ncu launch list + one `--set full` capture with source pages
#   small    one-leaf rollout launches under ncu (cycles per move of rollout_warp_kernel / rollout_small_kernel) + call latency
#   groups   root-parallel search against the number of leaf batches in flight
set -u
tag=${1:-rXX}; shift
what=${*:-tests bench profile}
out=gpurun_out
mkdir -p $out
rc=0
for w in $what; do
  case $w in
    tests)   python -m pytest tests -m gpu -q -s > $out/pytest_$tag.log 2>&1; r=$?; tail -5 $out/pytest_$tag.log; [ $r -ne 0 ] && rc=$r ;;
    gtests)  python -m pytest tests/test_guided.py tests/test_gpu_eval.py -m gpu -q -s > $out/pytest_$tag.log 2>&1; r=$?; tail -4 $out/pytest_$tag.log; [ $r -ne 0 ] && rc=$r ;;
    bench)   python bench.py --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; r=$?; head -c 600 $out/bench_$tag.json; echo; [ $r -ne 0 ] && { tail -5 $out/bench_$tag.err; rc=$r; } ;;
    refarm)  python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err || rc=$? ;;
    profile) bash scripts/profile_round.sh $tag || rc=$? ;;
    guided)  python scripts/bench_guided.py 256 1024 8192 65536 > $out/guided_$tag.json 2> $out/guided_$tag.err; r=$?; cat $out/guided_$tag.err | cut -c1-400; [ $r -ne 0 ] && rc=$r ;;
    gprofile) python scripts/profile_guided.py > $out/gplain_$tag.log 2>&1 && \
             ncu --set full --clock-control none --import-source on -f -o /tmp/gprof_$tag python scripts/profile_guided.py > $out/gncu_$tag.log 2>&1; \
             ncu -i /tmp/gprof_$tag.ncu-rep --page source --csv --kernel-name regex:guided_kernel --launch-count 1 > $out/src_guided1k_$tag.csv 2>/dev/null; \
             ncu -i /tmp/gprof_$tag.ncu-rep --page raw --csv > $out/gprof_${tag}_raw.csv 2>/dev/null; tail -2 $out/gncu_$tag.log ;;
    parity)  python tests/tools/full_parity.py > $out/full_parity_$tag.json 2> $out/full_parity_$tag.err; r=$?; head -c 700 $out/full_parity_$tag.json; echo; [ $r -ne 0 ] && { tail -3 $out/full_parity_$tag.err; rc=$r; } ;;
    benchN)  n=$(nvidia-smi -L | wc -l); python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > $out/bench_n${n}_$tag.json 2> $out/bench_n${n}_$tag.err; r=$?; head -c 400 $out/bench_n${n}_$tag.json; echo; [ $r -ne 0 ] && { tail -8 $out/bench_n${n}_$tag.err; rc=$r; } ;;
    fuzz)    python tests/tools/fuzz_guided.py > $out/fuzz_guided_$tag.json 2> $out/fuzz_guided_$tag.err; r=$?; cat $out/fuzz_guided_$tag.json; tail -3 $out/fuzz_guided_$tag.err; [ $r -ne 0 ] && rc=$r
             python tests/tools/fuzz_eval.py 100000 > $out/fuzz_eval_$tag.txt 2>&1; r=$?; tail -1 $out/fuzz_eval_$tag.txt; [ $r -ne 0 ] && rc=$r
             python tests/tools/fuzz_rollout.py 20000 > $out/fuzz_rollout_$tag.txt 2>&1; r=$?; tail -1 $out/fuzz_rollout_$tag.txt; [ $r -ne 0 ] && rc=$r
             python tests/tools/fuzz_search.py 600 > $out/fuzz_search_$tag.json 2> $out/fuzz_search_$tag.err; r=$?; tail -1 $out/fuzz_search_$tag.json; tail -3 $out/fuzz_search_$tag.err; [ $r -ne 0 ] && rc=$r ;;
    sweep)   python scripts/sweep_root_parallel.py > $out/rp_sweep_$tag.json 2> $out/rp_sweep_$tag.err; r=$?; tail -6 $out/rp_sweep_$tag.err | cut -c1-300; [ $r -ne 0 ] && rc=$r ;;
    small)   python scripts/profile_small_rollout.py > $out/small_plain_$tag.json 2> $out/small_$tag.err && \
             ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none --csv --log-file $out/small_launches_$tag.csv python scripts/profile_small_rollout.py > $out/small_lengths_$tag.json 2>> $out/small_$tag.err; \
             python scripts/small_rollout_summary.py $out/small_launches_$tag.csv $out/small_lengths_$tag.json > $out/small_rollout_$tag.json 2>> $out/small_$tag.err; head -c 900 $out/small_rollout_$tag.json; echo; \
             ncu --set full --import-source on --clock-control none -k regex:rollout_warp -c 1 -f -o /tmp/small_$tag python scripts/profile_small_rollout.py > /dev/null 2>> $out/small_$tag.err; \
             ncu -i /tmp/small_$tag.ncu-rep --page source --csv > $out/src_rollout_warp_$tag.csv 2>/dev/null; \
             ncu -i /tmp/small_$tag.ncu-rep --page raw --csv > $out/small_raw_$tag.csv 2>/dev/null; \
             python scripts/bench_small_rollout.py > $out/small_latency_$tag.json 2>> $out/small_$tag.err ;;
    groups)  python scripts/ab_root_parallel.py 8192 1024 512 128 > $out/rp_groups_$tag.json 2> $out/rp_groups_$tag.err; r=$?; tail -2 $out/rp_groups_$tag.err; [ $r -ne 0 ] && rc=$r ;;
    *)       echo "unknown step $w" ;;
  esac
done
exit $rc
