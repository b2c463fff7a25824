set -u
o=gpurun_out
python -m pytest tests/test_gpu_rollout.py tests/test_search_parity.py tests/test_gpu_mcts.py -m gpu -q -x > $o/pytest_small.log 2>&1; tail -3 $o/pytest_small.log
python scripts/profile_small_rollout.py > $o/sp_plain.json 2> $o/sp_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none --csv --log-file $o/sp_warp.csv python scripts/profile_small_rollout.py > $o/sp_warp.json 2> $o/sp_warp.err
python scripts/bench_small_rollout.py > $o/small_warp.json 2> $o/small_warp.err
python scripts/ab_root_parallel.py 1024 128 > $o/ab_warp.json 2> $o/ab_warp.err
tail -2 $o/ab_warp.err
