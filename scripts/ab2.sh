o=gpurun_out
python scripts/ab_root_parallel.py 1024 512 256 > $o/ab2_warp.json 2> $o/ab2_warp.err
GK_AB_NO_WARP=1 python scripts/ab_root_parallel.py 1024 512 256 > $o/ab2_thread.json 2> $o/ab2_thread.err
tail -2 $o/ab2_thread.err
