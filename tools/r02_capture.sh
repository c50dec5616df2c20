#!/bin/bash
# r02_capture.sh -- the profiler evidence of round 2, one GPU (run through gpurun).  Writes into gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
# every launch of one encode (x2) and one decode of a C4 block, with its device time (cold-cache, serialised)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_profile_one.csv python tools/profile_one.py > $O/r02_launches_profile_one.log 2>&1
# the top kernels with the full metric set and source correlation (second encode = warm)
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_pass1d|k_pass2d|k_row_census|k_p2d_summary|k_dec_write_delta|k_dec_tile_maps|k_dec_rows|k_dec_strip_summary" -f -o $O/r02_full python tools/profile_one.py > $O/r02_full.log 2>&1
# the same launch list for the bench command (short run)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench.csv python bench.py --blocks 4 --lanes 1 --e2e-lanes 1 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r02_launches_bench.log 2>&1
tail -2 $O/r02_launches_profile_one.log; tail -2 $O/r02_full.log; tail -c 300 $O/r02_launches_bench.log
