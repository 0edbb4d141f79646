ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 700 -c 420 --csv --log-file gpurun_out/launches_objects4.csv python tools/bench_objects.py --steps 3 --warmup 3 > gpurun_out/o_ncu.log 2>&1
tail -1 gpurun_out/o_ncu.log | cut -c1-200
