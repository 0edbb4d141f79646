ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_segment.csv python tools/time_segment.py > gpurun_out/seg_prof.log 2>&1
tail -3 gpurun_out/seg_prof.log
