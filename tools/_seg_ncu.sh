ncu --set full --import-source on --clock-control none --launch-skip 44 --launch-count 22 -o gpurun_out/segment_full -f python tools/time_segment.py > gpurun_out/seg_ncu.log 2>&1
tail -2 gpurun_out/seg_ncu.log
