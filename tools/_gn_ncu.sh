ncu --set full --import-source on --clock-control none -k regex:gn_persistent --launch-skip 2 --launch-count 1 -o gpurun_out/gn_full -f python tools/prof_track.py 640 4 > gpurun_out/gn_ncu.log 2>&1
tail -2 gpurun_out/gn_ncu.log
