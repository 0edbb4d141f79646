ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --cpu-frames 0 > gpurun_out/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"bilateral_kernel|clean_evaluate_kernel|gn_persistent_kernel" --launch-skip 12 --launch-count 3 -o gpurun_out/top3_full -f python bench.py --steps 3 --warmup 3 --cpu-frames 0 > gpurun_out/b_ncu2.log 2>&1
tail -2 gpurun_out/b_ncu2.log
