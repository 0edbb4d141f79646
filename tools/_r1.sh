set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 1000 --warmup 30 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 2500 gpurun_out/bench_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --cpu-frames 0 > gpurun_out/b_ncu.log 2>&1
python tools/bench_objects.py > gpurun_out/bench_objects4.json 2>&1; tail -c 900 gpurun_out/bench_objects4.json
