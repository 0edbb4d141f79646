python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; cut -c1-420 gpurun_out/bench_r01_final.json
python tools/bench_objects.py > gpurun_out/bench_objects4.json 2>&1; tail -1 gpurun_out/bench_objects4.json | cut -c1-400
