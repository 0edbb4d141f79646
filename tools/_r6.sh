python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; cut -c1-300 gpurun_out/bench_r01_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_final.json')); print('e2e', d['e2e'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-160
