python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 2600 gpurun_out/bench_r01_final.json
python tools/trace_persistent.py > gpurun_out/tracker_phase_trace.txt 2>&1; tail -3 gpurun_out/tracker_phase_trace.txt
python tools/time_tracker.py > gpurun_out/time_tracker.txt 2>&1; head -3 gpurun_out/time_tracker.txt
python tools/time_segment.py > gpurun_out/time_segment.txt 2>&1; cat gpurun_out/time_segment.txt
python bench.py --impl reference --steps 4 --warmup 1 | cut -c1-200
