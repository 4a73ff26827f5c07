#!/bin/bash
# quick GPU check used during development: digit / int8 / SVGP tests, CTA-0 counters, kernel timings, a short bench line
python -m pytest tests/test_digits_gpu.py tests/test_ozaki_gpu.py tests/test_svgp_gpu.py -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
python tools/rq_counters.py > gpurun_out/q_counters.json 2>&1
python tools/bench_digits.py > gpurun_out/q_digits.jsonl 2>&1
python bench.py --legs none --steps 20 --warmup 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo bench rc=$?
