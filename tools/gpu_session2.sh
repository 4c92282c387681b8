#!/bin/bash
# round-2 GPU session 2: full GPU parity suite (incl. long-horizon goldens), full default bench, racecheck
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s2_pytest.log
tail -15 gpurun_out/s2_pytest.log
timeout 900 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s2_bench_ref.json 2> gpurun_out/s2_bench_ref.err; echo "ref rc $?"
timeout 120 python tools/diag_sanitize.py > gpurun_out/s2_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tools/diag_sanitize.py > gpurun_out/s2_racecheck.log 2>&1; echo "racecheck rc $?"
tail -5 gpurun_out/s2_racecheck.log
