#!/bin/bash
# Quick check of the RoI-pool kernels: their parity tests and the stage timings.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -x -k "roi or golden or smoke or graph" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_roi.log 2>&1; tail -n 3 gpurun_out/pytest_roi.log
timeout 200 python scripts/bench_stages.py > gpurun_out/stages.jsonl 2> gpurun_out/stages.err
grep -i "roi_pool" gpurun_out/stages.jsonl | cut -c1-130
