mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 200 python scripts/bench_stages.py > gpurun_out/stages.jsonl 2> gpurun_out/stages.err
grep -i "postprocess" gpurun_out/stages.jsonl | cut -c1-260
tail -n 3 gpurun_out/stages.err
