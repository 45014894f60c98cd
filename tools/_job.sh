python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; tail -5 gpurun_out/s3_pytest.log
python bench.py --sub '' --no-cpu-baseline --steps 100 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -3 gpurun_out/s3_bench.err
