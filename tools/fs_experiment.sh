python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/fs_tests.log
for m in 1 0 2; do python bench.py --no-cpu-baseline --face-swap $m > gpurun_out/fs_bench$m.log 2>&1; done
cat gpurun_out/fs_tests.log
