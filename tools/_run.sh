python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity_configs0.py tests/test_gpu_encode.py -q -m gpu -x -s 2>&1 | grep -vE "^$|Warning|warn|torch.frombuffer|self.blob" | tail -14
python tools/parity_configs0.py 1000 2>&1 | tail -1 | tee gpurun_out/r2_parity_configs0.json
python tools/bench_cli.py 100000 AB 2>&1 | tail -12 | tee gpurun_out/r2_bench_cli.log
