python tools/bench_conv.py 1 128x32768 > gpurun_out/r2_conv_plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:lconv_kernel -s 4 -c 1 -o gpurun_out/r2_lconv_d python tools/bench_conv.py 1 128x32768 > gpurun_out/r2_ncu_d.log 2>&1
python tools/bench_smooth.py --reads 1000000 --iters 3 > gpurun_out/r2_smooth_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:smooth_chop_kernel -s 1 -c 1 -o gpurun_out/r2_smooth_a python tools/bench_smooth.py --reads 1000000 --iters 3 > gpurun_out/r2_ncu_s.log 2>&1
cat gpurun_out/r2_smooth_plain.log | tail -2
python -m pytest tests/test_gpu_encode.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -3
python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_bench2.json 2>gpurun_out/r2_bench2.err; python tools/kernel_table.py gpurun_out/r2_bench2.json | head -4
