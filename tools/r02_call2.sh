mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_unet.py -m gpu -q -s 2>&1 | grep -v "^\[tap\|sampling loop" > gpurun_out/r02c2_tests.txt; tail -60 gpurun_out/r02c2_tests.txt | cut -c1-400
