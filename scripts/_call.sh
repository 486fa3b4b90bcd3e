mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpu.txt
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/c1_pytest.txt
./scripts/ubench/atomics_bench > gpurun_out/c1_atomics.txt 2>&1
python bench.py --steps 128 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
TNF_OVERLAP_SCATTER=1 python bench.py --steps 128 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/c1_bench_overlap.json 2> gpurun_out/c1_bench_overlap.err
tail -3 gpurun_out/c1_pytest.txt; cat gpurun_out/c1_atomics.txt
