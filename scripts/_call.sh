set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_tests.log 2>&1; tail -25 gpurun_out/r2b_tests.log
{
for v in 0 2 3 4 5; do WN_VARIANT=$v timeout 300 python scripts/quick_bench.py --chains 65536 --reps 2 --integrator R2P; done
for v in 0 1 2 3 5; do WN_VARIANT=$v timeout 300 python scripts/quick_bench.py --chains 65536 --reps 2 --integrator fixed --H0 0.008; done
for v in 0 2 3 4; do WN_VARIANT=$v timeout 300 python scripts/config_sweep.py --only c3; done
timeout 300 python scripts/config_sweep.py --only c5
timeout 300 python scripts/config_sweep.py --only c1
} > gpurun_out/r2b_variants.log 2>&1
cat gpurun_out/r2b_variants.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 2000 gpurun_out/r2b_bench.err
