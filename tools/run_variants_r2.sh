set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_posterior.py -x -q -m gpu -k "tall" 2>&1 | tail -5
bash tools/bench_variants.sh r2w "panel=1024" "" "panel=1024 tall_min=4096" "panel=2048"
