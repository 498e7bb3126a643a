set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
bash tools/bench_variants.sh r2y "" "panel=2048" "panel=512"
