# final validation of a round: smoke, the GPU suite, the bench line, the other single-GPU configurations
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err || tail -20 gpurun_out/r2_bench_final.err
timeout 600 python tools/bench_configs.py > gpurun_out/r2_configs_c1_c2_c3.json 2> gpurun_out/r2_configs.err || tail -20 gpurun_out/r2_configs.err
cat gpurun_out/r2_configs_c1_c2_c3.json
