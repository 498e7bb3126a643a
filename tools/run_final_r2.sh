# final validation of a round: smoke, the GPU suite, the bench line, the launch list of one draw
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err || tail -20 gpurun_out/r2_bench_final.err
CMD="ncu --metrics gpu__time_duration.sum --clock-control none -s 14246 -c 620 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dist --opt enqueue_threads=0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 14246 -c 620 --csv --log-file gpurun_out/r2_launches_one_draw_final.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dist --opt enqueue_threads=0 > gpurun_out/r2_prof_launches.log 2>&1
python tools/launch_list_summary.py gpurun_out/r2_launches_one_draw_final.csv gpurun_out/r2_launches_one_draw_final_summary.json "$CMD" \
    "620 consecutive launches of a timed step of the final build (1024-wide diagonal blocks: one posterior draw is 593 launches; the window holds one draw plus the head of the next). Serialised per-launch times, caches flushed between launches: compare SHARES with the step, not absolutes."
