#!/bin/bash
# usage: tools/bench_variants.sh tag "opt1" "opt2 opt3" ...   (each arg = space separated key=value list, "" for defaults)
tag=$1; shift
for o in "$@"; do
  args=""; for kv in $o; do if [[ $kv == streams=* ]]; then args="$args --streams ${kv#streams=}"; elif [[ $kv == draws=* ]]; then args="$args --draws ${kv#draws=}"; else args="$args --opt $kv"; fi; done
  name=$(echo "$o" | tr ' =' '__'); [ -z "$name" ] && name=default
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline $args > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/${tag}_$name.json").read().strip().splitlines()[-1])
    print("[$o]", "value %.2f e2e %.2f launches %d enqueue_ms %.1f clocks %s roof %.3f (%.2f ms)" % (b["value"], b["e2e"]["value"], b["gpu_launches"], b["host_enqueue_ms_per_step"], b["clocks"]["sm_mhz"], b["roofline"]["frac"], b["roofline"]["ms_per_launch"]))
except Exception as e:
    print("[$o] FAILED", e); print(open("gpurun_out/${tag}_$name.err").read()[-1500:])
PY
done
