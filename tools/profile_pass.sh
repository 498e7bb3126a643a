#!/bin/bash
# One GPU: ncu --set full capture of each kernel the design names (summarised on the box by tools/ncu_summary.py: the
# reports themselves are 10-30 MB each and gpurun brings back at most 64 MiB; the dominant kernel's report is kept) + the
# launch list of one posterior draw of the bench command.
N="ncu --set full --clock-control none"
cap() {  # name regex skip -- command...
  name=$1; rx=$2; skip=$3; shift 3
  $N -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/r2_prof_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/r2_$name.json "$*" 2>> gpurun_out/r2_prof_$name.log | tail -1
}
cap gram_fast gram_fast 2 python tools/gram_once.py 16384 0
cap oz_mma6 oz_mma 1 python tools/gemm_once.py 8192 1 6
cp /tmp/oz_mma6.ncu-rep gpurun_out/r2_oz_mma6.ncu-rep
cap oz_mma7 oz_mma 1 python tools/gemm_once.py 8192 1 7
cap oz_slice oz_slice 1 python tools/gemm_once.py 8192 1 6
cap gemm_tma gemm_tma 1 python tools/gemm_once.py 8192 1 0
cap potrf_diag potrf_diag 8 python tools/potrf_once.py 4096
cap trsm_strip trsm_strip 4 python tools/potrf_once.py 4096
ncu --metrics gpu__time_duration.sum --clock-control none -s 14630 -c 640 --csv --log-file gpurun_out/r2_launches_one_draw.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dist --opt enqueue_threads=0 > gpurun_out/r2_prof_launches.log 2>&1
ls -la gpurun_out | tail -15
