"""ncu_summary.py REPORT.ncu-rep OUT.json [note] -- the judged numbers of one `ncu --set full` capture (first profiled launch):
duration, DRAM bytes, achieved DRAM GB/s, pipe utilisation (tensor / fp64), registers, shared memory, occupancy.
Run where ncu is installed (no GPU needed): it only reads the report."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_subpipe_pct_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.sum": "dmma_instructions",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_active",
    "sm__inst_executed_pipe_fp64.sum": "fp64_pipe_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active_avg",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__cluster_dim_x": "cluster_dim_x",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
    "smsp__inst_executed.sum": "instructions_executed",
    "sm__inst_executed_pipe_uniform.sum": "uniform_pipe_instructions",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    first = rows[2]
    col = {h: i for i, h in enumerate(header)}
    res = {"report": rep, "kernel": first[col["Kernel Name"]], "note": note}
    for m, k in WANT.items():
        if m in col:
            v = first[col[m]].replace(",", "")
            try:
                res[k] = float(v)
            except ValueError:
                res[k] = v
            res.setdefault("units", {})[k] = units[col[m]]
    for m in header:   # anything naming the tcgen05 / UTC pipes
        if ("utc" in m.lower() or "tmem" in m.lower()) and m in col and first[col[m]]:
            try:
                res.setdefault("tcgen05_metrics", {})[m] = float(first[col[m]].replace(",", ""))
            except ValueError:
                pass
    if "duration_ns" in res:
        dur = res["duration_ns"] * (1e-3 if res["units"].get("duration_ns", "ns") in ("us", "usecond") else 1.0)
        u = res["units"].get("duration_ns", "")
        scale = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0}.get(u, 1e-9)
        sec = res["duration_ns"] * scale
        res["duration_ms"] = sec * 1e3
        def b(k):
            v, uu = res.get(k, 0.0), res["units"].get(k, "byte")
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(uu, 1)
        tot = b("dram_bytes_read") + b("dram_bytes_write")
        res["dram_bytes_per_launch"] = tot
        res["dram_gb_per_s"] = tot / sec / 1e9 if sec > 0 else None
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps({k: res.get(k) for k in ("kernel", "duration_ms", "dram_bytes_per_launch", "dram_gb_per_s", "tensor_pipe_pct_active",
                                              "fp64_pipe_pct_active", "registers_per_thread")}))


if __name__ == "__main__":
    main()
