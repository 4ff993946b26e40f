#!/bin/bash
# Off-box: turns gpurun_out/<name>.ncu-rep into profiles/<name>_ncu_summary.txt (the metrics DESIGN.md quotes).
#   scripts/ncu_summary.sh r02_tc2_construct r02_tc2_small r02_bptt_sweep
# OUTDIR (default profiles/) receives the summaries: on the GPU box the capture script sets OUTDIR=gpurun_out so that only the
# small text files travel back (gpurun returns at most 64 MiB).
OUTDIR=${OUTDIR:-profiles}
for n in "$@"; do
  rep=gpurun_out/$n.ncu-rep
  [ -f $rep ] || { echo "missing $rep"; continue; }
  out=$OUTDIR/${n}_ncu_summary.txt
  {
    echo "# ncu --set full --clock-control none --import-source on (one launch), extracted by scripts/ncu_summary.sh from $rep"
    ncu -i $rep --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__icc_requests_lookup_hit.sum", "sm__icc_requests_lookup_miss.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
for v in vals:
    d = dict(zip(hdr, v)); u = dict(zip(hdr, units))
    for k in want:
        if k in d: print(f"{k:90s} {d[k]:>22s} {u.get(k, str())}")
    stalls = sorted(((float(d[k].replace(",", "")), k) for k in d if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)
    print("-- warp stall reasons (average warps stalled per issue-active cycle), top 8")
    pre, suf = len("smsp__average_warps_issue_stalled_"), len("_per_issue_active.ratio")
    for val, k in stalls[:8]: print("   %-40s %8.2f" % (k[pre:-suf], val))
'
  } > $out
  echo wrote $out
done
