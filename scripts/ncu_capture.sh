#!/bin/bash
# Run ON the GPU box (through gpurun): ncu captures of the three kernels DESIGN.md quotes, each after its plain command has
# exited 0.  Outputs under gpurun_out/ (.ncu-rep + launch lists); summaries are extracted off-box by scripts/ncu_summary.sh.
#   scripts/ncu_capture.sh <tag> [construct] [small] [bptt] [layered]      (default: all)
set -x
T=${1:-r02}
shift
WHAT="${*:-construct small bptt layered}"
has() { case " $WHAT " in *" $1 "*) return 0;; *) return 1;; esac; }
O=gpurun_out
mkdir -p $O
FWD="python bench.py --steps 3 --warmup 3 --no-bptt --no-fp32 --no-e2e --no-cpu-baseline --no-parity"
$FWD > $O/${T}_plain_fwd.json 2> $O/${T}_plain_fwd.err || exit 1
# launch list of the same command (kernel SHARES of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${T}_fwd_launches.csv $FWD > $O/${T}_ncu_fwd_list.log 2>&1
# launches of rollout_tc2_kernel in that command: 6 of construct_NN (3 warm-up + 3 timed), then 6 of the 50/20 mish net
has construct && ncu --set full --clock-control none --import-source on -k regex:rollout_tc2_kernel -s 4 -c 1 -o $O/${T}_tc2_construct $FWD > $O/${T}_ncu_tc2.log 2>&1
has small && ncu --set full --clock-control none --import-source on -k regex:rollout_tc2_kernel -s 10 -c 1 -o $O/${T}_tc2_small $FWD > $O/${T}_ncu_tc2s.log 2>&1
BPTT="python bench.py --steps 3 --warmup 3 --no-small --no-fp32 --no-e2e --no-cpu-baseline --no-parity --bptt-steps 1 --columns 65536"
if has bptt; then
  $BPTT > $O/${T}_plain_bptt.json 2> $O/${T}_plain_bptt.err || exit 1
  ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o $O/${T}_bptt_sweep $BPTT > $O/${T}_ncu_bptt.log 2>&1
fi
if has layered; then   # the GEMM kernel of the layered path on configs[4]'s 1024 x 1024 layer (third layer_gemm launch of the warm region)
  LAY="python scripts/layered_time.py"
  $LAY > $O/${T}_plain_layered.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:layer_gemm_kernel -s 16 -c 1 -o $O/${T}_layer_gemm $LAY > $O/${T}_ncu_layer.log 2>&1
fi
ls -la $O/${T}_*.ncu-rep
# summaries on the box; the reports themselves stay behind except the headline kernel's (gpurun returns at most 64 MiB)
OUTDIR=$O scripts/ncu_summary.sh ${T}_tc2_construct ${T}_tc2_small ${T}_bptt_sweep ${T}_layer_gemm
for n in ${T}_tc2_construct ${T}_tc2_small ${T}_bptt_sweep ${T}_layer_gemm; do
  [ -f $O/$n.ncu-rep ] && ncu -i $O/$n.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${n}_source.csv.gz
done
rm -f $O/${T}_tc2_small.ncu-rep $O/${T}_bptt_sweep.ncu-rep $O/${T}_layer_gemm.ncu-rep
du -sh $O
