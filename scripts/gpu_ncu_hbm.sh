#!/bin/bash
# DRAM / L2 byte counters of the memory-bound kernels (north_star: "achieved HBM GB/s against peak for the norm and elementwise
# kernels"): one metrics pass over two replays of the sampling step, one over three fused optimizer steps.  Usage: gpu_ncu_hbm.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-hbm}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 300 python scripts/step_n.py 64 2 > /dev/null 2>&1 && timeout 600 ncu --metrics $M --clock-control none -k regex:'gn_mish|posterior|layernorm|linattn|s2d|im2col|tick' --csv --log-file gpurun_out/hbm_step_$tag.csv python scripts/step_n.py 64 2 > gpurun_out/ncu_hbm_step_$tag.log 2>&1; echo "step exit $?"
timeout 300 python scripts/opt_n.py 3 > /dev/null 2>&1 && timeout 600 ncu --metrics $M --clock-control none -k regex:'adam_ema|grad_sqnorm|grad_norm_finish|ema_update' --csv --log-file gpurun_out/hbm_opt_$tag.csv python scripts/opt_n.py 3 > gpurun_out/ncu_hbm_opt_$tag.log 2>&1; echo "opt exit $?"
python scripts/ncu_hbm_summary.py gpurun_out/hbm_step_$tag.csv gpurun_out/hbm_opt_$tag.csv > gpurun_out/hbm_summary_$tag.txt 2>&1; head -40 gpurun_out/hbm_summary_$tag.txt
