#!/bin/bash
# Evidence pass: full parity suite, smoke, both bench arms, ncu launch list of the sampling step, one `ncu --set full` capture per
# kernel class of the sampling step and of the training step.   usage: bash scripts/gpu_r2_evidence.sh <tag> [notrain]
cd "$(dirname "$0")/.."
tag=${1:-r02_ev}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err; echo "ref exit $?"
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_$tag.json
# launch list of three step replays (plain run first)
timeout 300 python scripts/step_n.py 64 3 > gpurun_out/plain_$tag.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python scripts/step_n.py 64 3 > gpurun_out/ncu_list_$tag.log 2>&1
python scripts/ncu_list.py gpurun_out/launches_$tag.csv x > gpurun_out/launches_${tag}_summary.txt 2>&1; head -24 gpurun_out/launches_${tag}_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/conv_dram_$tag.csv python scripts/step_n.py 64 2 > gpurun_out/ncu_dram_$tag.log 2>&1; echo "dram exit $?"
# one full capture per kernel class of the sampling step: "<regex>:<skip>:<name>"
# (-k matches the base name only: the skip counts pick the template instance / layer, see profiles/launches_*_summary.txt)
for spec in "conv_tc_halo_persist_kernel:1:persist32" "conv_tc_halo_persist_kernel:4:persist16_128to256" "conv_tc_halo_persist_kernel:7:persist16_256to256" \
            "conv_tc_kernel:3:gen8x8" "conv_tc_kernel:11:splitk4x4" "conv_tc_kernel:0:gen_1x1" "conv_tc_gemm_persist_kernel:1:gemm_qkv32" \
            "linattn_ctxmix_kernel:0:linattn32" "gn_mish_sum_kernel:0:gnsum4x4"; do
  IFS=: read -r rx skip name <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -o gpurun_out/prof_${name}_$tag -f python scripts/step_n.py 64 2 > gpurun_out/ncu_${name}_$tag.log 2>&1
  echo "ncu $name exit $?"
done
if [ "$2" != "notrain" ]; then
  timeout 900 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train_$tag.json 2> gpurun_out/bench_train_$tag.err; echo "train exit $?"; cut -c1-200 gpurun_out/bench_train_$tag.json
  for spec in "conv_tc32_persist_kernel:40:train_tc32" "conv_tc32_halo_kernel:4:train_tc32_halo" "wgrad_tc32_kernel:20:train_wgrad"; do
    IFS=: read -r rx skip name <<< "$spec"
    DD_NO_RELAYOUT_PLAN=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -o gpurun_out/prof_${name}_$tag -f python scripts/train_n.py 32 2 > gpurun_out/ncu_${name}_$tag.log 2>&1
    echo "ncu $name exit $?"
  done
fi
