#!/bin/bash
# Round-end evidence pass: parity suite, both bench arms, launch list + DRAM traffic + one full capture, training bench.
cd "$(dirname "$0")/.."
tag=${1:-r01_f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err; echo "ref exit $?"
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_$tag.json
timeout 900 python bench.py --workload train > gpurun_out/bench_train_$tag.json 2> gpurun_out/bench_train_$tag.err; echo "train exit $?"; cut -c1-200 gpurun_out/bench_train_$tag.json
# ncu passes (numbers printed under ncu are never bench values)
timeout 300 python scripts/step_n.py 64 3 > gpurun_out/plain_step_$tag.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python scripts/step_n.py 64 3 > gpurun_out/ncu_step_$tag.log 2>&1
python scripts/ncu_list.py gpurun_out/launches_$tag.csv x > gpurun_out/list_$tag.txt 2>&1; head -12 gpurun_out/list_$tag.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/conv_dram_$tag.csv python scripts/step_n.py 64 2 > gpurun_out/ncu_dram_$tag.log 2>&1; echo "dram exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo -s 3 -c 1 -o gpurun_out/prof_conv_halo_$tag -f python scripts/step_n.py 64 2 > gpurun_out/ncu_full_$tag.log 2>&1; echo "full exit $?"
