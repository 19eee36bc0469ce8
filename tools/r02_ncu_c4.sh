#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; tag=${1:-r02c}
ncu --set full --clock-control none --import-source on -k regex:raymarchBatchRefill -s 4 -c 1 -f -o gpurun_out/${tag}_k1_refill \
    python bench.py --workload c4_ao_sdf2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_k1_refill.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_k1_refill.ncu-rep --json gpurun_out/${tag}_k1_refill.json > gpurun_out/${tag}_k1_refill.txt 2>&1
SVO_BENCH_NO_LANE_REFILL=1 ncu --set full --clock-control none --import-source on -k regex:raymarchBatchKernel -s 4 -c 1 -f -o gpurun_out/${tag}_k1_plain \
    python bench.py --workload c4_ao_sdf2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_k1_plain.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_k1_plain.ncu-rep --json gpurun_out/${tag}_k1_plain.json > gpurun_out/${tag}_k1_plain.txt 2>&1
paste <(head -48 gpurun_out/${tag}_k1_refill.txt | cut -c1-95) <(head -48 gpurun_out/${tag}_k1_plain.txt | cut -c74-95)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_c4.csv \
    python bench.py --workload c4_ao_sdf2048 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/${tag}_launches_c4.csv | awk -F'","' '{print $5, $(NF)}' | sort | uniq -c | sort -rn | head -12
