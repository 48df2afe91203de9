#!/bin/bash
# Everything the round's profiles/ directory is refreshed from, in one GPU call:
#   gpurun --timeout 1500 -- 'bash tools/round_profiles.sh'
# Bench numbers come from plain runs; the ncu passes are separate and only feed launch lists / metrics.
set -u
O=gpurun_out/final
mkdir -p $O
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $O/tests.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/bench_rows.py > $O/rows.json 2> $O/rows.err
python tools/bench_configs.py 94m 512 > $O/other_configs.json 2> $O/other_configs.err
python tools/bench_torch_gpu.py > $O/torch_gpu.json 2> $O/torch_gpu.err
./tools/micro/store_bw > $O/store_bw.txt 2>&1
python tools/ablate.py 0 1 2 3 4 15 > $O/ablate.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv3_umma|conv3_rows|stem_umma" --launch-skip 21 --launch-count 21 -f -o $O/fwd python tools/ncu_forward.py > $O/ncu_full.log 2>&1
ncu -i $O/fwd.ncu-rep --page raw --csv > $O/fwd_raw.csv 2>/dev/null
rm -f $O/fwd.ncu-rep     # 45 MB; the raw page holds every metric the summaries use
cat $O/tests.log; head -c 600 $O/bench_n1.json; echo; tail -2 $O/ncu_full.log
