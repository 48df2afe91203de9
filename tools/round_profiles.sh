#!/bin/bash
# Everything the round's profiles/ directory is refreshed from, in one single-GPU call:
#   gpurun --timeout 1500 -- 'bash tools/round_profiles.sh'
# Bench numbers come from plain runs; the ncu passes are separate and only feed launch lists / metrics.
set -u
O=gpurun_out/final
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "PARITY|G8|G9|row-form|pre-norm tap|passed|failed|skipped" ) > $O/tests.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/bench_rows.py > $O/rows.json 2> $O/rows.err
python tools/bench_torch_gpu.py > $O/torch_gpu.json 2> $O/torch_gpu.err
# every launch of one timed step with its device time (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launches.log 2>&1
# full metric set of every tcgen05 launch of one forward (second forward of tools/ncu_forward.py)
ncu --set full --clock-control none --import-source on -k regex:"conv3_umma|conv3_rows|stem_umma" --launch-skip 21 --launch-count 21 -f -o $O/fwd python tools/ncu_forward.py > $O/ncu_full.log 2>&1
ncu -i $O/fwd.ncu-rep --page raw --csv > $O/fwd_raw.csv 2>/dev/null
rm -f $O/fwd.ncu-rep     # tens of MB; the raw page holds every metric the summaries use
# the 94M model's launches
ncu --set full --clock-control none -k regex:"conv3_umma|stem_umma|upsample2|inorm|pool2" --launch-skip 57 --launch-count 57 -f -o $O/fwd94 python tools/ncu_forward_94m.py > $O/ncu_full94.log 2>&1
ncu -i $O/fwd94.ncu-rep --page raw --csv > $O/fwd94_raw.csv 2>/dev/null
rm -f $O/fwd94.ncu-rep
tail -3 $O/tests.log; head -c 400 $O/bench_n1.json; echo; tail -2 $O/ncu_full.log
