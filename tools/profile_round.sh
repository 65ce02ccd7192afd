#!/bin/bash
# Round-end measurement suite, run ON the GPU box (gpurun -- tools/profile_round.sh): the default
# bench line (headline + every BASELINE config as `secondary` entries), the reference arm, the ncu
# launch list of the bench command and one `ncu --set full` capture per dominant kernel.  Everything
# lands in gpurun_out/; profiles/*.txt are produced from the reports afterwards with
# profiles/summarize_ncu.py, tools/ncu_lines.py and tools/collect_profiles.py.
set -x
O=gpurun_out
mkdir -p $O
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_lv_adj_reference_arm.json 2> $O/bench_ref.err
Q="--no-cpu-baseline --no-e2e --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_lv_adj.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $O/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sb_backward$ -c 1 -o $O/ncu_sb_backward_lv -f \
    python bench.py --steps 1 --warmup 1 $Q > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sb_forward$ -c 1 -o $O/ncu_sb_forward_lv -f \
    python bench.py --workload lv_fwd --steps 1 --warmup 1 $Q > $O/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sb_backward$ -c 1 -o $O/ncu_sb_backward_seir -f \
    python bench.py --workload seir_adj --batch 32768 --steps 1 --warmup 1 $Q > $O/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sb_backward_flat$ -c 1 -o $O/ncu_sb_backward_flat_robertson -f \
    python bench.py --workload robertson_adj --steps 1 --warmup 1 $Q > $O/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sb_backward_fund$ -c 1 -o $O/ncu_sb_backward_fund_lv -f \
    python bench.py --backward fundamental --steps 1 --warmup 1 $Q > $O/ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sb_tables$ -c 1 -o $O/ncu_sb_tables_lv -f \
    python bench.py --steps 1 --warmup 1 $Q > $O/ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sb_forward$ -c 1 -o $O/ncu_sb_forward_seir -f \
    python bench.py --workload seir_adj --batch 32768 --steps 1 --warmup 1 $Q > $O/ncu7.log 2>&1
ls -la $O
