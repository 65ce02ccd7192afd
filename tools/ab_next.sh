#!/bin/bash
# A/B measurements queued at the end of round 1, when the GPU budget was already spent: run ON the
# GPU box (gpurun --timeout 1500 -- tools/ab_next.sh), results in gpurun_out/ab_*.json.
#  1. the option tests that have only run on the emulated device code so far;
#  2. the headline workload compiled by the toolkit's NVRTC 12.9 (default) against torch's
#     NVRTC 12.8 (the LV cubin used to come from whichever the process had loaded);
#  3. the reference README's backward-tolerance override (1e-8 instead of the hard-coded 1e-10,
#     SURVEY.md 8(d) asks for both), the restart-free fundamental-matrix backward pass
#     (--backward fundamental: ~6x fewer backward steps on LV, one lane per instance, spills
#     328 B -- the number that decides whether the lane-per-column layout is worth building)
#     and Hermite interpolation, device-resident legs only.
set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_zz_options_gpu.py -q -m gpu > $O/ab_option_tests.log 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $Q > $O/ab_lv_nvrtc129.json 2> $O/ab_lv_nvrtc129.err
NVRTC128=$(python -c "import nvidia.cuda_nvrtc, os; print(os.path.join(list(nvidia.cuda_nvrtc.__path__)[0], 'lib', 'libnvrtc.so.12'))")
SUNODE_B200_NVRTC=$NVRTC128 python bench.py $Q > $O/ab_lv_nvrtc128.json 2> $O/ab_lv_nvrtc128.err
SUNODE_B200_NVRTC=$NVRTC128 python bench.py --workload seir_adj --batch 32768 $Q > $O/ab_seir_nvrtc128.json 2> $O/ab_seir_nvrtc128.err
python bench.py --workload seir_adj --batch 32768 $Q > $O/ab_seir_nvrtc129.json 2> $O/ab_seir_nvrtc129.err
python bench.py --backward-tol 1e-8 $Q > $O/ab_lv_bwdtol1e-8.json 2> $O/ab_lv_bwdtol.err
python bench.py --backward fundamental $Q > $O/ab_lv_fundamental.json 2> $O/ab_lv_fundamental.err
python bench.py --interpolation hermite $Q > $O/ab_lv_hermite.json 2> $O/ab_lv_hermite.err
ncu --set full --clock-control none --import-source on -k regex:^sb_backward_fund$ -c 1 \
    -o $O/ncu_sb_backward_fund_lv -f \
    python bench.py --backward fundamental --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ab_ncu_fund.log 2>&1
for f in $O/ab_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['value'], d['roofline']['kernel_ms_all'])
except Exception as e:
    print(sys.argv[1], 'unreadable:', e)
PY
done
