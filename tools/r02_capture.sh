#!/bin/bash
# Round-2 evidence run on a 1-GPU box: full GPU test-suite, smoke, ncu --set full captures of the two headline kernels
# (one 128-ply dynamically scheduled launch at steady state, f32; plus the 9x9 kernel with u8 observations), the ncu
# launch list of a short bench run, and the bench lines themselves.  Everything lands in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NCU="ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 9 -c 1 -f"
$NCU -o $O/prof9_f32 python tools/prof_kernel.py --size 9 --boards 65536 --ppl 128 > $O/prof9_f32.log 2>&1
$NCU -o $O/prof9_u8 python tools/prof_kernel.py --size 9 --boards 65536 --ppl 128 --obs u8 > $O/prof9_u8.log 2>&1
$NCU -o $O/prof19_f32 python tools/prof_kernel.py --size 19 --boards 16384 --ppl 128 > $O/prof19_f32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_9x9.csv \
    python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/bench_9x9_steps20_warmup5.json 2> $O/bench_s20.err
python bench.py > $O/bench_9x9_default.json 2> $O/bench_default.err
python bench.py --workload 19x19 --steps 20 --warmup 5 > $O/bench_19x19.json 2> $O/bench_19.err
for obs in u8 bf16; do
  python bench.py --obs $obs --steps 20 --warmup 5 --quick --no-cpu-baseline > $O/bench_9x9_$obs.json 2>> $O/bench_s20.err
  python bench.py --workload 19x19 --obs $obs --steps 20 --warmup 5 --quick --no-cpu-baseline > $O/bench_19x19_$obs.json 2>> $O/bench_s20.err
done
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_9x9.json 2>> $O/bench_s20.err
python bench.py --impl reference --workload 19x19 --steps 20 --warmup 5 > $O/bench_reference_19x19.json 2>> $O/bench_s20.err
python tools/kernel_ab.py --quick --out $O/kernel_ab_final.json > $O/kernel_ab_final.log 2>&1
cuobjdump -sass gymgo_b200/_lib/libgymgo_b200.so 2>/dev/null | grep -c ATOMS > $O/atoms_count.txt
tail -c 300 $O/bench_s20.err $O/bench_default.err $O/bench_19.err
ls -la $O | tail -20
python tools/e2e_phases.py > $O/e2e_phases.json 2>> $O/bench_s20.err
python tools/e2e_phases.py --size 19 --boards 16384 > $O/e2e_phases_19x19.json 2>> $O/bench_s20.err
