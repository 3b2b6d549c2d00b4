#!/bin/bash
# One gpurun call that produces everything profiles/ is built from (run from the repo root on the GPU box):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'
# Each ncu pass runs only after its own command exited 0 without ncu.  Numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $O/pytest_gpu.log
  tail -3 $O/pytest_gpu.log
fi

timeout 600 python bench.py --impl reference > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
echo "bench reference exit $?"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
rc=$?
echo "bench exit $rc"
cat $O/bench_n1.json

if [ $rc -eq 0 ] && [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 300 python tools/prof_run.py --count > $O/prof_run_counts.txt 2>&1
  if timeout 300 python tools/prof_run.py > $O/prof_run.log 2>&1; then
    # one bench step = two waves x 7 passes: ALL 14 launches of each kernel (wave 0 is the sky half of the frame, wave 1 the
    # heavy half -- capturing only the first seven, as round 1 did, misses 62 % of the extend time)
    timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend -c 14 -f -o $O/prof_extend \
        python tools/prof_run.py --integrator=one_sample_mis > $O/ncu_extend.log 2>&1
    echo "ncu extend exit $?"
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 14 -f -o $O/prof_shade \
        python tools/prof_run.py --integrator=one_sample_mis > $O/ncu_shade.log 2>&1
    echo "ncu shade exit $?"
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow -c 4 -f -o $O/prof_shadow \
        python tools/prof_run.py --integrator=mis --spp=16 > $O/ncu_shadow.log 2>&1
    echo "ncu shadow exit $?"
  fi
fi
if [ "${EXTRA_NCU:-0}" = "1" ]; then
  # the two captures behind profiles/rNN_ncu_shade_mis_config4.csv and rNN_ncu_traversal_config5.csv (tools/ncu_stalls.py)
  A4="--scene=multi_light --integrator=mis --spp=16"
  A5="--scene=instanced --integrator=mis --spp=4"
  if timeout 300 python tools/prof_run.py $A4 > $O/prof_run_c4.log 2>&1; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 3 -f -o $O/prof_shade_mis \
        python tools/prof_run.py $A4 > $O/ncu_shade_mis.log 2>&1
    echo "ncu shade (config 4) exit $?"
  fi
  if timeout 300 python tools/prof_run.py $A5 > $O/prof_run_c5.log 2>&1; then
    timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_extend|k_shadow' -c 4 -f -o $O/prof_inst \
        python tools/prof_run.py $A5 > $O/ncu_inst.log 2>&1
    echo "ncu traversal (config 5) exit $?"
  fi
fi
if [ "${REPORT_SCENES:-0}" = "1" ]; then
  timeout 900 python tools/report_scenes.py > $O/report_scenes.jsonl 2> $O/report_scenes.err
  echo "report_scenes exit $?"
fi
ls -la $O | tail -30
