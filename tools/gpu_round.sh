#!/bin/bash
# One gpurun call that produces everything profiles/ is built from (run from the repo root on the GPU box):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'
# Each ncu pass runs only after its own command (tools/prof_run.py) exited 0 without ncu; the bench arms run last.  Numbers printed under ncu are never bench values.
# The .ncu-rep files are turned into raw CSV pages ON THE BOX and deleted (gpurun_out/ may carry 64 MiB back); set
# KEEP_REP=extend|shade|shadow to keep one report (for tools/ncu_lines.py's per-source-line view).
# Switches: SKIP_TESTS=1, SKIP_BENCH=1, SKIP_NCU=1, EXTRA_NCU=1 (config-4 shade, config-5 traversal), TESTS="-k expr"
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 2400 python -m pytest tests -m gpu -x -q ${TESTS:-} > $O/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $O/pytest_gpu.log
  tail -5 $O/pytest_gpu.log
fi

capture() {  # capture <name> <ncu kernel regex> <count> <timeout> <prof_run args...>
  local name=$1 rx=$2 cnt=$3 to=$4; shift 4
  # (TAKE_PROVISIONAL=0: the first render waits for the tie-break ranks, so the captured kernels are the regular ones and not
  #  the tie-counting instantiations a render ahead of the reference tree uses)
  TAKE_PROVISIONAL=0 timeout $to ncu --set full --clock-control none --import-source on -k "regex:$rx" -c $cnt -f -o $O/prof_$name \
      python tools/prof_run.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  if [ -f $O/prof_$name.ncu-rep ]; then
    ncu -i $O/prof_$name.ncu-rep --page raw --csv > $O/prof_${name}_raw.csv 2>/dev/null
    if [ "${KEEP_REP:-}" != "$name" ]; then rm -f $O/prof_$name.ncu-rep; fi
  fi
}

if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scenes '' > $O/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 300 python tools/prof_run.py --count > $O/prof_run_counts.txt 2>&1
  if timeout 300 python tools/prof_run.py > $O/prof_run.log 2>&1; then
    # one bench step = two waves x 7 passes: ALL 14 launches of each kernel (wave 0 is the sky half of the frame, wave 1 the
    # heavy half -- capturing only the first seven, as round 1 did, misses 62 % of the extend time)
    capture extend k_extend 14 1500 --integrator=one_sample_mis
    capture shade k_shade 14 1200 --integrator=one_sample_mis
    capture shadow k_shadow 4 600 --integrator=mis --spp=16
  fi
fi
if [ "${EXTRA_NCU:-0}" = "1" ]; then
  A4="--scene=multi_light --integrator=mis --spp=16"
  A5="--scene=instanced --integrator=mis --spp=4"
  if timeout 300 python tools/prof_run.py $A4 > $O/prof_run_c4.log 2>&1; then
    capture shade_mis_config4 k_shade 7 900 $A4
  fi
  if timeout 300 python tools/prof_run.py $A5 > $O/prof_run_c5.log 2>&1; then
    capture traversal_config5 'k_extend|k_shadow' 8 900 $A5
  fi
fi
# the bench arms come AFTER the captures: tools/ncu_summary.py stamps profiles/ncu_metrics.json with the hash of the kernel
# sources on this box, and bench.py quotes ncu figures only from a capture of the sources it runs
if [ "${SKIP_NCU:-0}" != "1" ]; then python tools/ncu_summary.py ${TAG:-r02} > $O/ncu_summary.log 2>&1; cp profiles/ncu_metrics.json $O/ 2>/dev/null; fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py --impl reference > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
  echo "bench reference exit $?"
  timeout 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
  rc=$?
  echo "bench exit $rc"
  python - <<'EOF'
import json
try:
    d = json.load(open("gpurun_out/bench_n1.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "scene_create_ms")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
    for r in d.get("scenes", []):
        print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items() if k in ("key", "job_ms", "mrays_per_s", "samples_per_s", "scene_create_ms", "e2e_job_ms", "e2e_job_cold_ms", "provisional", "error")})
except Exception as e:
    print("bench line unreadable:", e)
EOF
fi

ls -la $O | tail -30
du -sh $O
