#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list and --set full captures.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh [tests] [bench] [ref] [launches] [full]'
# Everything is written under gpurun_out/ (merged back into the build container).
set -u
OUT=gpurun_out
mkdir -p $OUT
WHAT="${*:-micro tests bench ref launches full}"
has() { [[ " $WHAT " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1

if has micro; then
  ./scripts/micro/mma_shapes.bin > $OUT/mma_shapes.txt 2>&1
  cat $OUT/mma_shapes.txt
  [ -x scripts/micro/gather_peak.bin ] && timeout 60 ./scripts/micro/gather_peak.bin > $OUT/gather_peak.txt 2>&1 && cat $OUT/gather_peak.txt
  [ -x scripts/micro/tmem_ld.bin ] && timeout 60 ./scripts/micro/tmem_ld.bin > $OUT/tmem_ld.txt 2>&1 && cat $OUT/tmem_ld.txt
fi
if has tests; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
  ( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1
  tail -3 $OUT/smoke.log
fi
if has trace; then
  LNB200_LIB=lidar-nerf_b200/lib/liblnb200_trace.so timeout 300 python scripts/diag_bwd_trace.py > $OUT/bwd_trace.txt 2>&1
  tail -2 $OUT/bwd_trace.txt
fi
if has overlap; then
  timeout 300 python scripts/diag_overlap.py > $OUT/overlap.txt 2>&1
  tail -25 $OUT/overlap.txt
fi
if has bench; then
  ( time timeout 600 python bench.py ) > $OUT/bench.log 2> $OUT/bench.err
  tail -1 $OUT/bench.log
fi
if has ref; then
  ( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference.log 2> $OUT/bench_reference.err
  tail -1 $OUT/bench_reference.log
  ( time timeout 400 python bench.py --impl reference-cuda --steps 30 ) > $OUT/bench_reference_cuda.log 2> $OUT/bench_reference_cuda.err
  tail -1 $OUT/bench_reference_cuda.log
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 320 --no-graph --no-cpu-baseline --no-profile --no-api-b2 --no-reference-cuda --profiler-range > $OUT/launches_bench.log 2>&1
  echo "launch list rows: $(wc -l < $OUT/launches.csv)"
fi
if has full; then
  # the heavy kernels of one steady-state step (eager launches; skip the preparation phase's launches)
  timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_field_fused_fwd|k_field_fwd|k_mlp_bwd|k_grid_fwd|k_grid_bwd|k_march_train|k_adam|k_composite_train|k_lidar_composite_step|k_ray_dir_terms|k_pack_field_weights' \
    --profile-from-start off -c ${NCU_COUNT:-12} -f -o $OUT/full \
    python bench.py --steps 2 --warmup 320 --no-graph --no-cpu-baseline --no-profile --no-api-b2 --no-reference-cuda --profiler-range > $OUT/full_bench.log 2>&1
  ls -la $OUT/full.ncu-rep
fi
