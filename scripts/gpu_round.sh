#!/bin/bash
# One GPU-box session: parity tests, A/B timings, launch list, full ncu captures, bench.  Outputs under gpurun_out/.
# usage (through gpurun): bash scripts/gpu_round.sh <tag> [steps...]   steps default: tests probe list full trace bench
set -u
TAG=${1:-run}; shift || true
STEPS=${*:-tests probe list full trace bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for s in $STEPS; do
  case $s in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests exit $?" >> $OUT/tests.log; tail -5 $OUT/tests.log;;
    probe) timeout 600 python tests/tools/gpu_probe.py 256 > $OUT/probe_tile.log 2>&1; tail -25 $OUT/probe_tile.log
           DMX_ASM_LEGACY=1 timeout 600 python scripts/ilu_probe.py 256 5 > $OUT/probe_legacy.log 2>&1; tail -5 $OUT/probe_legacy.log;;
    list)  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
             python scripts/profile_step.py 256 6 > $OUT/list.log 2>&1; tail -2 $OUT/list.log;;
    listbench) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv \
             python bench.py --steps 2 --warmup 1 --no-strong --no-amg --no-cpu-baseline --no-parity-check > $OUT/listbench.log 2>&1; tail -2 $OUT/listbench.log;;
    full)  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"assemble_tile|ilu_sweep|ilu_diag|ilu_skew|stencil_spmv|vec_skew|axpy3|axpy_r|p_update|dot_kernel" \
             -c 16 -f -o $OUT/prof python scripts/profile_step.py 256 2 > $OUT/full.log 2>&1; tail -2 $OUT/full.log;;
    fullamg) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"amg_|stencil_spmv|ilu_sweep|vec_skew" \
             -c 60 -f -o $OUT/prof_amg python scripts/profile_step.py 256 1 amg > $OUT/fullamg.log 2>&1; tail -2 $OUT/fullamg.log;;
    trace) timeout 300 python scripts/sweep_trace.py 256 > $OUT/trace.log 2>&1; tail -30 $OUT/trace.log;;
    bench) timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json;;
    stream) nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_probe scripts/probes/stream_probe.cu && timeout 600 /tmp/stream_probe > $OUT/stream.log 2>&1; tail -80 $OUT/stream.log;;
    asm)   timeout 900 python -m pytest tests/test_gpu_assembly.py -m gpu -x -q > $OUT/asm_tests.log 2>&1; tail -3 $OUT/asm_tests.log
           timeout 600 python scripts/ilu_probe.py 256 5 > $OUT/probe_asm.log 2>&1; tail -5 $OUT/probe_asm.log;;
    asmfull) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"assemble_tile" -c 1 -f -o $OUT/prof_asm python scripts/profile_step.py 256 1 > $OUT/asmfull.log 2>&1; tail -2 $OUT/asmfull.log;;
    lin)   timeout 300 python -m pytest tests/test_gpu_linear.py tests/test_gpu_newton.py -m gpu -x -q > $OUT/lin_tests.log 2>&1; tail -3 $OUT/lin_tests.log
           timeout 600 python scripts/ilu_probe.py 256 10 > $OUT/probe_ilu.log 2>&1; tail -5 $OUT/probe_ilu.log;;
    new)   timeout 600 python -m pytest tests/test_tracer.py tests/test_gpu_compressible.py tests/test_abi_exports.py -m gpu -x -q > $OUT/new_tests.log 2>&1; tail -15 $OUT/new_tests.log;;
    smoke) timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log;;
  esac
done
