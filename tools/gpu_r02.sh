#!/bin/bash
# One gpurun call of round-2 evidence: tests, smoke, bench, layer/operator reports, ncu captures.
# usage: tools/gpu_r02.sh <tag> [stage ...]   (default: all stages); logs under gpurun_out/
TAG=${1:-r02}; shift
STAGES=${@:-"tests smoke bench layers ops ncu_list ncu_all ncu_src ncu_ops san"}
mkdir -p gpurun_out
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name: $*"
  timeout $to "$@" > gpurun_out/${name}_$TAG.log 2>&1
  local rc=$?
  echo "=== $name rc=$rc"; tail -n ${TAILN:-6} gpurun_out/${name}_$TAG.log
  return $rc
}
for s in $STAGES; do
  case $s in
    tests)  TAILN=40 run tests 1500 python -m pytest tests -q -m gpu --timeout 600 -rxXs -s ;;
    testsx) TAILN=40 run tests 1500 python -m pytest tests -q -m gpu --timeout 600 -x -rxX ;;
    smoke)  run smoke 300 python __graft_entry__.py smoke ;;
    bench)  TAILN=3 run bench 900 python bench.py --steps 20 --warmup 5 ;;
    benchq) TAILN=3 run bench 900 python bench.py --steps 20 --warmup 5 --no-configs --no-eager --no-cpu-baseline ;;
    layers) TAILN=22 run layers 300 python tools/layer_report.py ;;
    layers6) TAILN=22 run layers6 300 python tools/layer_report.py --batch 6 ;;
    ops)    TAILN=18 run ops 300 python tools/op_bench.py ;;
    ncu_list)  # launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
          python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-eager --no-sustained > gpurun_out/ncu_list_$TAG.log 2>&1
      echo "=== ncu_list rc=$?"; tail -2 gpurun_out/ncu_list_$TAG.log | cut -c1-300 ;;
    ncu_all)   # full capture (no source) of every launch of one step: 21 kernels after 3 warm-up steps; summarised on the box
      ncu --set full --clock-control none \
          -k regex:'conv_umma_kernel|conv_smerge_kernel|conv_first_umma|conv_last|conv_ups4|adain_fold|adain_nhwc' -s 63 -c 21 -f \
          -o gpurun_out/prof_conv_$TAG python tools/layer_report.py --iters 1 --batch 32 > gpurun_out/ncu_all_$TAG.log 2>&1
      echo "=== ncu_all rc=$?"
      python tools/ncu_summary.py gpurun_out/prof_conv_$TAG.ncu-rep > gpurun_out/ncu_conv_summary_$TAG.txt 2>&1
      python tools/ncu_traffic.py gpurun_out/prof_conv_$TAG.ncu-rep 32 > gpurun_out/ncu_conv_traffic_$TAG.json 2>gpurun_out/ncu_traffic_$TAG.err
      cat gpurun_out/ncu_conv_summary_$TAG.txt | cut -c1-200
      rm -f gpurun_out/prof_conv_$TAG.ncu-rep ;;
    ncu_src)   # the five kernels below the roofline, with source: conv1_1, conv1_2, dec7, dec8, dec9
      ncu --set full --clock-control none --import-source on \
          -k regex:'conv_smerge_kernel|conv_first_umma|conv_last|conv_ups4' -s 15 -c 5 -f \
          -o gpurun_out/prof_src_$TAG python tools/layer_report.py --iters 1 --batch 32 > gpurun_out/ncu_src_$TAG.log 2>&1
      echo "=== ncu_src rc=$?"; ls -la gpurun_out/prof_src_$TAG.ncu-rep ;;
    ncu_ops)   # HBM operators at [32,512,64,64]: calc_mean_std (MODE 0), AdaIN (MODE 2), Welford accumulate (one launch)
      ncu --set full --clock-control none --import-source on -k regex:'plane_bulk_kernel|welford_bulk_kernel|merge_planes' -c 12 -f \
          -o gpurun_out/prof_ops_$TAG python tools/op_bench.py --iters 1 > gpurun_out/ncu_ops_$TAG.log 2>&1
      echo "=== ncu_ops rc=$?"
      python tools/ncu_summary.py gpurun_out/prof_ops_$TAG.ncu-rep > gpurun_out/ncu_ops_summary_$TAG.txt 2>&1
      python tools/ncu_ops_traffic.py gpurun_out/prof_ops_$TAG.ncu-rep > gpurun_out/ncu_ops_traffic_$TAG.json 2>gpurun_out/ncu_ops_traffic_$TAG.err
      cat gpurun_out/ncu_ops_summary_$TAG.txt | cut -c1-220 ;;
    san)    # compute-sanitizer memcheck: the stats / AdaIN operators (cp.async.bulk + mbarrier rings) and one small
            # style_transfer per engine
      TAILN=12 run san_ops 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_ops.py -q -m gpu -x --timeout 800
      TAILN=12 run san_net32 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/san_small.py fp32
      TAILN=12 run san_net16 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/san_small.py fp16
      TAILN=12 run san_net16x3 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/san_small.py fp16x3
      TAILN=12 run san_netb16x3 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/san_small.py bf16x3 ;;
  esac
done
du -sh gpurun_out
