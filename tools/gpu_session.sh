#!/bin/bash
# One GPU-box session: tests, bench, launch lists and ncu captures; everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_session.sh <tag> [steps...]   steps: tests bench qrlist svdlist ovncu gemmshapes
tag=$1; shift
steps="$@"
[ -z "$steps" ] && steps="tests bench"
mkdir -p gpurun_out
for s in $steps; do
  case $s in
    tests) ( time python -m pytest tests -x -q -m gpu ) > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log ;;
    bench) python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json ;;
    benchnocpu) python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json ;;
    qrlist) ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_qr.csv python tools/profile_sweep.py qr > gpurun_out/${tag}_qrlist.log 2>&1; tail -2 gpurun_out/${tag}_qrlist.log ;;
    svdlist) ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_svd.csv python tools/profile_sweep.py svd > gpurun_out/${tag}_svdlist.log 2>&1; tail -2 gpurun_out/${tag}_svdlist.log ;;
    ovncu) ncu --set full --clock-control none --import-source on -k regex:overlap -s 2 -c 1 -f -o gpurun_out/${tag}_overlap python tools/overlap_bench.py 2048 > gpurun_out/${tag}_ovncu.log 2>&1; tail -3 gpurun_out/${tag}_ovncu.log ;;
    gemmshapes) python tools/gemm_shapes.py svd > gpurun_out/${tag}_gemm_shapes_svd.txt 2>&1; python tools/gemm_shapes.py qr > gpurun_out/${tag}_gemm_shapes_qr.txt 2>&1; head -12 gpurun_out/${tag}_gemm_shapes_svd.txt ;;
    benchc4) python bench.py --workload c4 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; tail -c 400 gpurun_out/${tag}_bench_c4.json ;;
    benchc5) python bench.py --workload c5 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; tail -c 400 gpurun_out/${tag}_bench_c5.json ;;
    sanitize) tools/sanitize.sh ${tag} ;;
    gemmncu) ncu --set full --clock-control none --import-source on -k regex:gemm_f64_persistent -s 70 -c 1 -f -o gpurun_out/${tag}_gemm_p1 python tools/profile_sweep.py svd > gpurun_out/${tag}_gemmncu.log 2>&1; tail -2 gpurun_out/${tag}_gemmncu.log ;;
    purifyncu) ncu --set full --clock-control none --import-source on -k regex:purify_fused -s 30 -c 1 -f -o gpurun_out/${tag}_purify python tools/profile_sweep.py svd > gpurun_out/${tag}_purifyncu.log 2>&1; tail -2 gpurun_out/${tag}_purifyncu.log ;;
    purifybncu) ncu --set full --clock-control none --import-source on -k regex:purify_batched -s 3 -c 1 -f -o gpurun_out/${tag}_purify_batched python tools/purify_batched_bench.py > gpurun_out/${tag}_purifybncu.log 2>&1; tail -2 gpurun_out/${tag}_purifybncu.log ;;
    *) echo "unknown step $s" ;;
  esac
done
