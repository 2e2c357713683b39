set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/r2_bench_under_ncu.json 2> $O/r2_bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:stream3d -s 1 -c 1 -o $O/r2_s3_c5_1024 python benchmarks/prof_one.py c5 1024 > $O/ncu_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:extrema2d -s 1 -c 1 -o $O/r2_extrema2d_c4 python benchmarks/prof_one.py c4 64 > $O/ncu_c4.log 2>&1
PROF_ACCUM=1 ncu --set full --clock-control none --import-source on -k regex:stream2d -s 1 -c 1 -o $O/r2_stream2d_c1_f64_fma python benchmarks/prof_one.py c1f64 16 > $O/ncu_c1.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream3d_parity or slab_form or long_separable" > $O/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream3d_parity and symmetric" > $O/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r2_sanitizer_racecheck.log
tail -5 $O/r2_sanitizer_memcheck.log $O/r2_sanitizer_racecheck.log
ls -la $O | tail -12
