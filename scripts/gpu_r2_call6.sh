# round 2, call 6: compute-sanitizer (memcheck / racecheck / synccheck) over the kernel cases, memcheck over one infer
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 120 python scripts/sanitize_cases.py kernels infer ) > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit $?" >> gpurun_out/summary.txt
( timeout 900 $CS --tool memcheck --error-exitcode 9 python scripts/sanitize_cases.py kernels ) > gpurun_out/sanitize_memcheck_kernels.log 2>&1; echo "memcheck kernels exit $?" >> gpurun_out/summary.txt
( timeout 900 $CS --tool synccheck --error-exitcode 9 python scripts/sanitize_cases.py kernels ) > gpurun_out/sanitize_synccheck_kernels.log 2>&1; echo "synccheck kernels exit $?" >> gpurun_out/summary.txt
( timeout 1200 $CS --tool racecheck --error-exitcode 9 python scripts/sanitize_cases.py kernels ) > gpurun_out/sanitize_racecheck_kernels.log 2>&1; echo "racecheck kernels exit $?" >> gpurun_out/summary.txt
( DEPTHPRO_GRAPH=0 timeout 1500 $CS --tool memcheck --error-exitcode 9 python scripts/sanitize_cases.py infer ) > gpurun_out/sanitize_memcheck_infer.log 2>&1; echo "memcheck infer exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; for f in gpurun_out/sanitize_*.log; do echo "== $f"; tail -6 $f | cut -c1-220; done
