# What the driver runs at round end, in one go.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/reh_pytest.log 2>&1; echo "pytest exit $?"
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/reh_smoke.log 2>&1; echo "smoke exit $?"
( time timeout 900 python bench.py --impl reference ) > gpurun_out/reh_bench_ref.json 2> gpurun_out/reh_bench_ref.err; echo "bench ref exit $?"
( time timeout 900 python bench.py ) > gpurun_out/reh_bench.json 2> gpurun_out/reh_bench.err; echo "bench exit $?"
tail -3 gpurun_out/reh_pytest.log; tail -4 gpurun_out/reh_smoke.log; tail -4 gpurun_out/reh_bench_ref.err; head -c 600 gpurun_out/reh_bench_ref.json; echo; tail -4 gpurun_out/reh_bench.err; head -c 400 gpurun_out/reh_bench.json
