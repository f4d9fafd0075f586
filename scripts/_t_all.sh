( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r3c_all_tests.log 2>&1
( time python bench.py > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err ) 2> gpurun_out/r3c_bench.time
( time python bench.py --impl reference > gpurun_out/r3c_bench_ref.json 2> gpurun_out/r3c_bench_ref.err ) 2> gpurun_out/r3c_bench_ref.time
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3c_smoke.log 2>&1
