python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest12.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench12_ref.json 2>> gpurun_out/r2_bench12.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke12.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke12.log
