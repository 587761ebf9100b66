# configs coverage + bulk-staging parity (ONE gpurun command)
out=gpurun_out; mkdir -p $out
PHI_GPU_READ_BULK=1 timeout 900 python -m pytest tests -m gpu -x -q > $out/g6_pytest_bulk.log 2>&1; echo "pytest (bulk staging on) rc=$?"; tail -2 $out/g6_pytest_bulk.log
timeout 600 python bench.py --config readme > $out/g6_bench_readme.json 2> $out/g6_bench_readme.err; echo "readme rc=$?"
timeout 600 python bench.py --config c2 --coverage 0.1 > $out/g6_bench_c2_cov0.1.json 2> $out/g6_bench_c2_cov0.1.err; echo "c2 0.1x rc=$?"
timeout 600 python bench.py --config c2 --coverage 1 > $out/g6_bench_c2_cov1.json 2> $out/g6_bench_c2_cov1.err; echo "c2 1x rc=$?"
timeout 900 python bench.py --config c3long > $out/g6_bench_c3long.json 2> $out/g6_bench_c3long.err; echo "c3long rc=$?"
timeout 900 python bench.py --impl reference --config c2 --steps 2 --warmup 1 > $out/g6_reference_c2.json 2> $out/g6_reference_c2.err; echo "reference c2 rc=$?"
PHI_ADAPTER_TIMES=1 timeout 600 python profiles/readme_dropin_times.py > $out/g6_readme_dropin.json 2> $out/g6_readme_dropin.err; echo "dropin rc=$?"
