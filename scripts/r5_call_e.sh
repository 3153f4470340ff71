set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r5e_smoke.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r5e_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r5e_tests.log
timeout 300 python bench.py > gpurun_out/r5e_bench.json 2> gpurun_out/r5e_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5e_ref.json 2> gpurun_out/r5e_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r5e_launches.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-cpu-baseline --child > gpurun_out/r5e_ncu_bench.log 2>&1
timeout 200 python bench.py --workload versa_large --steps 5 --reps 3 --no-cpu-baseline > gpurun_out/r5e_versa_large.json 2> gpurun_out/r5e_versa_large.err
echo done
