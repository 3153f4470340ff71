set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r5c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r5c_tests.log
timeout 300 python bench.py > gpurun_out/r5c_bench.json 2> gpurun_out/r5c_bench.err
echo done
