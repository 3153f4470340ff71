set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 60 ./iisan_b200/lib/probe_gather4 > gpurun_out/r5a_probe_gather4.jsonl 2>&1
timeout 300 python -m pytest tests/test_gpu_ce_fast.py tests/test_gpu_store.py tests/test_gpu_pipeline.py tests/test_gpu_parity.py tests/test_gpu_user_encoder.py -x -q -m gpu > gpurun_out/r5a_tests.log 2>&1
timeout 200 python scripts/ce_gather_bench.py > gpurun_out/r5a_ce_gather.json 2> gpurun_out/r5a_ce_gather.err
timeout 300 python bench.py > gpurun_out/r5a_bench.json 2> gpurun_out/r5a_bench.err
echo done
