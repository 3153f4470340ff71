set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_ce_fast.py -x -q -m gpu > gpurun_out/r5d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r5d_tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r5d_bench2.json 2> gpurun_out/r5d_bench2.err
echo done
