set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --reps 4 --no-store --no-host-e2e --no-cpu-baseline > gpurun_out/r5b_bench2_default.json 2> gpurun_out/r5b_bench2_default.err
NCCL_MAX_CTAS=8 timeout 200 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --reps 4 --no-store --no-host-e2e --no-cpu-baseline > gpurun_out/r5b_bench2_ctas8.json 2> gpurun_out/r5b_bench2_ctas8.err
echo done
