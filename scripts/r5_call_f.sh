set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 85 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --reps 5 --no-host-e2e --no-cpu-baseline > gpurun_out/r5f_bench4.json 2> gpurun_out/r5f_bench4.err
echo done
