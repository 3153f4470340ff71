set -x
cd $GRAFT_REPO_ROOT
LR_GENS=3 LR_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"san_chain3_fwd_kernel|lr_rank_kernel|lr_combine_kernel" -s 3 -c 3 -o gpurun_out/r02f_chain3 python scripts/lr_check.py 512 1.0 > gpurun_out/r02f_ncu1.log 2>&1
LR_GENS=3 LR_REPS=1 timeout 400 ncu --set full --clock-control none -k regex:umma_gemm_kernel -s 9 -c 1 -o gpurun_out/r02f_wgrad python scripts/lr_check.py 512 1.0 > gpurun_out/r02f_ncu2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
IISAN_B200_LIB=$PWD/iisan_b200/lib/libiisan_b200_trace.so timeout 200 python scripts/chain3_trace.py 512 > gpurun_out/r02f_chain3_trace.json 2> gpurun_out/r02f_trace.err
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1
echo done
