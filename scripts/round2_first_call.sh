#!/bin/bash
# First GPU call of round 2 (one B200, ~4 minutes of box time):  gpurun --timeout 420 -- 'bash scripts/round2_first_call.sh'
# 1. the whole GPU suite (includes the tests added after the round-1 budget was spent: poisoned-workspace probe of the forward,
#    Versa real shapes with the calibrated emulation bars, the reference's batch loop under DDP / autocast / GradScaler)
# 2. smoke()
# 3. measurement #1 of DESIGN.md 4.1: hidden-state tile loads alone, strided vs tile-contiguous
# 4. the bench line (no CPU leg) and the ncu launch list of the same command
# 5. one full ncu capture of the CE kernels (instruction mix per logit after the snapshot-c changes)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -3 gpurun_out/r2a_tests.log
python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2a_smoke.log
python scripts/probe_tile_stream.py 512 > gpurun_out/r2a_probe_tile_stream.jsonl 2> gpurun_out/r2a_probe_tile_stream.err; echo "probe rc=$?"
tail -4 gpurun_out/r2a_probe_tile_stream.jsonl
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-store --no-graph > gpurun_out/r2a_ncu_bench.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ce_tile -s 6 -c 2 -o gpurun_out/r2a_ce \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-store --no-graph > gpurun_out/r2a_ncu_ce.log 2>&1; echo "ncu ce rc=$?"
# 6. (only if the trace variant was built before the call: IISAN_B200_BUILD_VARIANT=trace python -m iisan_b200.build)
#    measurement #2 of DESIGN.md 4.1: where the chain kernels wait, per role and wait site
if [ -f iisan_b200/lib/libiisan_b200_trace.so ]; then
  IISAN_B200_LIB=$PWD/iisan_b200/lib/libiisan_b200_trace.so python scripts/chain_trace.py 512 > gpurun_out/r2a_chain_trace.json 2> gpurun_out/r2a_chain_trace.err; echo "trace rc=$?"
fi
