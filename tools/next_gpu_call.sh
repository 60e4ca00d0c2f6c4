#!/bin/bash
# First GPU call of the next round (run as: gpurun --timeout 900 -- 'bash tools/next_gpu_call.sh').
# 1. the parity tests that have not run on hardware yet, with their real outcomes (--runxfail shows tracebacks);
# 2. the whole GPU suite as the driver runs it;
# 3. one short bench line.
# Everything lands in gpurun_out/ (merged back into the repo's gpurun_out/ by gpurun).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_zz_unverified_modes.py tests/test_oracle_numpy_step.py tests/test_dist_gpu.py -m gpu -q --runxfail --tb=short -p no:cacheprovider > gpurun_out/pending_modes.log 2>&1
echo "pending modes: rc=$?"; tail -15 gpurun_out/pending_modes.log
timeout 500 python -m pytest tests -q -m gpu -x -rxX -p no:cacheprovider > gpurun_out/gpu_suite.log 2>&1
echo "gpu suite: rc=$?"; tail -8 gpurun_out/gpu_suite.log
timeout 200 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
echo "bench: rc=$?"; cut -c1-600 gpurun_out/bench_short.json
# 3b. A/B of the 4-row schedule of the sweep kernels
ASPH_ROWS4=1 timeout 200 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_short_rows4.json 2> gpurun_out/bench_short_rows4.err
echo "bench ROWS4: rc=$?"; cut -c1-600 gpurun_out/bench_short_rows4.json
# 3c. A/B of the bulk-copy stage fill
ASPH_BULK=1 timeout 200 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_short_bulk.json 2> gpurun_out/bench_short_bulk.err
echo "bench BULK: rc=$?"; cut -c1-600 gpurun_out/bench_short_bulk.json
# 4. the adaptive workloads that have no number yet: BASELINE configs[2] (4 M particles) and the 16 M north-star case
timeout 600 python tools/bench_adaptive.py --spacing 5.612e-4 --warmup 60 --steps 40 > gpurun_out/adaptive_4m.json 2> gpurun_out/adaptive_4m.err
echo "adaptive 4M: rc=$?"; cut -c1-700 gpurun_out/adaptive_4m.json
timeout 900 python tools/bench_adaptive.py --spacing 2.806e-4 --warmup 40 --steps 20 > gpurun_out/adaptive_16m.json 2> gpurun_out/adaptive_16m.err
echo "adaptive 16M: rc=$?"; cut -c1-700 gpurun_out/adaptive_16m.json
# 5. launch lists (per-kernel durations, cold and serialised under ncu: shares only) of the default and the 4-row schedule
for v in 0 1; do
  ASPH_ROWS4=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rows4_$v.csv \
    python tools/profile_step.py 1.122e-3 2 > gpurun_out/profile_rows4_$v.log 2>&1
  echo "ncu launch list ROWS4=$v: rc=$?"
done
# 6. (separate call, `gpurun --gpus 2`): the 2-GPU parity tests incl. the columns layout and the 4-row peer kernels, then the
#    two weak-scaling scenes side by side:
#      python -m pytest tests/test_dist_gpu.py tests/test_zz_unverified_modes.py -m gpu -q --runxfail -k "two_gpu"
#      for s in wide columns; do ASPH_BENCH_SCENE=$s python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#        --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 16 --warmup 3; done
