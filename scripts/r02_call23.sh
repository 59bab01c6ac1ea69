#!/usr/bin/env bash
# Round-2 call 23: validation of the committed build of record - smoke(), the whole GPU suite, the default bench line
# (roofline.traffic now comes from the r02t capture of the same kernel sources) and the reference arm.
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02w_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02w_smoke.log
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider -rs > gpurun_out/r02w_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02w_gpu_tests.log
tail -10 gpurun_out/r02w_gpu_tests.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r02w_bench_n1.err | cut -c1-300; cut -c1-300 gpurun_out/r02w_bench_n1.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02w_bench_ref.json 2> gpurun_out/r02w_bench_ref.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/r02w_bench_ref.json
