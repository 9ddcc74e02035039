#!/bin/bash
# GPU-box run of round 2: full GPU test suite, the full bench line, the launch list of a short bench run and one
# `ncu --set full` capture of the dominant kernel per tensor-core mode
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_full.log 2>&1; tail -3 gpurun_out/r02_pytest_full.log
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; tail -c 300 gpurun_out/r02_bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train > /dev/null 2>&1
for m in bf16x3 fp16f8 bf16; do
  ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 3 -c 1 -o gpurun_out/r02_prof_render_$m \
      python profiles/run_render_points.py 65536 $m > gpurun_out/r02_ncu_$m.log 2>&1
done
ncu --set full --clock-control none -k regex:render_tail_kernel\|coarse_to_fine_kernel\|ray_head_kernel -s 6 -c 3 \
    -o gpurun_out/r02_prof_fused_ray_kernels python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train --no-extra > /dev/null 2>&1
ls -la gpurun_out | tail -12
