#!/bin/bash
# One gpurun call that refreshes the round-2 evidence: launch list + full ncu capture of the dominant
# kernel (profiles/summarize.py turns them into r02_summary.md), the bench line, the other-config
# timings, the FAB-loss step timing, and compute-sanitizer runs of the kernels added this round.
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python profiles/profile_chain.py --chains 1 > gpurun_out/ncu_list_r02.log 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:k_hmc_step_u -s 20 -c 1 -f \
    -o gpurun_out/hmc_step_u_r02 python profiles/profile_chain.py > gpurun_out/ncu_full_r02.log 2>&1
timeout -k 5 400 python bench.py 2> gpurun_out/bench_r02.err | grep '^{' | head -1 > gpurun_out/bench_r02.json
timeout -k 5 400 python profiles/bench_configs.py 2>/dev/null | grep '^{' > gpurun_out/r02_bench_configs.jsonl
timeout -k 5 200 python profiles/bench_param_grad.py > gpurun_out/r02_param_grad.log 2>&1
{
  echo "== memcheck: row-tile engine + parameter-gradient kernels"
  timeout -k 5 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rowtile.py tests/test_gpu_param_grad.py \
      -x -q -k "not 2048 and not full and not config" 2>&1 | grep -v "Warning\|warn" | tail -6
  echo "== racecheck: row-tile engine (small cases) + parameter-gradient kernels (small cases)"
  timeout -k 5 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_rowtile.py tests/test_gpu_param_grad.py \
      -x -q -k "(flow_logprob and (1-2-64 or 2-10-200)) or (hmc_transition and 3-4-4) or ragged or odd or repeat" 2>&1 | grep -v "Warning\|warn" | tail -25
} > gpurun_out/sanitizer_r02c.log 2>&1
tail -3 gpurun_out/ncu_list_r02.log; tail -2 gpurun_out/ncu_full_r02.log; cut -c1-300 gpurun_out/bench_r02.json; cat gpurun_out/r02_bench_configs.jsonl | cut -c1-400; cat gpurun_out/r02_param_grad.log | tail -4; cat gpurun_out/sanitizer_r02c.log
