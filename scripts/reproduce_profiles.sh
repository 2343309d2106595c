#!/bin/bash
# The GPU-side commands behind the numbers in DESIGN.md / profiles/ (round 1), as they were run through
#   /usr/local/graft/bin/gpurun [--gpus N] --timeout T -- '<command>'
# Every ncu command follows a plain run of the same command line that exited 0.  Outputs go to gpurun_out/ (scratch);
# the summaries under profiles/ were cut from them with `ncu -i <rep> --page raw --csv`.
set -e
mkdir -p gpurun_out

# 1. parity: all GPU tests (218), smoke
python -m pytest tests -m gpu -q
python -c "import __graft_entry__ as g; g.smoke()"

# 2. the bench line (N = 1) -> profiles/bench_r01_k7_1gpu.json
python bench.py > gpurun_out/bench_final.json

# 3. launch list of the timed region -> profiles/launches_r01c.csv (+ _summary.csv)
python bench.py --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2500 --csv \
    --log-file gpurun_out/launches_r01c.csv python bench.py --no-cpu > gpurun_out/ncu_launches.log 2>&1

# 4. full captures of the hot kernels at k = 7 -> profiles/hot_kernels_full_r01d_summary.txt
python scripts/ncu_probe.py 7 amg > gpurun_out/ncu_probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_star_op -c 4 -o gpurun_out/prof_starop_r01d \
    python scripts/ncu_probe.py 7 amg > gpurun_out/ncu_probe2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_jacobian -c 1 -o gpurun_out/prof_jac_r01d \
    python scripts/ncu_probe.py 7 > gpurun_out/ncu_probe3.log 2>&1

# 5. multigrid variants (V(1,1), V(3,3), Chebyshev, omega, pre/post, Galerkin instead of re-discretised coarse operators)
bash scripts/bench_variants.sh "" "--prec-steps 1" "--prec-steps 3" "--solver-opt amg_smoother=1" "--solver-opt amg_omega=0.8" \
     "--solver-opt amg_pre_steps=1 --solver-opt amg_post_steps=2" "--solver-opt amg_rediscretise=0"

# 6. strong scaling (gpurun --gpus N): -> profiles/bench_r01_k7_{2,8}gpu.json
# python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
#     bench.py --gpus N --steps 3 --warmup 3 --no-cpu
