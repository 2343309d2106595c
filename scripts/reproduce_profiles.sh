#!/bin/bash
# The GPU-side commands behind the numbers in DESIGN.md / profiles/ (round 2; round 1's are in the git history), as they
# were run through   /usr/local/graft/bin/gpurun [--gpus N] --timeout T -- '<command>'
# Every ncu command follows a plain run of the same command line that exited 0.  Outputs go to gpurun_out/ (scratch);
# the summaries under profiles/ were cut from them with `ncu -i <rep> --page raw --csv` / `--page source --csv`.
set -e
mkdir -p gpurun_out

# 1. parity: all GPU tests (275 on one GPU), smoke
python -m pytest tests -m gpu -q
python -c "import __graft_entry__ as g; g.smoke()"

# 2. the bench line (N = 1, with the CPU sample) -> profiles/bench_r02_k7_1gpu.json
python bench.py > gpurun_out/r2_bench_final.json

# 3. launch list of one timed step -> profiles/launches_r02.csv (+ launches_r02_summary.csv)
python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r02b.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_launch.log 2>&1

# 4. full capture of the streaming SpMV at k = 7 -> profiles/hot_kernels_full_r02_summary.txt
python scripts/ncu_probe.py 7 amg > gpurun_out/r2_probe_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_star_op_tma -s 2 -c 4 -o gpurun_out/r2_prof_tma_final \
    python scripts/ncu_probe.py 7 amg > gpurun_out/r2_ncu_final.log 2>&1
#    (the first streaming version, source-level: profiles/spmv_tma_r02_summary.txt came from the same command with -c 4)

# 5. SpMV kernel variants alone (plain-load kernel, stages, lanes per row) and the stand-alone feed probe -> profiles/stream_probe_r02.txt
python scripts/spmv_probe.py 7 tma=0 tma=1,tma_stages=2,tma_lpr=1 tma=1,tma_stages=2,tma_lpr=2 tma=1,tma_stages=3,tma_lpr=2
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/probes/stream_probe scripts/probes/stream_probe.cu
scripts/probes/stream_probe

# 6. multigrid variants (smoothing steps, damping, size of the dense coarsest level)
bash scripts/bench_variants_r02.sh "amg_pre_steps=2 amg_post_steps=1" "amg_pre_steps=1 amg_post_steps=2" "amg_pre_steps=3 amg_post_steps=1" \
     "amg_pre_steps=1 amg_post_steps=1" "amg_pre_steps=2 amg_post_steps=2 amg_omega=0.8" "amg_dense_max=2048" "amg_dense_max=1024" "amg_dense_max=400"

# 7. entry-wise parity figures -> profiles/parity_entrywise_r02.txt
python scripts/entrywise_parity.py

# 8. multi-GPU (gpurun --gpus N; wrap in `timeout`): NCCL parity worker, C++ driver, strong scaling -> profiles/bench_r02_k7_{2,8}gpu.json
# python -m pytest tests/test_gpu_nccl.py -m gpu -q
# python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
#     bench.py --gpus N --steps 3 --warmup 2 --no-cpu

# 9. quadratic / cubic elements: kernel timings -> profiles/pk_bench_r02.json; device vs host probe of the
#    element functions (scripts/probes/pk_probe.cu, built with nvcc -fmad=false like pnp_p2.cu)
# python scripts/bench_pk.py --levels 3 > gpurun_out/pk_bench_r02.json
# nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false --expt-relaxed-constexpr scripts/probes/pk_probe.cu -o scripts/probes/pk_probe && scripts/probes/pk_probe
