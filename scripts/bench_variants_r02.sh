# usage: bash scripts/bench_variants_r02.sh "<solver-opt list>" ...   -- one bench run per argument (N = 1), prints ms/step, iterations, defect
for v in "$@"; do
  args=""; for kv in $v; do args="$args --solver-opt $kv"; done
  python bench.py --steps 3 --warmup 2 --no-cpu $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],1), d['krylov_iterations'], d['defect_after'], d['gpu_launches'], round(d['roofline']['step']['frac'],3))"
done
