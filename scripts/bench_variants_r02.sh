for v in "amg_pre_steps=2 amg_post_steps=1" "amg_pre_steps=1 amg_post_steps=2" "amg_pre_steps=3 amg_post_steps=1" "amg_pre_steps=1 amg_post_steps=1" "amg_pre_steps=2 amg_post_steps=2 amg_omega=0.8" "amg_pre_steps=2 amg_post_steps=1 amg_omega=0.8"; do
  args=""; for kv in $v; do args="$args --solver-opt $kv"; done
  python bench.py --steps 3 --warmup 2 --no-cpu $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],1), d['krylov_iterations'], d['defect_after'])"
done
