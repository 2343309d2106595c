# experiments on the timed step's multigrid: one bench line per variant (development aid; bench.py is the bench)
i=0
for v in "$@"; do
  i=$((i+1))
  python bench.py --no-cpu --steps 2 --warmup 3 $v > gpurun_out/var_$i.json 2> gpurun_out/var_$i.err
  python - "$v" gpurun_out/var_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-70s step %.4f s  its %d  defect %.3e  spmv_share %.2f launches %d e2e %.0f M" % (sys.argv[1], d["newton_step_s"], d["krylov_iterations"], d["defect_after"], d["spmv_share_of_step"], d["gpu_launches"], d["e2e"]["value"]/1e6))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
