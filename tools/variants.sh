for v in "OMP_WAIT_POLICY=active" "OMP_WAIT_POLICY=active OMP_PROC_BIND=true" "OMP_NUM_THREADS=8 OMP_WAIT_POLICY=active OMP_PROC_BIND=true" "OMP_NUM_THREADS=12 OMP_WAIT_POLICY=active"; do
  echo "== $v" >> gpurun_out/variants.txt
  env $v python bench.py --no-cpu-baseline --sweep-level 0 --steps 60 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['split_ms_per_step'], d['host_stages_ms_per_step'], d['e2e']['ms_per_step'])" >> gpurun_out/variants.txt
done
