for cfg in "OMP_NUM_THREADS=16" "OMP_NUM_THREADS=8" "OMP_NUM_THREADS=4" "OMP_NUM_THREADS=12" "OMP_NUM_THREADS=16 OMP_WAIT_POLICY=active" "OMP_NUM_THREADS=8 OMP_WAIT_POLICY=active OMP_PROC_BIND=close" "OMP_NUM_THREADS=16 OMP_PROC_BIND=spread" "GOMP_SPINCOUNT=1000000 OMP_NUM_THREADS=16"; do
  echo "== $cfg nproc=$(nproc)"
  env $cfg python bench.py --steps 60 --warmup 3 --no-cpu-baseline --sweep-level 0 --level-3d 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['split_ms_per_step'], {k:round(v,2) for k,v in d['host_stages_ms_per_step'].items()})"
done
