python -m pytest tests -m gpu -q -x -k "hmc or hamilton or verlet or force or neighbour" 2>&1 | tail -3
for skin in 0.7 0.5 0.4; do
  echo -n "skin=$skin "; QMC_HMC_SKIN=$skin python bench_configs.py c5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['trajectories_per_s'], d['acceptance'])"
done
