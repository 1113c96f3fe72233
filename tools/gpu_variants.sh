# A/B of step-kernel variants (EARL_TT_VARIANT / EARL_TT_TILE / EARL_TT_PDL)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
EARL_TT_PDL=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
EARL_TT_PDL=1 EARL_TT_VARIANT=3 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for n in 262144 1048576 2097152 4194304 8388608; do
  for cfg in "0 0" "0 1" "6 0" "6 1" "3 0" "3 1"; do
    set -- $cfg
    EARL_TT_VARIANT=$1 EARL_TT_PDL=$2 python bench.py --num-envs $n --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('variant $1 pdl $2 n', d['config']['envs_per_gpu'], '%.3e'%d['value'], 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'])"
  done
done
