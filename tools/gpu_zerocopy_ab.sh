#!/bin/bash
# A/B of the zero-copy host step (EARL_TT_HOST_ZEROCOPY): e2e through PersistentStateWrapper.step with pinned host buffers
O=gpurun_out/zc2
mkdir -p $O
timeout 200 python -m pytest tests/test_tabletop_gpu.py -m gpu -x -q -k "not benchmarked_horizon and not config1" 2>&1 | tail -2
for n in 1048576; do
  for z in 0 1 0 1; do
    EARL_TT_HOST_ZEROCOPY=$z timeout 120 python bench.py --steps 20 --warmup 5 --num-envs $n --no-door --no-hbm-check --no-cpu-baseline 2>$O/err_${n}_$z.txt | tail -1 > $O/b_${n}_$z.json
    python -c "
import json,sys
d=json.load(open('$O/b_${n}_$z.json')); e=d['e2e']
print('n $n zerocopy $z e2e %.4g ceiling %.4g frac %.3f value %.4g' % (e['value'], e['pcie_ceiling']['value'], e['frac_of_pcie_ceiling'], d['value']))"
  done
done
EARL_TT_HOST_ZEROCOPY=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-door --no-cpu-baseline 2>$O/err_full.txt | tail -1 > $O/b_full.json
python -c "
import json
d=json.load(open('$O/b_full.json')); print('full: value %.4g e2e %.4g all_hbm %.4f' % (d['value'], d['e2e']['value'], d['roofline']['frac_all_hbm']))
for p in d['sweep']: print('  ', p['point'][:50], '%.4g' % p['value'], '%.3f' % p['frac'])"
