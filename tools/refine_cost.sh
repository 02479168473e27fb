# cost of the float64 refinement of ill-conditioned pixels (HIPR_LNE2D_REFINE=0 disables it), by number of streams
for r in 1 0; do for s in ${STREAMS:-2 1}; do HIPR_LNE2D_REFINE=$r python bench.py --steps 100 --warmup 5 --streams $s --pool ${POOL:-2} --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('refine $r streams $s', d['ms_per_step'], d['value'], 'K1 alone', d['roofline']['ms_per_launch'])"; done; done
