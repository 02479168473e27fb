import random
def sim(STAGES, GROUPS, nchunks=40, seed=0):
    rnd = random.Random(seed)
    # barrier state: phase (int), pending
    full = [dict(phase=0, pend=1, tx=0) for _ in range(STAGES)]
    empty = [dict(phase=0, pend=1) for _ in range(STAGES)]  # count 1 (group as a unit)
    def complete(b, cnt):
        b['phase'] += 1; b['pend'] = cnt
    def wait_ok(b, parity):
        # true iff phase with given parity has completed: current phase parity != parity
        return (b['phase'] & 1) != parity
    prod = dict(it=0, state='start')
    cons = [dict(it=g, state='waitfull') for g in range(GROUPS)]
    inflight = []  # stages loading
    stage_owner = [None]*STAGES  # which 'it' data is in stage
    steps = 0
    while True:
        steps += 1
        if steps > 100000: return "deadlock/livelock"
        done = prod['it'] >= nchunks and all(c['it'] >= nchunks for c in cons) and not inflight
        if done: return "ok"
        actors = ['p'] + ['c%d' % g for g in range(GROUPS)] + (['land'] if inflight else [])
        a = rnd.choice(actors)
        if a == 'p' and prod['it'] < nchunks:
            it = prod['it']; s = it % STAGES; r = it // STAGES
            if r > 0 and not wait_ok(empty[s], (r - 1) & 1): continue
            # expect_tx arrive
            if full[s]['pend'] != 1 or full[s]['tx'] != 0: return "full[%d] bad state at issue it=%d: %s" % (s, it, full[s])
            full[s]['pend'] -= 1; full[s]['tx'] += 1
            if stage_owner[s] is not None: return "overwrite stage %d holding it=%s by it=%d" % (s, stage_owner[s], it)
            inflight.append((s, it)); prod['it'] += 1
        elif a == 'land':
            s, it = inflight.pop(0)
            full[s]['tx'] -= 1; stage_owner[s] = it
            if full[s]['pend'] == 0 and full[s]['tx'] == 0: complete(full[s], 1)
        elif a.startswith('c'):
            g = int(a[1:]); c = cons[g]
            if c['it'] >= nchunks: continue
            it = c['it']; s = it % STAGES; r = it // STAGES
            if not wait_ok(full[s], r & 1): continue
            if stage_owner[s] != it: return "group %d consumed stage %d expecting it=%d but holds %s" % (g, s, it, stage_owner[s])
            stage_owner[s] = None
            empty[s]['pend'] -= 1
            if empty[s]['pend'] < 0: return "empty underflow"
            if empty[s]['pend'] == 0: complete(empty[s], 1)
            c['it'] += GROUPS
for cfg in [(4,2),(3,3),(3,2),(4,3),(7,2),(5,2),(6,3),(2,2),(3,1)]:
    res = set(sim(*cfg, seed=s) for s in range(300))
    print(cfg, res)
