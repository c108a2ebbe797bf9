"""Randomised model of the mbarrier protocol of conv3x3_dx_kernel (csrc/conv3x3_dx.cu): one TMA producer, kMmaWarps
UMMA-issuing warps on alternate units, the epilogue groups, and asynchronous completions (TMA transactions, tcgen05.commit
arrivals) delivered after random delays.  mbarrier waits are parity based, exactly as the hardware's: wait(P) succeeds
iff the barrier's current phase has parity != P -- so a waiter that skips a phase, or is lapped, is caught here as a
deadlock or as a read of a stage / buffer in the wrong fill (`check` asserts).  Run:  python tools/sim_dx_protocol.py"""
import itertools
import random
import sys


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def ready(self, parity):
        return (self.phase & 1) != parity


def simulate(seed, units, m_tiles, a_stages, kchunks, nsub, nbuf, ngroups, res, mma_warps, shadow=True):
    rng = random.Random(seed)
    a_full = [Bar(1) for _ in range(a_stages)]
    a_empty = [Bar(1) for _ in range(a_stages)]
    b_full, b_empty = Bar(1), [Bar(1) for _ in range(mma_warps)]
    acc_full = [Bar(1) for _ in range(nbuf)]
    acc_empty = [Bar(1) for _ in range(nbuf)]          # one arrival per epilogue group (models group_warps arrivals)
    res_stages = ngroups
    res_full = [Bar(1) for _ in range(res_stages)]
    res_empty = [Bar(1) for _ in range(res_stages)]
    events = []                                         # (time, seq, fn)
    seq = itertools.count()
    now = [0.0]
    # ground truth of what each stage / buffer currently holds, to detect reads of the wrong fill
    a_content = [None] * a_stages                       # (unit, kc) landed in the stage
    acc_content = [None] * nbuf                         # item whose MMAs completed into the buffer
    errors = []

    def later(lo, hi, fn):
        events.append((now[0] + rng.uniform(lo, hi), next(seq), fn))

    ring = a_stages // mma_warps                      # activation stages owned by each issuing warp

    def producer():
        b_par = rs = rph = 0
        rpos, rph_a = [0] * mma_warps, [0] * mma_warps  # ring position / phase per issuing warp
        new_panel, first, m = True, True, 0
        for u in range(units):
            w_ = u % mma_warps
            if new_panel:
                if not first:
                    for w in range(mma_warps):
                        yield (b_empty[w], b_par)
                    b_par ^= 1
                later(0.5, 3.0, b_full.arrive)
            first = False
            for kc in range(kchunks):
                st = w_ * ring + rpos[w_]
                yield (a_empty[st], rph_a[w_] ^ 1)

                def land(st=st, u=u, kc=kc):
                    a_content[st] = (u, kc)
                    a_full[st].arrive()
                later(0.5, 4.0, land)
                rpos[w_] += 1
                if rpos[w_] == ring:
                    rpos[w_], rph_a[w_] = 0, rph_a[w_] ^ 1
            if res:
                for s in range(nsub):
                    yield (res_empty[rs], rph ^ 1)
                    later(0.5, 4.0, res_full[rs].arrive)
                    rs += 1
                    if rs == res_stages:
                        rs, rph = 0, rph ^ 1
            m += 1
            new_panel = m == m_tiles
            if new_panel:
                m = 0

    def mma(me):
        stage = phase = buf = aph = item = panel_idx = 0
        m, turn = 0, 0
        have_panel = issued = False
        last_commit = [0.0]

        def commit(fn):                                # commits of one thread complete in order
            t = max(last_commit[0], now[0]) + rng.uniform(0.2, 2.0)
            last_commit[0] = t
            events.append((t, next(seq), fn))
        for u in range(units):
            panel_ends = (u + 1 == units) or (m + 1 == m_tiles)
            if not have_panel:
                yield (b_full, panel_idx & 1)
                have_panel = True
            if turn == me:
                issued = True
                bf, bph = buf, aph
                for s in range(nsub):
                    st, ph = stage, phase
                    for kc in range(kchunks):
                        if s == 0:
                            yield (a_full[me * ring + st], ph)
                        if kc == 0:
                            yield (acc_empty[bf], bph ^ 1)
                        if a_content[me * ring + st] != (u, kc):
                            errors.append(f"warp {me} unit {u} kc {kc}: stage {me * ring + st} holds {a_content[me * ring + st]}")
                        if s == nsub - 1:
                            commit(a_empty[me * ring + st].arrive)
                        if kc == kchunks - 1:
                            def done(bf=bf, it=item + s):
                                acc_content[bf] = it
                                acc_full[bf].arrive()
                            commit(done)
                        st += 1
                        if st == ring:
                            st, ph = 0, ph ^ 1
                        yield None                      # issue time
                    bf += 1
                    if bf == nbuf:
                        bf, bph = 0, bph ^ 1
            if turn == me:                                # the ring of this warp advances with its own units only
                for kc in range(kchunks):
                    stage += 1
                    if stage == ring:
                        stage, phase = 0, phase ^ 1
            for s in range(nsub):
                buf += 1
                if buf == nbuf:
                    buf, aph = 0, aph ^ 1
            item += nsub
            turn = (turn + 1) % mma_warps
            if panel_ends:
                if issued:
                    commit(b_empty[me].arrive)
                else:
                    b_empty[me].arrive()
                have_panel = issued = False
                panel_idx += 1
            m += 1
            if m == m_tiles:
                m = 0

    def epilogue(group):
        buf = aph = rs = rph = turn = item = 0
        for u in range(units):
            for s in range(nsub):
                my = (buf, aph, rs, rph)
                mine = turn == group
                buf += 1
                if buf == nbuf:
                    buf, aph = 0, aph ^ 1
                rs += 1
                if rs == res_stages:
                    rs, rph = 0, rph ^ 1
                turn = (turn + 1) % ngroups
                it = item
                item += 1
                if not mine:
                    continue
                yield (acc_full[my[0]], my[1])
                if acc_content[my[0]] != it:
                    errors.append(f"group {group} item {it}: buffer {my[0]} holds item {acc_content[my[0]]}")
                if res:
                    yield (res_full[my[2]], my[3])
                for _ in range(rng.randint(1, 6)):
                    yield None
                acc_empty[my[0]].arrive()
                if res:
                    res_empty[my[2]].arrive()

    procs = {"producer": producer()}
    for w in range(mma_warps):
        procs[f"mma{w}"] = mma(w)
    for g in range(ngroups):
        procs[f"epi{g}"] = epilogue(g)
    waiting = {k: None for k in procs}                  # None = runnable
    steps = 0
    while procs:
        steps += 1
        if steps > 2_000_000:
            return "livelock", errors
        runnable = [k for k in procs if waiting[k] is None or waiting[k][0].ready(waiting[k][1])]
        if not runnable:
            if not events:
                return "deadlock: " + ", ".join(
                    f"{k} waits {'phase parity ' + str(waiting[k][1])}" for k in procs), errors
            events.sort()
            t, _, fn = events.pop(0)
            now[0] = t
            fn()
            continue
        # deliver due events with some probability, else advance a random runnable process
        if events and rng.random() < 0.3:
            events.sort()
            t, _, fn = events.pop(0)
            now[0] = max(now[0], t)
            fn()
            continue
        k = rng.choice(runnable)
        waiting[k] = None
        try:
            w = next(procs[k])
            now[0] += rng.uniform(0.0, 0.3)
            waiting[k] = w
        except StopIteration:
            del procs[k]
            del waiting[k]
    while events:
        events.sort()
        _, _, fn = events.pop(0)
        fn()
    return "ok", errors


def main():
    configs = []
    for a_stages in (2, 3, 4, 5, 6):        # with two issuing warps each owns a_stages // 2 stages (>= 1)
        for kchunks in (1, 2, 3):
            for nsub, nbuf, ngroups in ((1, 4, 4), (2, 4, 4), (1, 2, 2)):
                for res in (False, True):
                    configs.append(dict(a_stages=a_stages, kchunks=kchunks, nsub=nsub, nbuf=nbuf, ngroups=ngroups, res=res))
    shadow = "--no-shadow" not in sys.argv
    bad = 0
    for cfg in configs:
        for mma_warps in (1, 2):
            for units, m_tiles in ((20, 7), (9, 3), (4, 12), (13, 1)):
                for seed in range(40):
                    r, errs = simulate(seed, units, m_tiles, mma_warps=mma_warps, shadow=shadow, **cfg)
                    if r != "ok" or errs:
                        bad += 1
                        if bad <= 12:
                            print("FAIL", cfg, "mma_warps", mma_warps, "units", units, "m_tiles", m_tiles, "seed", seed, r, errs[:2])
    print("configs x warps x shapes x seeds:", len(configs) * 2 * 4 * 40, "failures:", bad)


if __name__ == "__main__":
    main()
