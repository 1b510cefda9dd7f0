"""Randomised model of the mbarrier protocol of the multi-issuer convolution kernels (igemm_kmajor_kernel<.., MI = true>; the halo and
wgrad kernels follow the same scheme): one TMA producer filling an S-deep stage ring, I MMA-issuing threads that own the stages
`it % I` of every tile, an epilogue draining two accumulators, and the barriers full[S], empty[S], tfull[2] (count I), tempty[2], zinit[2].
mbarrier semantics as in PTX: `try_wait.parity P` succeeds iff the phase of parity P has completed, which is only meaningful while the
waiter is at most one phase behind; tcgen05.commit / TMA completions arrive asynchronously, in order per issuing thread.

The model checks, under random schedules: no deadlock, an issuer never passes `full` onto a slot that does not hold its (tile, stage)
(phase aliasing), nobody accumulates into an accumulator another tile owns, the epilogue reads the tile it expects.

Result (tests/test_host_logic.py runs it): sound for I <= S; with more issuers than ring stages (I = 4, S = 3) an issuer waits for a slot
whose PREVIOUS pass has not been filled yet and the parity test lets it through -- igemm.cu therefore clamps issuers to the stage count.
The hang / fault of the batch-256 step with I = 2 (profiles/r01_issuers_status.txt) is NOT reproduced by this model, i.e. it is not a
flaw of the barrier protocol as written but of an assumption about the hardware (candidates: what tcgen05.commit tracks when two threads
of a CTA have MMAs in flight; MMAs of different shapes from two threads interleaving at a tile boundary)."""
import random


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):
        return (self.phase & 1) != parity


def run(S, iters_list, I, seed):
    """-> ('OK' | 'HANG' | 'ERR', errors)"""
    rnd = random.Random(seed)
    full, empty = [Bar(1) for _ in range(S)], [Bar(1) for _ in range(S)]
    tfull, tempty, zinit = [Bar(I) for _ in range(2)], [Bar(1) for _ in range(2)], [Bar(1) for _ in range(2)]
    slot, acc_owner = [None] * S, [None, None]
    pending, now, errors = [], [0], []

    def later(state, fn):  # asynchronous completion, in order per issuing thread
        t = max(state.get("last", 0), now[0]) + rnd.randint(1, 6)
        state["last"] = t
        pending.append((t, fn))

    def producer():
        st, stage, phase = {}, 0, 0
        for tile, iters in enumerate(iters_list):
            for it in range(iters):
                while not empty[stage].test(phase ^ 1):
                    yield

                def land(stage=stage, tile=tile, it=it):
                    slot[stage] = (tile, it)
                    full[stage].arrive()
                later(st, land)
                stage += 1
                if stage == S:
                    stage, phase = 0, phase ^ 1
                yield

    def issuer(x):
        st, stage, phase, acc, accphase = {}, 0, 0, 0, 0
        for tile, iters in enumerate(iters_list):
            while not (tempty[acc].test(accphase ^ 1) if x == 0 else zinit[acc].test(accphase)):
                yield
            for it in range(iters):
                if it % I == x:
                    while not full[stage].test(phase):
                        yield
                    if slot[stage] != (tile, it):
                        errors.append(("stale slot", x, tile, it, slot[stage]))
                    if it == 0:
                        acc_owner[acc] = tile
                        if I > 1:
                            later(st, lambda a=acc: zinit[a].arrive())
                    elif acc_owner[acc] != tile:
                        errors.append(("accumulator hazard", x, tile, it, acc_owner[acc]))
                    later(st, lambda s_=stage: empty[s_].arrive())
                    yield
                stage += 1
                if stage == S:
                    stage, phase = 0, phase ^ 1
            later(st, lambda a=acc: tfull[a].arrive())
            acc ^= 1
            if acc == 0:
                accphase ^= 1
            yield

    def epilogue():
        acc, accphase = 0, 0
        for tile, _ in enumerate(iters_list):
            while not tfull[acc].test(accphase):
                yield
            if acc_owner[acc] != tile:
                errors.append(("epilogue reads another tile", tile, acc_owner[acc]))
            for _ in range(rnd.randint(0, 8)):
                yield
            tempty[acc].arrive()
            acc ^= 1
            if acc == 0:
                accphase ^= 1
            yield

    threads = [producer()] + [issuer(x) for x in range(I)] + [epilogue()]
    alive, steps = [True] * len(threads), 0
    while any(alive):
        now[0] += 1
        for p in sorted([p for p in pending if p[0] <= now[0]], key=lambda q: q[0]):
            pending.remove(p)
            p[1]()
        i = rnd.randrange(len(threads))
        if alive[i]:
            try:
                next(threads[i])
            except StopIteration:
                alive[i] = False
        steps += 1
        if errors:
            return "ERR", errors
        if steps > 200000:
            return "HANG", errors
    return "OK", errors


def sweep(seeds=40, tiles=10):
    """-> {(S, iters, I): number of failing seeds}"""
    out = {}
    for cfg in [(3, 4, 2), (4, 9, 2), (3, 4, 4), (4, 4, 4), (8, 1, 2), (8, 2, 4), (4, 7, 2), (3, 5, 2), (4, 36, 4)]:
        S, iters, I = cfg
        out[cfg] = sum(run(S, [iters] * tiles, I, seed)[0] != "OK" for seed in range(seeds))
    out["mixed groups, S=4, I=2"] = sum(run(4, [8, 4, 4, 2] * 3, 2, seed)[0] != "OK" for seed in range(seeds))
    return out


if __name__ == "__main__":
    for k, v in sweep(seeds=200).items():
        print(k, "failing seeds:", v)
