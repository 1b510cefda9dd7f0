"""Randomised model of the mbarrier protocol of the multi-issuer convolution kernels (igemm_kmajor_kernel<.., MI = true>; the halo and
wgrad kernels follow the same scheme): one TMA producer filling an S-deep stage ring, I MMA-issuing threads, an epilogue draining two
accumulators, and the barriers full[S], empty[S], tfull[2] (count I), tempty[2], zinit[2].
mbarrier semantics as in PTX: `try_wait.parity P` succeeds iff the phase of parity P has completed, which is only meaningful while the
waiter is at most one phase behind; tcgen05.commit completions arrive asynchronously, in order per issuing thread; TMA loads complete in
issue order (tma_in_order) or in any order (what the hardware does).

The model checks, under random schedules: no deadlock, an issuer never passes `full` onto a slot that does not hold its (tile, stage)
(phase aliasing), nobody accumulates into an accumulator another tile owns, the epilogue reads the tile it expects.

Two ownership schemes:
 * by stage index (issuer x takes the stages it % I == x of every tile; the first version of the kernels): sound only if TMA loads
   land in order.  With out-of-order completion it fails exactly where the hardware did (profiles/r01_issuers_status.txt): 3-stage rings
   with 2 issuers, and 4-stage rings whose work items have an odd stage count (the stem's wgrad: 339).  An issuer that skips a pass of a
   slot waits for the NEXT pass with the same parity; if the skipped pass has not landed yet the barrier is one phase behind and the
   parity test lets the issuer through onto stale data, after which its `empty` arrival corrupts the producer's accounting.
 * by ring slot (issuer x owns the slots s % I == x, S % I == 0; every issuer waits `tempty`, the owner of a tile's first stage
   overwrites the accumulator and commits `zinit`; what igemm.cu does now): every waiter sees every phase of its barriers; no failure
   in any configuration tried, in-order or not.
tests/test_host_logic.py runs both."""
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


def run(S, iters_list, I, seed, by_slot=False, tma_in_order=True):
    """-> ('OK' | 'HANG' | 'ERR', errors).  by_slot: issuer x owns the ring SLOTS s % I == x (needs S % I == 0) instead of the
    stages it % I == x of every tile.  tma_in_order = False: TMA loads complete in any order (they do on the hardware)."""
    rnd = random.Random(seed)
    full, empty = [Bar(1) for _ in range(S)], [Bar(1) for _ in range(S)]
    tfull, tempty, zinit = [Bar(I) for _ in range(2)], [Bar(1) for _ in range(2)], [Bar(1) for _ in range(2)]
    assert not by_slot or S % I == 0
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
                if tma_in_order:
                    later(st, land)
                else:
                    pending.append((now[0] + rnd.randint(1, 40), land))
                stage += 1
                if stage == S:
                    stage, phase = 0, phase ^ 1
                yield

    def issuer(x):
        st, stage, phase, acc, accphase = {}, 0, 0, 0, 0
        for tile, iters in enumerate(iters_list):
            if by_slot:  # every issuer waits for the drained accumulator; the owner of the tile's first stage overwrites it, the others wait for that
                while not tempty[acc].test(accphase ^ 1):
                    yield
                if stage % I != x:
                    while not zinit[acc].test(accphase):
                        yield
            else:
                while not (tempty[acc].test(accphase ^ 1) if x == 0 else zinit[acc].test(accphase)):
                    yield
            for it in range(iters):
                if (stage % I == x) if by_slot else (it % I == x):
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


CONFIGS = [(4, 4, 2), (4, 9, 2), (4, 339, 2), (4, 7, 2), (8, 9, 2), (8, 1, 2), (8, 2, 4), (4, 4, 4), (8, 36, 4), (2, 8, 2), (6, 5, 2)]


def sweep(seeds=40, tiles=8, by_slot=True, tma_in_order=False, configs=CONFIGS):
    """-> {(S, stages per tile, I): number of failing seeds}"""
    out = {}
    for S, iters, I in configs:
        out[(S, iters, I)] = sum(run(S, [iters] * tiles, I, seed, by_slot, tma_in_order)[0] != "OK" for seed in range(seeds))
    if by_slot:
        out["mixed tile lengths, S=4, I=2"] = sum(run(4, [8, 4, 4, 2, 1, 3] * 2, 2, seed, True, tma_in_order)[0] != "OK" for seed in range(seeds))
    return out




def run_halo(AS, BS, kchunks, ntaps, tiles, I, seed, resident=False):
    """The same model for igemm_halo_kernel: an A ring of AS patches (one per K chunk, read by ALL issuers: afull waited by everyone, aempty
    count I) and a B ring of BS weight tiles (one per (K chunk, tap), owned by slot: bs % I == x; resident weights: taps t % I == x, no B ring)."""
    rnd = random.Random(seed)
    afull, aempty = [Bar(1) for _ in range(AS)], [Bar(I) for _ in range(AS)]
    bfull, bempty = [Bar(1) for _ in range(max(BS, 1))], [Bar(1) for _ in range(max(BS, 1))]
    tfull, tempty, zinit = [Bar(I) for _ in range(2)], [Bar(1) for _ in range(2)], [Bar(1) for _ in range(2)]
    aslot, bslot, acc_owner = [None] * AS, [None] * max(BS, 1), [None, None]
    pending, now, errors = [], [0], []
    assert resident or BS % I == 0

    def later(state, fn):
        t = max(state.get("last", 0), now[0]) + rnd.randint(1, 6)
        state["last"] = t
        pending.append((t, fn))

    def producer():
        a, aph, b, bph = 0, 0, 0, 0
        for tile in range(tiles):
            for kc in range(kchunks):
                while not aempty[a].test(aph ^ 1):
                    yield

                def landa(a=a, tile=tile, kc=kc):
                    aslot[a] = (tile, kc)
                    afull[a].arrive()
                pending.append((now[0] + rnd.randint(1, 40), landa))
                a += 1
                if a == AS:
                    a, aph = 0, aph ^ 1
                if not resident:
                    for t in range(ntaps):
                        while not bempty[b].test(bph ^ 1):
                            yield

                        def landb(b=b, tile=tile, kc=kc, t=t):
                            bslot[b] = (tile, kc, t)
                            bfull[b].arrive()
                        pending.append((now[0] + rnd.randint(1, 40), landb))
                        b += 1
                        if b == BS:
                            b, bph = 0, bph ^ 1
                        yield
                yield

    def issuer(x):
        st, a, aph, b, bph, acc, accphase = {}, 0, 0, 0, 0, 0, 0
        for tile in range(tiles):
            while not tempty[acc].test(accphase ^ 1):
                yield
            if (0 if resident else b % I) != x:
                while not zinit[acc].test(accphase):
                    yield
            for kc in range(kchunks):
                while not afull[a].test(aph):
                    yield
                if aslot[a] != (tile, kc):
                    errors.append(("stale patch", x, tile, kc, aslot[a]))
                for t in range(ntaps):
                    if (t % I if resident else b % I) == x:
                        if not resident:
                            while not bfull[b].test(bph):
                                yield
                            if bslot[b] != (tile, kc, t):
                                errors.append(("stale weights", x, tile, kc, t, bslot[b]))
                        if kc == 0 and t == 0:
                            acc_owner[acc] = tile
                            if I > 1:
                                later(st, lambda c=acc: zinit[c].arrive())
                        elif acc_owner[acc] != tile:
                            errors.append(("accumulator hazard", x, tile, kc, t, acc_owner[acc]))
                        if not resident:
                            later(st, lambda s_=b: bempty[s_].arrive())
                        yield
                    if not resident:
                        b += 1
                        if b == BS:
                            b, bph = 0, bph ^ 1
                later(st, lambda s_=a: aempty[s_].arrive())
                a += 1
                if a == AS:
                    a, aph = 0, aph ^ 1
            later(st, lambda c=acc: tfull[c].arrive())
            acc ^= 1
            if acc == 0:
                accphase ^= 1
            yield

    def epilogue():
        acc, accphase = 0, 0
        for tile in range(tiles):
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
        if steps > 400000:
            return "HANG", errors
    return "OK", errors


def sweep_halo(seeds=20, tiles=5):
    out = {}
    for AS, BS, kc, I, res in [(4, 0, 1, 2, True), (4, 0, 1, 4, True), (3, 12, 2, 2, False), (3, 6, 4, 2, False), (3, 12, 2, 4, False), (3, 8, 4, 4, False), (2, 4, 1, 2, False)]:
        out[(AS, BS, kc, I, "resident" if res else "streamed")] = sum(run_halo(AS, BS, kc, 9, tiles, I, seed, res)[0] != "OK" for seed in range(seeds))
    return out


if __name__ == "__main__":
    print("ownership by ring slot, TMA out of order:")
    for k, v in sweep(seeds=100).items():
        print("  ", k, "failing seeds:", v)
    print("ownership by stage index (first version), TMA out of order:")
    for k, v in sweep(seeds=100, by_slot=False, configs=[(3, 4, 2), (3, 8, 2), (4, 339, 2), (4, 7, 2), (4, 8, 2), (8, 9, 2)]).items():
        print("  ", k, "failing seeds:", v)
    print("ownership by stage index, TMA in order:")
    for k, v in sweep(seeds=100, by_slot=False, tma_in_order=True, configs=[(3, 4, 2), (3, 8, 2), (4, 339, 2), (4, 7, 2), (3, 4, 4)]).items():
        print("  ", k, "failing seeds:", v)
    print("igemm_halo_kernel (A ring read by all issuers, weight ring owned by slot / resident weights split by tap), TMA out of order:")
    for k, v in sweep_halo(seeds=100).items():
        print("  ", k, "failing seeds:", v)
