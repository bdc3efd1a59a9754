"""A discrete model of the CTA-pair protocol of tc::k_eval_tc (facedeform_b200/csrc/fd_eval_tc.cu, the FP16 hi/lo tensor-core
evaluation): two CTAs of a cluster take the same column block of two neighbouring vertex tiles in lockstep; each loads HALF of
every weight tile and multicasts it into both CTAs' slot, each CTA's full_b barrier expects the whole tile, and a slot is free
only when the MMAs of BOTH CTAs have read it (empty barrier of count 2, tcgen05.commit multicast to both).

Roles per CTA as coroutines over mbarriers with the kernel's counts, ring sizes (4 stages, 16 centre tiles), slot indices and
parities; the tensor cores and the bulk loads complete asynchronously; random schedules.  Checked: progress (no deadlock), every
MMA reads Phi and BOTH weight halves of its own stage, no half lands in a slot an MMA in flight reads, no Phi slot is overwritten
under an MMA in flight, the epilogue drains its own unit, accumulators are released before they are rewritten.  A half may land
before the receiving CTA has posted its expect_tx: the transaction count goes negative for a while, as the hardware allows.
Usage: python tests/tools/tc_pair_protocol_model.py [seed]"""
import random

STAGES = 4
CDEPTH = 16
PRODUCER_GROUP_WARPS = 8
EPILOGUE_WARPS = 8


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self, tx=0):
        assert self.pending > 0, "more arrivals than the barrier counts"
        self.tx += tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, n):
        self.tx -= n  # may run ahead of the expect_tx of this phase
        self._maybe_complete()

    def passed(self, parity):
        return self.phase != parity


class CTA:
    def __init__(self, rank, pair):
        self.rank = rank
        self.full_a = [MBar(PRODUCER_GROUP_WARPS) for _ in range(STAGES)]
        self.full_b = [MBar(1) for _ in range(STAGES)]
        self.empty = [MBar(2 if pair else 1) for _ in range(STAGES)]
        self.tmem_full = [MBar(1), MBar(1)]
        self.tmem_empty = [MBar(EPILOGUE_WARPS), MBar(EPILOGUE_WARPS)]
        self.cfull = [MBar(1) for _ in range(CDEPTH)]
        self.cempty = [MBar(PRODUCER_GROUP_WARPS) for _ in range(CDEPTH)]
        self.a_slot = [None] * STAGES
        self.b_half = [[None, None] for _ in range(STAGES)]  # stage counter of each half of the weight tile
        self.c_slot = [None] * CDEPTH
        self.a_readers = [0] * STAGES
        self.b_readers = [0] * STAGES
        self.acc_unit = [None, None]
        self.acc_busy = [False, False]
        self.inflight = []
        self.done = {"epi": [], "mma": []}


class PairModel:
    def __init__(self, n_units, nk, pair, rng):
        self.rng, self.nk, self.pair = rng, nk, pair
        self.units = list(range(n_units))
        self.ctas = [CTA(r, pair) for r in range(2 if pair else 1)]
        self.loads = []

    def tma(self, c):
        it = ic = kc = 0
        total = len(self.units) * self.nk
        for u in self.units:
            for kb in range(self.nk):
                while ic < total and ic < it + CDEPTH:
                    s = ic % CDEPTH
                    if not c.cempty[s].passed(((ic // CDEPTH) & 1) ^ 1):
                        break
                    c.cfull[s].arrive(tx=1)
                    self.loads.append(("c", c, s, kc))
                    ic += 1
                    kc = (kc + 1) % self.nk
                s = it % STAGES
                yield lambda s=s, p=((it // STAGES) & 1) ^ 1: c.empty[s].passed(p)
                c.full_b[s].arrive(tx=2)  # the whole tile: both halves land in this CTA
                if self.pair:
                    self.loads.append(("b", self.ctas, s, c.rank, it))  # this CTA's half, delivered to both
                else:
                    self.loads.append(("b", [c], s, 0, it))
                    self.loads.append(("b", [c], s, 1, it))
                it += 1
                yield None

    def mma(self, c):
        it = 0
        for unit_iter, u in enumerate(self.units):
            ab = unit_iter & 1
            yield lambda ab=ab, p=((unit_iter >> 1) & 1) ^ 1: c.tmem_empty[ab].passed(p)
            assert not c.acc_busy[ab], "MMA into accumulators the epilogue has not released"
            for kb in range(self.nk):
                s = it % STAGES
                yield lambda s=s, p=(it // STAGES) & 1: c.full_a[s].passed(p)
                yield lambda s=s, p=(it // STAGES) & 1: c.full_b[s].passed(p)
                c.a_readers[s] += 1
                c.b_readers[s] += 1
                c.inflight.append(("mma", s, it, ab, u))
                c.inflight.append(("free", s))  # umma_commit_mc(bar_empty, 3) / umma_commit
                if kb == self.nk - 1:
                    c.inflight.append(("full", ab, u))
                it += 1
                yield None
            c.done["mma"].append(u)

    def tensor_core(self, c):
        while True:
            yield lambda: bool(c.inflight) or self.finished_issuing
            if not c.inflight:
                return
            op = c.inflight.pop(0)
            if op[0] == "mma":
                _, s, it, ab, u = op
                assert c.a_slot[s] == it, f"CTA {c.rank}: MMA of stage {it} read Phi slot {s} holding {c.a_slot[s]}"
                assert c.b_half[s] == [it, it], f"CTA {c.rank}: MMA of stage {it} read weight halves {c.b_half[s]}"
                c.a_readers[s] -= 1
                c.b_readers[s] -= 1
                c.acc_unit[ab] = u
            elif op[0] == "free":
                for d in self.ctas:  # the arrive lands on the barrier of every CTA of the pair
                    d.empty[op[1]].arrive()
            else:
                _, ab, u = op
                assert c.acc_unit[ab] == u
                c.acc_busy[ab] = True
                c.tmem_full[ab].arrive()

    def loader(self):
        while True:
            yield lambda: bool(self.loads) or self.finished_issuing
            if not self.loads:
                return
            op = self.loads.pop(self.rng.randrange(len(self.loads)))
            if op[0] == "c":
                _, c, s, kc = op
                c.c_slot[s] = kc
                c.cfull[s].complete_tx(1)
            else:
                _, dests, s, half, it = op
                for d in self.rng.sample(dests, len(dests)):  # the copies of a multicast land one by one
                    assert d.b_readers[s] == 0, f"a weight half landed in slot {s} of CTA {d.rank} under an MMA in flight"
                    d.b_half[s][half] = it
                    d.full_b[s].complete_tx(1)

    def producer(self, c, grp):
        it = 0
        for u in self.units:
            it0 = it
            it += self.nk
            for kb in range((grp ^ it0) & 1, self.nk, 2):
                itk = it0 + kb
                s, cs = itk % STAGES, itk % CDEPTH
                yield lambda cs=cs, p=(itk // CDEPTH) & 1: c.cfull[cs].passed(p)
                yield lambda s=s, p=((itk // STAGES) & 1) ^ 1: c.empty[s].passed(p)
                assert c.c_slot[cs] == kb
                assert c.a_readers[s] == 0, "Phi slot overwritten while an MMA in flight reads it"
                c.a_slot[s] = itk
                yield None
                c.full_a[s].arrive()
                c.cempty[cs].arrive()

    def epilogue(self, c, warp):
        for unit_iter, u in enumerate(self.units):
            ab = unit_iter & 1
            yield lambda ab=ab, p=(unit_iter >> 1) & 1: c.tmem_full[ab].passed(p)
            assert c.acc_unit[ab] == u and c.acc_busy[ab]
            yield None
            if c.tmem_empty[ab].pending == 1:
                c.acc_busy[ab] = False
            c.tmem_empty[ab].arrive()
            if warp == 0:
                c.done["epi"].append(u)

    def run(self):
        self.finished_issuing = False
        issuing = {}
        for c in self.ctas:
            issuing[f"{c.rank}.tma"] = self.tma(c)
            issuing[f"{c.rank}.mma"] = self.mma(c)
            for g in range(2):
                for w in range(PRODUCER_GROUP_WARPS):
                    issuing[f"{c.rank}.p{g}.{w}"] = self.producer(c, g)
            for w in range(EPILOGUE_WARPS):
                issuing[f"{c.rank}.e{w}"] = self.epilogue(c, w)
        roles = dict(issuing)
        for c in self.ctas:
            roles[f"{c.rank}.tc"] = self.tensor_core(c)
        roles["ld"] = self.loader()
        waiting = {}
        for name, gen in list(roles.items()):
            try:
                waiting[name] = next(gen)
            except StopIteration:
                del roles[name]
        steps = 0
        while roles:
            self.finished_issuing = not any(n in roles for n in issuing)
            ready = [n for n in roles if waiting[n] is None or waiting[n]()]
            if not ready:
                raise RuntimeError(f"deadlock after {steps} steps; waiting: {sorted(roles)[:12]}")
            name = self.rng.choice(ready)
            try:
                waiting[name] = roles[name].send(None)
            except StopIteration:
                del roles[name]
            steps += 1
        for c in self.ctas:
            assert c.done["epi"] == self.units and c.done["mma"] == self.units and not c.inflight
        assert not self.loads
        return steps


def check(n_units, nk, pair, seed):
    return PairModel(n_units, nk, pair, random.Random(seed)).run()


if __name__ == "__main__":
    import sys
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    for pair in (False, True):
        for n_units in (1, 2, 5):
            for nk in (1, 2, 3, 9, 20):
                print(f"pair={int(pair)} units={n_units} stages={nk}: ok ({check(n_units, nk, pair, seed)} scheduler steps)")
