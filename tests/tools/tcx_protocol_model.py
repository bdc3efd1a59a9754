"""A discrete model of the synchronisation protocol of tcx::k_eval_tcx (facedeform_b200/csrc/fd_eval_tcx.cu):
the roles of one CTA -- TMA warp, MMA warp + the tensor core's asynchronous completion, two producer groups, epilogue -- as
coroutines over mbarriers with the kernel's counts, ring sizes, slot indices and wait parities, run under random schedules.

What it checks (compute-sanitizer is closed on the GPU pool, so this is the independent look at the protocol):
  * progress: no schedule deadlocks, every role finishes all its units;
  * data: every MMA reads the Phi slot and the weight slot of ITS stage (tags), no slot is overwritten (by the producers' stores
    or by a landing bulk load) while an MMA that reads it is still in flight, the epilogue drains the accumulators of ITS unit
    and the MMAs never write accumulators the epilogue has not released.
The index arithmetic is transcribed from the kernel: a change there has to be made here too (tests/test_tcx_protocol.py runs it
over the shapes the kernel meets: one to seven column blocks, one to many stages, more CTAs than units...).
Usage: python tests/tools/tcx_protocol_model.py [seed]"""
import random

CDEPTH = 8
PRODUCER_GROUP_WARPS = 8  # arrivals on full_a / cempty per stage
EPILOGUE_WARPS = 4


class MBar:
    """mbarrier: `count` arrivals (+ all expected bytes) complete a phase; wait(parity) passes once the phase of that parity
    has completed -- on a fresh barrier parity 1 passes at once."""

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

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        assert self.tx >= 0
        self._maybe_complete()

    def passed(self, parity):
        return self.phase != parity


class Model:
    def __init__(self, n_units_total, grid, block, nk, ncb, cbu, wide, rng):
        self.rng = rng
        self.nk, self.ncb, self.cbu, self.wide = nk, ncb, cbu, wide
        self.SA = 4 if cbu == 1 else 3
        self.SB = 4 if cbu == 1 else (2 if wide else 5)
        self.ncbu = (ncb + cbu - 1) // cbu
        self.units = list(range(block, n_units_total, grid))
        self.full_a = [MBar(PRODUCER_GROUP_WARPS) for _ in range(self.SA)]
        self.empty_a = [MBar(1) for _ in range(self.SA)]
        self.full_b = [MBar(1) for _ in range(self.SB)]
        self.empty_b = [MBar(1) for _ in range(self.SB)]
        self.tmem_full = [MBar(1), MBar(1)]
        self.tmem_empty = [MBar(EPILOGUE_WARPS), MBar(EPILOGUE_WARPS)]
        self.cfull = [MBar(1) for _ in range(CDEPTH)]
        self.cempty = [MBar(PRODUCER_GROUP_WARPS) for _ in range(CDEPTH)]
        # contents (tags) and hazards
        self.a_slot = [None] * self.SA            # stage counter whose Phi tiles the slot holds
        self.b_slot = [None] * self.SB            # (stage counter, block group)
        self.c_slot = [None] * CDEPTH             # k block of the centre tile
        self.a_readers = [0] * self.SA            # MMAs in flight that read the slot
        self.b_readers = [0] * self.SB
        self.acc_unit = [None, None]              # unit whose sums the accumulator buffer holds
        self.acc_busy = [False, False]            # complete and not yet released by the epilogue
        self.inflight = []                        # issued MMAs / commits, completed in order by the tensor core
        self.loads = []                           # bulk loads in flight (complete in any order)
        self.done_units = {"epi": [], "mma": []}

    def unit_blocks(self, u):
        cb0 = (u % self.ncbu) * self.cbu
        return cb0, min(self.cbu, self.ncb - cb0)

    # ---- roles (generators: `yield cond` blocks until cond() is true) ----------------------------------------------
    def tma(self):
        it = ic = ib = kc = 0
        total = len(self.units) * self.nk
        for u in self.units:
            cb0, nj = self.unit_blocks(u)
            for kb in range(self.nk):
                while ic < total and ic < it + CDEPTH:  # centre tiles run ahead; a busy slot is retried at the next stage
                    c = ic % CDEPTH
                    if not self.cempty[c].passed(((ic // CDEPTH) & 1) ^ 1):
                        break
                    self.cfull[c].arrive(tx=1)
                    self.loads.append(("c", c, kc))
                    ic += 1
                    kc = (kc + 1) % self.nk
                groups = [list(range(nj))] if self.wide else [[j] for j in range(nj)]
                for g in groups:
                    s = ib % self.SB
                    bar = self.empty_a[s] if self.cbu == 1 else self.empty_b[s]
                    par = ((ib // self.SB) & 1) ^ 1
                    yield lambda bar=bar, par=par: bar.passed(par)
                    self.full_b[s].arrive(tx=len(g))
                    for j in g:
                        self.loads.append(("b", s, (it, tuple(g)), cb0 + j, kb))
                    ib += 1
                    yield None
                it += 1

    def mma(self):
        it = ib = 0
        for unit_iter, u in enumerate(self.units):
            cb0, nj = self.unit_blocks(u)
            ab = (unit_iter & 1) if self.cbu == 1 else 0
            par = (((unit_iter >> 1) if self.cbu == 1 else unit_iter) & 1) ^ 1
            yield lambda ab=ab, par=par: self.tmem_empty[ab].passed(par)
            assert not self.acc_busy[ab], "MMA into accumulators the epilogue has not released"
            for kb in range(self.nk):
                s = it % self.SA
                yield lambda s=s, p=(it // self.SA) & 1: self.full_a[s].passed(p)
                groups = [tuple(range(nj))] if self.wide else [(j,) for j in range(nj)]
                for gi, g in enumerate(groups):
                    sb = ib % self.SB
                    yield lambda sb=sb, p=(ib // self.SB) & 1: self.full_b[sb].passed(p)
                    self.a_readers[s] += 1
                    self.b_readers[sb] += 1
                    self.inflight.append(("mma", s, sb, it, g, ab, u))
                    if self.cbu != 1:
                        self.inflight.append(("commit", self.empty_b[sb]))
                    if gi == len(groups) - 1:
                        self.inflight.append(("commit", self.empty_a[s]))
                        if kb == self.nk - 1:
                            self.inflight.append(("full", ab, u))
                    ib += 1
                    yield None
                it += 1
            self.done_units["mma"].append(u)

    def tensor_core(self):
        """completes the issued MMAs and commits in order, some time after they were issued"""
        while True:
            yield lambda: bool(self.inflight) or self.finished_issuing
            if not self.inflight:
                return
            op = self.inflight.pop(0)
            if op[0] == "mma":
                _, s, sb, it, g, ab, u = op
                assert self.a_slot[s] == it, f"MMA of stage {it} read Phi slot {s} holding {self.a_slot[s]}"
                assert self.b_slot[sb] == (it, g), f"MMA of stage {it} {g} read weight slot {sb} holding {self.b_slot[sb]}"
                self.a_readers[s] -= 1
                self.b_readers[sb] -= 1
                self.acc_unit[ab] = u
            elif op[0] == "commit":
                op[1].arrive()
            else:
                _, ab, u = op
                assert self.acc_unit[ab] == u
                self.acc_busy[ab] = True
                self.tmem_full[ab].arrive()

    def loader(self):
        """lands the bulk loads in any order"""
        while True:
            yield lambda: bool(self.loads) or self.finished_issuing
            if not self.loads:
                return
            op = self.loads.pop(self.rng.randrange(len(self.loads)))
            if op[0] == "c":
                _, c, kc = op
                self.c_slot[c] = kc
                self.cfull[c].complete_tx(1)
            else:
                _, s, tag, cb, kb = op
                assert self.b_readers[s] == 0, "a weight tile landed in a slot an MMA in flight reads"
                self.b_slot[s] = tag
                self.full_b[s].complete_tx(1)

    def producer(self, grp, warp):
        it = 0
        for u in self.units:
            it0 = it
            it += self.nk
            for kb in range((grp ^ it0) & 1, self.nk, 2):
                itk = it0 + kb
                s, cs = itk % self.SA, itk % CDEPTH
                yield lambda cs=cs, p=(itk // CDEPTH) & 1: self.cfull[cs].passed(p)
                yield lambda s=s, p=((itk // self.SA) & 1) ^ 1: self.empty_a[s].passed(p)
                assert self.c_slot[cs] == kb, f"stage {itk} (k block {kb}) found centre tile {self.c_slot[cs]}"
                assert self.a_readers[s] == 0, "Phi slot overwritten while an MMA in flight reads it"
                self.a_slot[s] = itk
                yield None
                self.full_a[s].arrive()
                self.cempty[cs].arrive()

    def epilogue(self, warp):
        for unit_iter, u in enumerate(self.units):
            ab = (unit_iter & 1) if self.cbu == 1 else 0
            par = ((unit_iter >> 1) if self.cbu == 1 else unit_iter) & 1
            yield lambda ab=ab, par=par: self.tmem_full[ab].passed(par)
            assert self.acc_unit[ab] == u and self.acc_busy[ab], f"epilogue of unit {u} found {self.acc_unit[ab]}"
            yield None  # tcgen05.ld of the warp's lanes
            if self.tmem_empty[ab].pending == 1:
                self.acc_busy[ab] = False  # the last warp's arrival releases the buffer
            self.tmem_empty[ab].arrive()
            if warp == 0:
                self.done_units["epi"].append(u)

    # ---- scheduler --------------------------------------------------------------------------------------------------
    def run(self):
        self.finished_issuing = False
        issuing = {"tma": self.tma(), "mma": self.mma()}
        for g in range(2):
            for w in range(PRODUCER_GROUP_WARPS):
                issuing[f"p{g}.{w}"] = self.producer(g, w)
        for w in range(EPILOGUE_WARPS):
            issuing[f"e{w}"] = self.epilogue(w)
        engines = {"tc": self.tensor_core(), "ld": self.loader()}
        roles = dict(issuing, **engines)
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
                raise RuntimeError(f"deadlock after {steps} steps; waiting: {sorted(roles)}")
            name = self.rng.choice(ready)
            try:
                waiting[name] = roles[name].send(None)
            except StopIteration:
                del roles[name]
            steps += 1
        assert self.done_units["epi"] == self.units and self.done_units["mma"] == self.units
        assert not self.inflight and not self.loads
        return steps


def check(n_vt, ncb, nk, grid, cbu, wide, seed):
    rng = random.Random(seed)
    ncbu = (ncb + cbu - 1) // cbu
    n_units = n_vt * ncbu
    total = 0
    for block in range(min(grid, n_units)):
        total += Model(n_units, min(grid, n_units), block, nk, ncb, cbu, wide, rng).run()
    return total


if __name__ == "__main__":
    import sys
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    for cbu, wide in ((1, False), (2, False), (2, True)):
        for ncb in (1, 2, 3, 6, 7):
            for nk in (1, 2, 3, 9, 20):
                steps = check(n_vt=5, ncb=ncb, nk=nk, grid=3, cbu=cbu, wide=wide, seed=seed)
                print(f"CBU={cbu} wide={int(wide)} blocks={ncb} stages={nk}: ok ({steps} scheduler steps)")
