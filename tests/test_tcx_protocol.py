"""The synchronisation protocol of the exact-digit evaluation kernel (facedeform_b200/csrc/fd_eval_tcx.cu) under random
schedules of its roles: tests/tools/tcx_protocol_model.py restates the kernel's barriers, counts, ring sizes, slot indices and
wait parities; here it runs over the shapes the kernel meets (the vertex loop of SOP_FaceDeform.cpp:404-439 for 1 ... 280
frames = 1 ... 7 column blocks, 1 ... 20 stages of 32 centres, fewer units than CTAs), for the shipped form and the two
development forms.  No GPU; compute-sanitizer is closed on the GPU pool, so this is the independent look at the protocol."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
import tcx_protocol_model as model  # noqa: E402

FORMS = [(1, False), (2, False), (2, True)]  # (column blocks per unit, one N = 240 MMA for both)


@pytest.mark.parametrize("cbu,wide", FORMS)
@pytest.mark.parametrize("ncb", [1, 2, 3, 6, 7])
def test_no_schedule_deadlocks_or_reads_the_wrong_tile(cbu, wide, ncb):
    for seed, nk in enumerate((1, 2, 3, 9, 20)):
        assert model.check(n_vt=5, ncb=ncb, nk=nk, grid=3, cbu=cbu, wide=wide, seed=seed) > 0


@pytest.mark.parametrize("cbu,wide", FORMS)
def test_more_ctas_than_units_and_single_unit(cbu, wide):
    assert model.check(n_vt=1, ncb=1, nk=1, grid=148, cbu=cbu, wide=wide, seed=3) > 0
    assert model.check(n_vt=2, ncb=5, nk=4, grid=148, cbu=cbu, wide=wide, seed=4) > 0


def test_the_model_sees_a_protocol_error():
    """sensitivity: freeing the Phi slot before the stage's MMAs have completed must be caught"""

    class FreesTooEarly(model.Model):
        def mma(self):
            for step in super().mma():
                # move the Phi slot's commit in front of the stage's MMAs still in flight
                for i, op in enumerate(self.inflight):
                    if op[0] == "commit" and op[1] in self.empty_a:
                        self.inflight.insert(0, self.inflight.pop(i))
                        break
                yield step

    caught = 0
    for seed in range(8):
        try:
            FreesTooEarly(20, 3, 0, 9, 7, 2, False, random.Random(seed)).run()
        except (AssertionError, RuntimeError):
            caught += 1
    assert caught == 8


# ---- tc::k_eval_tc (FP16 hi/lo tensor-core evaluation): the CTA-pair protocol -----------------------------------------------
import tc_pair_protocol_model as pair_model  # noqa: E402


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("n_units", [1, 2, 5])
def test_pair_protocol_no_deadlock_and_the_right_halves(pair, n_units):
    """each CTA loads half of every weight tile and multicasts it to both; a slot is free when the MMAs of both have read it"""
    for seed, nk in enumerate((1, 2, 3, 9, 20)):
        assert pair_model.check(n_units, nk, pair, seed) > 0


def test_the_pair_model_sees_a_slot_freed_by_one_cta_alone():
    """sensitivity: with the `empty` barriers counting ONE commit in pair mode (the single-CTA value) a CTA refills a slot its
    peer's MMAs may still read, or runs a phase ahead -- caught as a wrong / overwritten tile or a deadlock"""

    class OneCommitFrees(pair_model.PairModel):
        def __init__(self, *a):
            super().__init__(*a)
            for c in self.ctas:
                c.empty = [pair_model.MBar(1) for _ in range(pair_model.STAGES)]

    caught = 0
    for seed in range(8):
        try:
            OneCommitFrees(5, 9, True, random.Random(seed)).run()
        except (AssertionError, RuntimeError):
            caught += 1
    assert caught == 8
