"""Index model of the CTC kernel's phase-1 lattice ring (csrc/ctc.cu, v4): for every sequence length and chunk size
the rows are fetched once each, PD rows ahead, into slots that no pending reader still needs, and the post-pass
reads the slot its row was written to.  Pure Python (CPU): it guards the arithmetic the three template
instantiations (32/16/8 frames) share."""
import pytest

PD = 4


def _simulate(Tn, TC, direction):
    EM = TC + PD
    tstar = Tn >> 1
    i_begin = tstar if direction == 0 else Tn - tstar
    i_end = Tn
    slot_row = [None] * EM          # which processing index a slot holds
    fetched = []
    pf_i, pf_slot = i_begin, 0

    def fetch_next():
        nonlocal pf_i, pf_slot
        if pf_i < i_end:
            fetched.append(pf_i)
            slot_row[pf_slot] = pf_i
        pf_i += 1
        pf_slot = 0 if pf_slot + 1 == EM else pf_slot + 1

    for _ in range(PD):
        fetch_next()
    slot = 0
    nchunks = (i_end - i_begin + TC - 1) // TC
    consumed = []
    for ci in range(nchunks):
        ic = i_begin + ci * TC
        ie = min(ic + TC, i_end)
        n = ie - ic
        slot0 = slot
        live = {}                    # slot -> row of this chunk already consumed (needed by the post-pass)
        for i in range(ic, ie):
            # the prefetch of this step must not land on a slot the post-pass of this chunk still needs
            target = pf_slot
            will_write = pf_i < i_end
            fetch_next()
            assert not (will_write and target in live), (Tn, TC, direction, i)
            assert slot_row[slot] == i, (Tn, TC, direction, i, slot_row[slot])
            live[slot] = i
            consumed.append(i)
            slot = 0 if slot + 1 == EM else slot + 1
        # post-pass: lane r (natural time order inside the chunk) reads slot0 + (r or n-1-r)
        for r in range(n):
            sidx = slot0 + (r if direction == 0 else n - 1 - r)
            if sidx >= EM:
                sidx -= EM
            i_of_r = ic + (r if direction == 0 else n - 1 - r)
            assert slot_row[sidx] == i_of_r and live.get(sidx) == i_of_r
    assert consumed == list(range(i_begin, i_end))
    assert fetched == list(range(i_begin, i_end))


@pytest.mark.parametrize("TC", [32, 16, 8])
def test_ring_slots_for_all_lengths(TC):
    for Tn in list(range(1, 140)) + [255, 256, 257, 998, 1000, 1001]:
        for direction in (0, 1):
            _simulate(Tn, TC, direction)
