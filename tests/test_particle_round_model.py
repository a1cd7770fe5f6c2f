"""The argument for the exactness of the particle rounds (csrc/fss_particles.cuh), checked on a CPU model of the scheme under
hostile schedules and colliding tables: see tests/model_particle_rounds.py."""
import numpy as np
import pytest

import model_particle_rounds as m


def _same(a, b):
    ka, aa, la = a
    kb, ab, lb = b
    return np.array_equal(ka, kb) and aa.tobytes() == ab.tobytes() and [p.key() for p in la] == [p.key() for p in lb]


@pytest.mark.parametrize("wake", [True, False])
@pytest.mark.parametrize("visibility", ["late", "now", "random"])
@pytest.mark.parametrize("slots,tile_slots,budget", [(7, 3, 1 << 30), (61, 7, 1 << 30), (100003, 1009, 1 << 30), (61, 7, 12), (100003, 1009, 4)])
def test_rounds_give_the_serial_result(visibility, slots, tile_slots, budget, wake):
    for seed in range(4):
        rng = np.random.default_rng(1000 * seed + slots + budget % 97)
        kind, amount, parts = m.random_scene(rng)
        want = m.serial(kind, amount, parts)
        kg, ag, left, n_rounds = m.rounds(kind, amount, parts, rng, slots=slots, tile_slots=tile_slots, visibility=visibility, budget=budget, wake=wake)
        assert _same((kg, ag, left), want), f"seed {seed}: differs from the serial loop after {n_rounds} rounds"
        assert n_rounds <= len(parts)


def test_the_model_can_fail():
    """a broken scheme (no check of the cells a lower undecided particle looked at before writing) is caught by the same
    comparison: the test above is not vacuous"""
    bad = 0
    for seed in range(12):
        rng = np.random.default_rng(seed)
        kind, amount, parts = m.random_scene(rng, n=200)
        want = m.serial(kind, amount, parts)

        class Careless(m.Tables):
            def clear(self):
                super().clear()
                self.r_act = _Never()
                self.r_pot = _Never()
                self.w_pot = _Never()

        orig = m.Tables
        m.Tables = Careless
        try:
            got = m.rounds(kind, amount, parts, rng, slots=100003, tile_slots=1009)[:3]
        except AssertionError:
            got = None
        finally:
            m.Tables = orig
        bad += got is None or not _same(got, want)
    assert bad > 0


class _Never(list):
    """a table that forgets everything it is told"""

    def __getitem__(self, k):
        return 1 << 30

    def __setitem__(self, k, v):
        pass
