"""Pins the C oracle (oracle/maexp_oracle.c) against outputs of the live reference.

The reference ships no tests or golden vectors; tests/golden/*.npz were
produced by oracle/make_golden.py from the reference itself under injected draws.
"""
import hashlib

import numpy as np
import pytest

from maexp_b200.config import make_config
from oracle.oracle import Oracle
from tests import replay as rp


class OracleBackend:
    def __init__(self, fix):
        self.o = Oracle(fix["config"], num_envs=1, first_global_env_id=fix["env_id"], extras=True,
                        **fix["oracle_kwargs"])

    def reset(self, tick, noise_tape=None):
        return self.o.reset(tick, noise_tape=noise_tape)

    def step(self, tick, noise_tape=None):
        return self.o.step(tick, noise_tape=noise_tape)

    def state(self):
        return self.o.state()


@pytest.mark.parametrize("name", rp.fixture_names())
def test_oracle_replays_reference_fixture(name):
    fix = rp.load(name)
    b = OracleBackend(fix)
    wall, area, los = b.o.tables()
    assert np.array_equal(wall, fix["tables_wall"])
    assert np.array_equal(area, fix["tables_area"])
    assert np.array_equal(los, fix["tables_los"])
    assert rp.replay(fix, b) == len(fix["ticks"])


# sha256(table.tobytes())[:16] of the reference's own tables
SURVEY_HASHES = {
    ("simple", 20, 20): ("674c9dbee04e5882", {5: (7056, "d7926a60a84bedde"), 7: (12996, "b3ea4bb221bcd999"), 9: (20164, "413060f2786cf096")}),
    ("four_rooms", 24, 24): ("729292878533450c", {5: (9196, "f9fa5372b80e3b3e"), 7: (16152, "718fdc57397b85f7"), 9: (23772, "60c4b71fc6f42361")}),
    ("three_rooms", 25, 25): ("27558b5e477b3a19", {5: (10522, "d6c24ecd8d180331"), 7: (18758, "474af93509a91b30"), 9: (27954, "87d23c81cdd6e236")}),
    ("four_rectangle", 40, 24): ("3f3150e7d54be01c", {5: (16144, "61cd9689238a1e30"), 7: (28692, "2d6c02a78099ec87"), 9: (42800, "9dda4b3f15b49d62")}),
}
TRANSPARENT = {
    "four_rooms": {5: (9636, "89623c2b5d6b3e68"), 7: (17952, "f0bff3b0f255abba"), 9: (28120, "1e37a3ee91b172fb")},
    "three_rooms": {5: (10896, "16401268812e8e51"), 7: (20346, "f45467c43bff44d3"), 9: (31972, "fe0337bb1748a2b0")},
    "four_rectangle": {5: (16856, "287a3073736d3fbc"), 7: (31700, "9a8f5a8efeb611c4"), 9: (50260, "145fa83a96b0600f")},
}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("key", list(SURVEY_HASHES))
@pytest.mark.parametrize("V", [5, 7, 9])
def test_oracle_tables_match_reference_hashes(key, V):
    name, X, Y = key
    wall_sha, los = SURVEY_HASHES[key]
    o = Oracle(make_config(map=dict(name=name, SIZE_X=X, SIZE_Y=Y), visible_range=V), num_envs=1)
    wall, _, table = o.tables()
    assert _sha(wall) == wall_sha
    assert (int(table.sum()), _sha(table)) == los[V]
    if name in TRANSPARENT:
        o = Oracle(make_config(map=dict(name=name, SIZE_X=X, SIZE_Y=Y), visible_range=V,
                               transparent_observation=True), num_envs=1)
        table = o.tables()[2]
        assert (int(table.sum()), _sha(table)) == TRANSPARENT[name][V]
