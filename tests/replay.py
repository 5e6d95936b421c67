"""Replays a tests/golden/env_*.npz fixture (outputs of the live reference) through a backend.

A backend exposes reset(tick, noise_tape=None) / step(tick, noise_tape=None) returning the
oracle-format dict (arrays with a leading env axis of size 1) and state().  Both the C oracle
(`-m "not gpu"`) and the CUDA library through its C-ABI (`-m gpu`) are checked this way.
"""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PACKED = ("objects", "destination", "state_objects")


def fixture_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "env_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"env_{name}.npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    for k in PACKED:
        shape = tuple(d.pop(k + "_shape"))
        d[k] = np.unpackbits(d[k])[: int(np.prod(shape))].reshape(shape).astype(np.int8)
    d["config"] = json.loads(str(d["config"]))
    d["oracle_kwargs"] = json.loads(str(d["oracle_kwargs"]))
    d["env_id"] = int(d["env_id"])
    return d


def replay(fix, backend, check_extras=True, check_state=True):
    """Bit-exact comparison of every recorded call; returns the number of calls checked."""
    has_noise_tape = "noise" in fix
    n = len(fix["ticks"])
    for s in range(n):
        tick = int(fix["ticks"][s])
        tape = fix["noise"][s][None] if has_noise_tape else None
        out = backend.reset(tick, noise_tape=tape) if fix["is_reset"][s] else backend.step(tick, noise_tape=tape)
        keys = ["obs", "dones"] + (["obs_rel"] if "obs_rel" in fix else [])
        if check_extras:
            keys += ["coordinates", "agents", "objects", "destination"]
        for k in keys:
            got, want = np.asarray(out[k][0]), fix[k][s]
            if not np.array_equal(got, want):
                idx = np.argwhere(got != want)[0]
                raise AssertionError(f"call {s} (tick {tick}, reset={int(fix['is_reset'][s])}): {k} differs at "
                                     f"{idx.tolist()}: got {got[tuple(idx)]} want {want[tuple(idx)]}")
        # the reference returns python floats / np.float64 (shared mean); the learner casts with .float()
        want_r = fix["rewards"][s].astype(np.float32)
        assert np.array_equal(np.asarray(out["rewards"][0], np.float32), want_r), (s, out["rewards"][0], want_r)
        if check_state:
            st = backend.state()
            assert np.array_equal(st["pos"][0], fix["pos"][s]), (s, "pos")
            assert np.array_equal(st["objects"][0], fix["state_objects"][s]), (s, "objects")
            assert np.array_equal(st["counters"][0], fix["counters"][s]), (s, "counters", st["counters"][0], fix["counters"][s])
    return n
