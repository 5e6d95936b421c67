"""Writes tests/golden/env_<case>.npz by running the LIVE reference (build container only).

    python -m oracle.make_golden            # from the repo root

TEST INFRASTRUCTURE.  The fixtures hold the reference's own outputs for injected
actions / orders / draws; tests/test_oracle_golden.py replays them through the C oracle on
any machine (the GPU box has no /root/reference).
"""
import hashlib
import json
import os
import sys

import numpy as np

from oracle import ref_runner as rr
from oracle.cases import cases

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ENV_ID = 3
KEYS = ("obs", "obs_rel", "rewards", "dones", "coordinates", "agents", "objects", "destination")


def run_case(cfg, steps, noise_kind, env_id=ENV_ID, seed=5):
    ref = rr.RefEnv(cfg, env_id=env_id)
    rng = np.random.default_rng(seed)
    shape = (ref.N, ref.C, ref.V, ref.V)
    rec = {k: [] for k in KEYS}
    rec.update(ticks=[], is_reset=[], noise=[], pos=[], state_objects=[], counters=[])
    tables = ref.tables()  # before the first step: dynamic_simple moves its obstacle cells in wall_matrix afterwards

    def tape():
        return rng.standard_normal(shape).astype(np.float32) if noise_kind == "normal" else None

    def push(out, tick, is_reset, t):
        for k in KEYS:
            if out[k] is not None:
                rec[k].append(out[k])
        rec["ticks"].append(tick)
        rec["is_reset"].append(is_reset)
        if t is not None:
            rec["noise"].append(t)
        st = ref.state()
        rec["pos"].append(st["pos"]); rec["state_objects"].append(st["objects"]); rec["counters"].append(st["counters"])

    tick = 0
    t = tape(); push(ref.reset(tick, t), tick, 1, t)
    for _ in range(steps):
        tick += 1
        t = tape(); out = ref.step(tick, noise_tape=t); push(out, tick, 0, t)
        if out["dones"][0]:  # default_trainer.py:85,99 — episode ends on dones[0], then reset()
            tick += 1
            t = tape(); push(ref.reset(tick, t), tick, 1, t)
    arrays = {k: np.stack(v) for k, v in rec.items() if len(v)}
    arrays["tables_wall"], arrays["tables_area"], arrays["tables_los"] = tables
    return arrays


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])
    mpath = os.path.join(GOLDEN, "MANIFEST.json")
    manifest = json.load(open(mpath))["files"] if only and os.path.exists(mpath) else {}
    for name, (cfg, steps, noise_kind, okw) in cases().items():
        if only and name not in only:
            continue
        arrays = run_case(cfg, steps, noise_kind)
        # objects/destination snapshots dominate the size: keep them bit-packed along Y
        for k in ("objects", "destination", "state_objects"):
            arrays[k + "_shape"] = np.array(arrays[k].shape, np.int64)
            arrays[k] = np.packbits(arrays[k].astype(np.uint8), axis=None)
        path = os.path.join(GOLDEN, f"env_{name}.npz")
        np.savez_compressed(path, config=json.dumps(dict(cfg)), noise_kind=str(noise_kind),
                            oracle_kwargs=json.dumps(okw), env_id=ENV_ID, **arrays)
        manifest[name] = hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
        print(name, os.path.getsize(path), "bytes", arrays["obs"].shape)
    with open(mpath, "w") as f:
        json.dump({"generator": "python -m oracle.make_golden", "reference": "Yoshi-0921/MAEXP (live import)",
                       "files": manifest}, f, indent=1)


if __name__ == "__main__":
    main()
