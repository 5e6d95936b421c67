"""CUDA library (through its C-ABI) presented in the oracle's output format, for parity tests."""
import numpy as np
import torch

from maexp_b200.config import Config
from maexp_b200.core.environments.batched import BatchedEnvironment
from maexp_b200.core.maps import generate_map
from maexp_b200.core.worlds import generate_world


def make_env(cfg, num_envs, **kw):
    cfg = cfg if isinstance(cfg, Config) else Config(cfg)
    world = generate_world(cfg, generate_map(cfg))
    return BatchedEnvironment(cfg, world, num_envs, **kw)


class CudaBackend:
    def __init__(self, cfg, num_envs=1, first_global_env_id=0, auto_reset=False, extras=True, **kw):
        self.env = make_env(cfg, num_envs, first_global_env_id=first_global_env_id, auto_reset=auto_reset, **kw)
        self.extras = extras
        self.out = self.env.new_outputs(extras=extras)

    def _tape(self, noise_tape):
        if noise_tape is None:
            return None
        e = self.env
        return torch.as_tensor(np.ascontiguousarray(noise_tape, np.float32)).reshape(e.E, e.N, e.K).to(e.device)

    def _convert(self):
        e, o = self.env, self.out
        torch.cuda.synchronize()
        r = {"obs": e.view(o["obs"]).cpu().numpy(),
             "obs_rel": e.view(o["obs_rel"]).cpu().numpy() if e.merged else None,
             "rewards": o["rewards"].cpu().numpy(), "dones": o["dones"].cpu().numpy(),
             "was_reset": o["was_reset"].cpu().numpy()}
        if self.extras:
            pos = o["positions"].cpu().numpy()
            r["coordinates"] = o["coordinates"].cpu().numpy()
            r["agents"] = np.repeat(pos[:, None], e.N, axis=1)
            r["objects"] = o["objects"].cpu().numpy()
            r["destination"] = e.destination_masks[o["destination"].cpu().numpy().astype(np.int64)]
        return r

    def reset(self, tick, noise_tape=None, env_mask=None):
        if env_mask is not None:
            env_mask = torch.as_tensor(np.asarray(env_mask, np.uint8)).to(self.env.device)
        self.env.reset(self.out, tick=tick, env_mask=env_mask, noise_tape=self._tape(noise_tape))
        return self._convert()

    def step(self, tick, actions=None, order=None, noise_tape=None):
        d = self.env.device
        if actions is not None:
            actions = torch.as_tensor(np.asarray(actions, np.int8)).reshape(self.env.E, self.env.N).to(d)
        if order is not None:
            order = torch.as_tensor(np.asarray(order, np.uint8)).reshape(self.env.E, self.env.N).to(d)
        self.env.step(self.out, actions=actions, order=order, tick=tick, noise_tape=self._tape(noise_tape))
        return self._convert()

    def state(self):
        return {k: v.cpu().numpy() for k, v in self.env.state().items()}
