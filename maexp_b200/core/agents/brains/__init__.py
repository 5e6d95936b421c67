"""Brains: TD target, loss, clipping and optimizer step on the GPU (maexp_learner_* C-ABI).

`LearnerGroup` updates G = num_agents independent networks in grouped launches; `generate_brain`
keeps the reference's per-agent factory (core/agents/brains/__init__.py:42-85) on top of it.
"""
import ctypes as C
import os

import numpy as np
import torch

from .... import _lib
from ...utils.logging import initialize_logging
from ..models import build_reference_modules

logger = initialize_logging(__name__)

# brain name -> (learner kind for conv models, clip max_norm)      brains/*.py
_BRAINS = {"dqn": ("dqn_conv_mlp", 0.1), "da3_baseline": ("dqn_conv_mlp", 0.1), "categorical_dqn": ("c51", 10.0),
           "qr_dqn": ("qr", 0.1), "iqn": ("iqn", 0.1), "da3": ("da3", 0.1), "da6": ("da6", 0.1),
           "da3_iqn": ("iqn", 0.1)}  # DA3_IQNBrain(IQNBrain) keeps IQNBrain.learn (da3_iqn_brain.py:14)

IQN_TRUNKS = {"iqn": 0, "da3_iqn": 1, "merged_da3_iqn": 2, "da6_iqn": 3}


def iqn_trunk(config):
    """Which network produces the state embedding under the IQN head (models/__init__.py:51-72)."""
    name, mask = config.model.name, config.observation_area_mask
    if name == "da3_iqn" and mask == "merged":
        return "merged_da3_iqn"
    return name


def learner_kind(config):
    if config.brain not in _BRAINS:
        logger.warning(f"Unexpected brain is given. config.brain: {config.brain}")
        raise ValueError()
    kind, max_norm = _BRAINS[config.brain]
    if config.model.name == "mlp":
        if kind != "dqn_conv_mlp":
            raise ValueError("model=mlp is only defined for the dqn brain")
        kind = "dqn_mlp"
    elif config.model.name in ("da3_iqn", "da6_iqn"):
        if config.brain not in ("da3_iqn", "iqn"):
            raise ValueError("model da3_iqn / da6_iqn needs an IQN brain (brain=da3_iqn)")
        mask = config.observation_area_mask
        if config.model.get("attention") == "aoa":
            raise ValueError("attention IQN models are implemented with soft / hard-shrink attention")
        if config.model.name == "da6_iqn" and mask != "merged":
            raise ValueError("da6_iqn reads the merged view (customs/da6_iqn.py:160-176)")
        if config.model.name == "da3_iqn" and mask == "local" and int(config.model.patch_size) != 1:
            raise ValueError("da3_iqn on the local view is implemented for patch_size=1")
    elif config.model.name == "da6" or kind == "da6":
        if kind != "da6" or config.model.name != "da6":
            raise ValueError("brain=da6 and model=da6 go together (da6_brain.py, customs/da6.py)")
        if config.observation_area_mask != "merged":
            raise ValueError("DA6 reads the merged view (customs/da6.py:161-179)")
    elif config.model.name == "da3" or kind == "da3":
        if kind != "da3" or config.model.name != "da3":
            raise ValueError("brain=da3 and model=da3 go together (da3_brain.py, customs/da3.py)")
        if config.observation_area_mask == "local" and int(config.model.patch_size) != 1:
            raise ValueError("DA3 on the local view is implemented for patch_size=1")
    elif config.model.name not in ("conv_mlp", "da3_baseline", "categorical_dqn", "qr_dqn", "iqn"):
        logger.warning(f"Unexpected network is given. config.model.name: {config.model.name}")
        raise ValueError()
    if config.observation_area_mask != "local" and kind not in ("dqn_conv_mlp", "da3", "da6", "iqn"):
        raise ValueError("relative / merged views are implemented for the conv_mlp DQN, IQN, DA3 / DA6 and their IQN variants")
    if config.observation_area_mask == "local" and kind in ("dqn_conv_mlp", "da3", "iqn") and canvas_flags(config, kind)[0]:
        # the reference sizes the first layer for C + extra channels but feeds it the C-channel crop (conv_mlp.py:25-27,61)
        raise ValueError("objects_channel / agents_channel / destination_channel need the relative or merged view")
    return kind, max_norm


AREA_MASKS = {"local": 0, "relative": 1, "merged": 2}


def los_table(config, world=None):
    """int8 [X,Y,V,V] line-of-sight table of the configuration's map (local_area_mask.py:50-97)."""
    from ...handlers.observations.masks import generate_observation_area_mask
    if world is None:
        from ...maps import generate_map
        from ...worlds import generate_world
        world = generate_world(config, generate_map(config))
    return generate_observation_area_mask(config.merged_with(observation_area_mask="local"), world)


def canvas_flags(config, kind):
    """(extra in_chans the model family's constructor adds, (objects, agents, destination) it reads in state_encoder).
    conv_mlp.py:25-27,61; da3.py:31-33,104; iqn.py:55-60,153-159; da3_iqn.py:96-104,195-211; da6.py:32-37,170-177;
    da6_iqn.py:42-47,167-174.  The merged subclasses keep the parent's first layer (sized with the flags) unused."""
    o, a, d = (bool(config.get(k)) for k in ("objects_channel", "agents_channel", "destination_channel"))
    name = config.model.name
    if kind == "dqn_conv_mlp" or (kind == "da3" and name == "da3"):
        a = d = False
    elif kind == "da6" or (kind == "iqn" and name == "iqn"):
        d = False
    elif kind == "iqn" and name in ("da3_iqn", "da6_iqn"):
        pass
    else:
        o = a = d = False
    extra = (int(config.type_objects) if o else 0) + (int(config.num_agents) if a else 0) + int(d)
    reads = config.observation_area_mask == "relative" or name in ("da6", "da6_iqn")
    if name == "da6_iqn":
        d = False  # its own relative stack takes objects + agents only; the destination flag sizes the unused parent
    return extra, (o and reads, a and reads, d and reads)


class LearnerGroup:
    def __init__(self, config, obs_shape, act_size, num_agents, batch_size, obs_stride=None, device="cuda:0"):
        if not torch.cuda.is_available():
            raise _lib.MaexpError("LearnerGroup needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.config, self.device = config, torch.device(device)
        self.kind, max_norm = learner_kind(config)
        C_, V = int(obs_shape[0]), int(config.visible_range)  # batches always hold the V x V crops
        self.G, self.B, self.A, self.C, self.V = int(num_agents), int(batch_size), int(act_size), C_, V
        self.K = C_ * V * V
        self.obs_stride = int(obs_stride or (self.K + 7) // 8 * 8)
        m, opt = config.model, config.opt
        if opt.name not in ("rmsprop", "adam"):  # optimizers/__init__.py:17-42
            logger.warning(f"Unexpected optimizer name is given. config.opt.name: {opt.name}")
            raise ValueError()
        cfg = _lib.LearnerCfg()
        cfg.kind = _lib.LEARNER_KINDS[self.kind]
        cfg.num_agents, cfg.batch_size, cfg.channels, cfg.visible_range = self.G, self.B, C_, V
        cfg.num_actions = self.A
        self.area_mask = config.observation_area_mask
        cfg.area_mask = AREA_MASKS[self.area_mask]
        cfg.size_x, cfg.size_y = int(config.map.SIZE_X), int(config.map.SIZE_Y)
        cfg.patch_size = int(m.get("patch_size") or 1)
        cfg.hidden = 0 if self.kind == "da6" or (self.kind == "da3" and self.area_mask != "merged") else \
            int(m.hidden_layer_sizes[0])
        self.trunk = iqn_trunk(config) if self.kind == "iqn" else None
        if self.kind == "da6":
            cfg.relative_patch_size, cfg.local_patch_size = int(m.relative_patch_size), int(m.local_patch_size)
        if self.kind == "iqn":
            cfg.iqn_trunk = IQN_TRUNKS[self.trunk]
            if self.trunk == "da6_iqn":  # hard-coded in da6_iqn.py:50-58
                cfg.relative_patch_size, cfg.local_patch_size = 5, 1
        if self.kind in ("da3", "da6") or self.trunk in ("da3_iqn", "merged_da3_iqn", "da6_iqn"):
            cfg.embed_dim, cfg.num_heads, cfg.block_loop = int(m.embed_dim), int(m.num_heads), int(m.block_loop)
            cfg.mlp_hidden = int(int(m.embed_dim) * float(m.mlp_ratio))
            # HardShrinkBlock (hard_shrink_attention.py:13-49): da3.py:46,144, da6.py:67, da3_iqn.py:118-122,245, da6_iqn.py:76-80
            if m.get("attention") == "hard":
                cfg.attention_hard, cfg.af_lambd = 1, float(m.af_lambd)
        cfg.conv1_channels = int(m.get("hidden_channels", [32])[0])
        cfg.conv2_channels = int(m.get("output_channel", 64))
        self.Z = int(m.num_atoms) if self.kind == "c51" else int(m.num_quantiles) if self.kind in ("qr", "iqn") else 1
        cfg.num_atoms, cfg.obs_stride = self.Z, self.obs_stride
        if self.kind == "iqn":
            self.Nd = int(m.num_quantiles_dash)
            cfg.num_quantiles_dash, cfg.num_cosines = self.Nd, int(m.num_cosines)
        cfg.gamma = float(config.gamma)
        if self.kind == "c51":
            cfg.v_min, cfg.v_max = float(m.V_min), float(m.V_max)
            cfg.delta_z = (float(m.V_max) - float(m.V_min)) / (self.Z - 1)  # python double -> fp32, as torch does
        cfg.learning_rate, cfg.eps = float(opt.learning_rate), float(opt.eps)
        if opt.name == "adam":
            cfg.optimizer, cfg.beta1, cfg.beta2 = 1, float(opt.betas[0]), float(opt.betas[1])
        else:
            cfg.alpha, cfg.one_minus_alpha = float(opt.alpha), 1.0 - float(opt.alpha)
        cfg.max_norm = max_norm
        for k in ("objects_channel", "agents_channel", "destination_channel"):
            setattr(cfg, k, int(bool(config.get(k))))
        cfg.type_objects, cfg.team_agents = int(config.type_objects), int(config.num_agents)
        self.T, self.N_team, self.X, self.Y = cfg.type_objects, cfg.team_agents, cfg.size_x, cfg.size_y
        self.canvas_reads = canvas_flags(config, self.kind)[1]  # which of objects / agents / destination the forward pass reads
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.maexp_learner_create(C.byref(cfg), C.byref(handle)))
        self.h = handle
        self.P = int(self.lib.maexp_learner_num_params(self.h))
        self.table = []
        for i in range(self.lib.maexp_learner_num_tensors(self.h)):
            name, off, cnt = C.create_string_buffer(96), C.c_int64(), C.c_int64()
            _lib.check(self.lib.maexp_learner_tensor(self.h, i, name, 96, C.byref(off), C.byref(cnt)))
            self.table.append((name.value.decode(), off.value, cnt.value))
        self.loss = torch.zeros((self.G,), dtype=torch.float32, device=self.device)
        self.grad_norm = torch.zeros((self.G,), dtype=torch.float32, device=self.device)
        self.q = torch.zeros((self.G, self.B, self.A), dtype=torch.float32, device=self.device)
        # support / tau_hat vector exactly as the reference brains build it
        if self.kind == "c51":  # categorical_dqn_brain.py:24-26
            self._write_support(torch.linspace(float(m.V_min), float(m.V_max), self.Z))
        elif self.kind == "qr":  # qr_dqn_brain.py:59-65
            taus = torch.arange(0, self.Z + 1, dtype=torch.float32) / self.Z
            self._write_support((taus[1:] + taus[:-1]) / 2.0)
        # PyTorch default initialisation, target <- online (abstract_brain.py:33)
        for g in range(self.G):
            self.load_state_dict(g, build_reference_modules(config, obs_shape, act_size).state_dict())
        self.synchronize()

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.lib.maexp_learner_destroy(h)

    # ------------------------------------------------------------------ buffers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _buffer(self, which, count):
        """Zero-copy torch view of a library-owned device buffer."""
        ptr = self.lib.maexp_learner_buffer(self.h, which)

        class _Holder:
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Holder(), device=self.device)

    def params(self, which=_lib.BUF_ONLINE):
        return self._buffer(which, self.G * self.P).view(self.G, self.P)

    def optimizer_state(self):
        """The optimizer's buffers [G, P] (RMSprop: square_avg; Adam: exp_avg, exp_avg_sq and the int32 step counters)."""
        bufs = [self.params(_lib.BUF_SQUARE_AVG)]
        if self.config.opt.name == "adam":
            bufs.append(self.params(_lib.BUF_EXP_AVG))
            bufs.append(self._buffer(_lib.BUF_ADAM_STEPS, self.G))  # int32 bits behind a float view: copied verbatim
        return bufs

    def broadcast(self, dist, process_group=None, src=0):
        """Data-parallel replicas must start from the same weights and optimizer state: every rank initialises its networks
        from its own torch RNG (build_reference_modules), so rank `src` sends its online and target parameters and the
        optimizer buffers to the others.  Call after construction and after load_state_dict."""
        from .... import distributed as mdist
        mdist.broadcast_tensors(dist, [self.params(_lib.BUF_ONLINE), self.params(_lib.BUF_TARGET)] + self.optimizer_state(),
                                process_group, src)

    def replica_spread(self, dist, process_group=None):
        """max |p_rank - p_rank0| over every parameter and rank (0.0 when the replicas are in sync)."""
        from .... import distributed as mdist
        return mdist.replica_spread(dist, self.params(_lib.BUF_ONLINE), process_group)

    def _write_support(self, values):
        self._buffer(_lib.BUF_SUPPORT, 128)[: values.numel()].copy_(values.to(self.device))

    # ------------------------------------------------------------------ state dicts (agent{i}.pth compatible)
    def state_dict(self, agent_id, which=_lib.BUF_ONLINE, shapes=None):
        flat = self.params(which)[agent_id].detach().cpu()
        shapes = shapes or self._shapes
        return {name: flat[off:off + cnt].clone().view(shapes[name]) for name, off, cnt in self.table}

    def load_state_dict(self, agent_id, state_dict, which=_lib.BUF_ONLINE):
        """abstract_loop_handler.py:168-180: copy the matching keys."""
        flat = self.params(which)[agent_id]
        self._shapes = getattr(self, "_shapes", {})
        for name, off, cnt in self.table:
            if name in state_dict:
                t = state_dict[name].detach().to(torch.float32)
                assert t.numel() == cnt, (name, tuple(t.shape), cnt)
                self._shapes[name] = tuple(t.shape)
                flat[off:off + cnt].copy_(t.reshape(-1).to(self.device))

    def synchronize(self):
        """hard_update of every agent's target network (abstract_brain.py:37-38)."""
        _lib.check(self.lib.maexp_learner_sync_target(self.h, self._stream()))

    # ------------------------------------------------------------------ compute
    def bind_positions(self, pos, next_pos=None):
        """int8 [G, b, 2] index-space positions of the observing agents (relative / merged models)."""
        self._pos_keep = (pos, next_pos)
        _lib.check(self.lib.maexp_learner_bind_positions(self.h, C.c_void_p(pos.data_ptr()),
                                                         None if next_pos is None else C.c_void_p(next_pos.data_ptr())))

    def bind_canvas_inputs(self, batch, keys=("obj_rows", "next_obj_rows", "agents", "next_agents", "dest_rows", "next_dest_rows")):
        """state["objects"] (bit rows int32 [G,b,T,X]), state["agents"] (int8 [G,b,num_agents,2]) and state["destination"]
        (bit rows int32 [G,b,X]) of the state and the new state, for the models that read the extra canvas channels."""
        o, a, d = self.canvas_reads
        ptrs = []
        for key, need in zip(keys, (o, o, a, a, d, d)):
            t = batch.get(key) if need else None
            if need and t is None and not key.startswith("next_"):
                raise ValueError(f"this model reads the {key} canvas input (objects_channel / agents_channel / destination_channel)")
            ptrs.append(None if t is None else C.c_void_p(t.data_ptr()))
        self._canvas_keep = [batch.get(k) for k in keys]
        _lib.check(self.lib.maexp_learner_bind_canvas_inputs(self.h, *ptrs))

    def set_los(self, table):
        """Merged models under observation noise: the int8 [X,Y,V,V] line-of-sight table (numpy) the noise-free area
        channel is rebuilt from (merged_observtion_handler.py:27,41-47)."""
        import numpy as _np
        t = _np.ascontiguousarray(table, _np.int8)
        assert t.shape == (self.X, self.Y, self.V, self.V), t.shape
        _lib.check(self.lib.maexp_learner_set_los(self.h, C.c_void_p(t.ctypes.data)))

    def q_values(self, obs, use_target=False, tau_sel=None, pos=None, canvas=None):
        """obs fp32 [G, b, obs_stride] on the device -> Q [G, b, A].  IQN: tau_sel [G, Nq] (drawn if None,
        customs/iqn.py:105-106).  Relative / merged models: pos int8 [G, b, 2]; canvas: dict with obj_rows / agents /
        dest_rows for the models that read the extra canvas channels."""
        b = obs.shape[1]
        if pos is not None:
            self.bind_positions(pos.contiguous())
        if any(self.canvas_reads):
            self.bind_canvas_inputs(canvas or {})
        q = self.q if b == self.B else torch.empty((self.G, b, self.A), dtype=torch.float32, device=self.device)
        if self.kind == "iqn":
            if tau_sel is None:
                tau_sel = torch.rand((self.G, self.Z), device=self.device)
            _lib.check(self.lib.maexp_learner_q_values_iqn(self.h, C.c_void_p(obs.data_ptr()), b, int(use_target),
                                                           C.c_void_p(tau_sel.data_ptr()), C.c_void_p(q.data_ptr()),
                                                           self._stream()))
            return q
        _lib.check(self.lib.maexp_learner_q_values(self.h, C.c_void_p(obs.data_ptr()), b, int(use_target),
                                                   C.c_void_p(q.data_ptr()), self._stream()))
        return q

    def attention(self, batch, stack=0, layer=0):
        """Attention weights [G, batch, heads, T, T] of the last q_values call (forward_attn of the attention models)."""
        T = self.lib.maexp_learner_attention_tokens(self.h, stack)
        if T < 0:
            raise ValueError("no such transformer stack in this model")
        heads = int(self.config.model.num_heads)
        out = torch.empty((self.G, batch, heads, T, T), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.maexp_learner_attention(self.h, stack, layer, batch, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def grads(self, batch):
        """Forward, TD target, loss [G] and gradients (kept in the library's grad buffer)."""
        p = lambda t: C.c_void_p(t.data_ptr())
        if self.area_mask != "local":
            self.bind_positions(batch["pos"], batch["next_pos"])
        if any(self.canvas_reads):
            self.bind_canvas_inputs(batch)
        if self.kind == "iqn":  # quantile fractions: injected for parity tests, else drawn like iqn_brain.py:74,87
            d = self.device
            taus = batch["taus"] if "taus" in batch else torch.rand((self.G, self.B, self.Z), device=d)
            dash = batch["tau_dashes"] if "tau_dashes" in batch else torch.rand((self.G, self.B, self.Nd), device=d)
            sel = batch["tau_sel"] if "tau_sel" in batch else torch.rand((self.G, self.Z), device=d)
            self._taus_keep = (taus, dash, sel)
            _lib.check(self.lib.maexp_learner_grads_iqn(self.h, p(batch["obs"]), p(batch["next_obs"]), p(batch["actions"]),
                                                        p(batch["rewards"]), p(batch["dones"]), p(taus), p(dash), p(sel),
                                                        p(self.loss), self._stream()))
            return self.loss
        _lib.check(self.lib.maexp_learner_grads(self.h, p(batch["obs"]), p(batch["next_obs"]), p(batch["actions"]),
                                                p(batch["rewards"]), p(batch["dones"]), p(self.loss), self._stream()))
        return self.loss

    def apply(self, grad_scale=1.0, active=None):
        a = None if active is None else C.c_void_p(active.data_ptr())
        _lib.check(self.lib.maexp_learner_apply(self.h, float(grad_scale), a, C.c_void_p(self.grad_norm.data_ptr()),
                                                self._stream()))

    def learn(self, batch, active=None, all_reduce=None):
        """One update of every agent's network.  all_reduce: optional callable(grads_tensor) -> world size,
        e.g. an NCCL all-reduce for the data-parallel learner (clip runs AFTER the reduce)."""
        loss = self.grads(batch)
        scale = 1.0
        if all_reduce is not None:
            scale = 1.0 / float(all_reduce(self.params(_lib.BUF_GRADS)))
        self.apply(scale, active)
        return loss


class GraphedUpdate:
    """A team update on a fixed batch buffer replayed from CUDA graphs: one graph for forward / backward of every agent's
    network and one for clip + optimizer step; the data-parallel learner's NCCL all-reduce of the flat gradient buffer runs
    between them (capturing it into the graph was measured no faster at 2 GPUs, 0.290 vs 0.289 ms per update, and a captured
    collective has to be destroyed before the process group).  The first step() runs eagerly (it is a real update and
    initialises lazily created state such as the NCCL communicator), the second captures, later ones replay.

    dist: torch.distributed module (or None for a single GPU); the replay sampler refills `batch` in place before step()."""

    def __init__(self, group, batch, active=None, dist=None, process_group=None, fused_collective=True):
        self.group, self.batch, self.active, self.dist, self.pg = group, batch, active, dist, process_group
        self.world = dist.get_world_size(process_group) if dist is not None else 1
        self.g_grads = self.g_apply = None
        self.single_graph = False
        self.calls = 0
        # data parallel on NVLink peers: the gradient buffer moves into symmetric memory and ONE cooperative kernel does the
        # exchange and the optimizer step (maexp_learner_apply_data_parallel); otherwise NCCL all-reduce + apply
        self.peers = None
        if dist is not None and self.world > 1 and fused_collective and group.config.opt.name == "rmsprop" and \
                dist.get_backend(process_group) == "nccl":
            try:
                self.peers = self._symmetric_gradients()
            except Exception as exc:  # no peer access / symmetric memory: keep the NCCL path
                logger.warning(f"fused data-parallel apply unavailable ({exc}); using NCCL all-reduce")
        self.grads_view = group.params(_lib.BUF_GRADS)
        self.draw_taus = False
        if group.kind == "iqn" and not all(k in batch for k in ("taus", "tau_dashes", "tau_sel")):
            # the fractions live in fixed buffers refilled before every replay (a captured torch.rand would replay one
            # draw): one Philox launch over the three tensors (iqn_brain.py:74,87, customs/iqn.py:105-106)
            g = group
            self.tau_flat = torch.empty((g.G * g.B * (g.Z + g.Nd) + g.G * g.Z,), dtype=torch.float32, device=g.device)
            n1, n2 = g.G * g.B * g.Z, g.G * g.B * g.Nd
            batch["taus"] = self.tau_flat[:n1].view(g.G, g.B, g.Z)
            batch["tau_dashes"] = self.tau_flat[n1:n1 + n2].view(g.G, g.B, g.Nd)
            batch["tau_sel"] = self.tau_flat[n1 + n2:].view(g.G, g.Z)
            self.draw_taus = True
            self.tau_seed = int(group.config.seed) & 0xFFFFFFFF
            if dist is not None:  # every rank samples its own shard of the global batch: its own fractions, too
                self.tau_seed = (self.tau_seed + 0x9E3779B9 * dist.get_rank(process_group)) & 0xFFFFFFFF

    def _symmetric_gradients(self):
        import torch.distributed._symmetric_memory as symm
        g, dist = self.group, self.dist
        pg = self.pg if self.pg is not None else dist.group.WORLD
        name = pg.group_name
        if hasattr(symm, "enable_symm_mem_for_group"):
            try:
                symm.enable_symm_mem_for_group(name)
            except Exception:
                pass
        n, W = g.G * g.P, self.world
        n_part = (W * g.G + 3) // 4 * 4
        self.symm_buf = symm.empty(n + n_part + 64, dtype=torch.float32, device=g.device)
        self.symm_buf.zero_()
        self.symm_hdl = symm.rendezvous(self.symm_buf, name)
        torch.cuda.synchronize(g.device)
        dist.barrier(group=self.pg)  # every rank's buffer is zero before the first signal can arrive
        peers = _lib.DpPeers()
        peers.world_size, peers.rank = W, dist.get_rank(self.pg)
        for q, base in enumerate(self.symm_hdl.buffer_ptrs):
            peers.grads[q], peers.partials[q], peers.signals[q] = base, base + 4 * n, base + 4 * (n + n_part)
        assert peers.grads[peers.rank] == self.symm_buf.data_ptr()
        mc = int(getattr(self.symm_hdl, "multicast_ptr", 0) or 0)
        # NVSwitch: in-switch reduction / broadcast (NVLS multimem).  Measured on 8 x B200: 38.7 us per exchange + apply
        # against 41.5 us with peer loads / stores (NCCL all-reduce + apply: 63 us); at 2 GPUs peer accesses win (32.9 vs 40.1)
        want = os.environ.get("MAEXP_DP_MULTIMEM")
        if mc and (want == "1" or (want is None and W > 2)):
            peers.multicast_grads = mc
        _lib.check(g.lib.maexp_learner_bind_grads(g.h, C.c_void_p(self.symm_buf.data_ptr())))
        return peers

    def _reduce(self):
        if self.dist is not None:
            self.dist.all_reduce(self.grads_view, op=self.dist.ReduceOp.SUM, group=self.pg)

    def _apply(self):
        """all ranks: exchange + optimizer step"""
        g = self.group
        if self.peers is not None:
            a = None if self.active is None else C.c_void_p(self.active.data_ptr())
            _lib.check(g.lib.maexp_learner_apply_data_parallel(g.h, C.byref(self.peers), a, C.c_void_p(g.grad_norm.data_ptr()),
                                                               g._stream()))
            return True
        return False

    def step(self):
        self.calls += 1
        if self.draw_taus:
            _lib.check(self.group.lib.maexp_fill_uniform(C.c_void_p(self.tau_flat.data_ptr()), self.tau_flat.numel(),
                                                         self.tau_seed, self.calls & 0xFFFFFFFF, self.group._stream()))
        if self.calls == 1:
            loss = self.group.grads(self.batch)
            if not self._apply():
                self._reduce()
                self.group.apply(1.0 / self.world, self.active)
            return loss
        if self.g_grads is None:
            torch.cuda.synchronize(self.group.device)
            self.g_grads = torch.cuda.CUDAGraph()
            # one GPU: nothing runs between the two halves, so they are ONE graph (no graph boundary before the optimizer)
            self.single_graph = self.dist is None and os.environ.get("MAEXP_SINGLE_GRAPH", "1") != "0"
            with torch.cuda.graph(self.g_grads):
                self.group.grads(self.batch)
                if self.single_graph:
                    self.group.apply(1.0, self.active)
            if self.peers is None and not self.single_graph:
                self.g_apply = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g_apply):
                    self.group.apply(1.0 / self.world, self.active)
        self.g_grads.replay()
        if self.single_graph:
            return self.group.loss
        if not self._apply():  # fused: one cooperative launch (exchange over NVLink + clip + RMSprop)
            self._reduce()
            self.g_apply.replay()
        return self.group.loss


class Brain:
    """Single-agent brain with the reference's interface (abstract_brain.py:20-67)."""

    def __init__(self, config, obs_shape, act_size):
        self.config = config
        self.device = torch.device("cuda:0")
        self.group = LearnerGroup(config, obs_shape, act_size, 1, int(config.batch_size), device=self.device)
        self.view = "local" if config.observation_area_mask in ("local", "merged") else "relative"
        if config.observation_area_mask == "merged" and config.observation_noise:
            self.group.set_los(los_table(config))  # the noise-free "relative" key's area channel

    def _pack(self, x):
        g = self.group
        x = x.float().reshape(x.shape[0], -1).to(self.device)
        out = torch.zeros((1, x.shape[0], g.obs_stride), dtype=torch.float32, device=self.device)
        out[0, :, : g.K] = x
        return out

    def _positions(self, state, batched):
        """Observing agent's index-space cell from the `coordinates` dict (local_area_mask.py:18-47)."""
        c = state["coordinates"]
        r = int(self.config.visible_range) // 2
        px = torch.as_tensor(c["map_x_min"]) - torch.as_tensor(c["obs_x_min"]) + r
        py = torch.as_tensor(c["map_y_min"]) - torch.as_tensor(c["obs_y_min"]) + r
        pos = torch.stack([px.reshape(-1), py.reshape(-1)], dim=1).to(torch.int8)
        return pos.reshape(1, -1, 2).contiguous().to(self.device)

    @staticmethod
    def _bit_rows(t):
        """[..., X, Y] 0/1 matrices -> int32 [..., X] bit rows (bit y of word x), the library's layout."""
        t = (torch.as_tensor(t) != 0).to(torch.int64)
        weights = (torch.ones((), dtype=torch.int64) << torch.arange(t.shape[-1], dtype=torch.int64))
        return (t * weights).sum(dim=-1).to(torch.int32)

    def _canvas(self, state, prefix=""):
        """The reference's "objects" / "agents" / "destination" state keys (observation_handler.py:95-105) in the
        library's layout, for the models that concatenate them onto the relative canvas."""
        o, a, d = self.group.canvas_reads
        out = {}
        if o:
            x = torch.as_tensor(state["objects"])
            x = x.reshape(-1, *x.shape[-3:])
            out[prefix + "obj_rows"] = self._bit_rows(x).reshape(1, x.shape[0], -1).contiguous().to(self.device)
        if a:
            x = torch.as_tensor(state["agents"]).to(torch.int8)
            out[prefix + "agents"] = x.reshape(1, -1, x.shape[-2], 2).contiguous().to(self.device)
        if d:
            x = torch.as_tensor(state["destination"])
            x = x.reshape(-1, *x.shape[-2:])
            out[prefix + "dest_rows"] = self._bit_rows(x).reshape(1, x.shape[0], -1).contiguous().to(self.device)
        return out

    @torch.no_grad()
    def get_action(self, state):
        """int for the plain brains (abstract_brain.py:52-64); (action, attns) for the attention brains:
        da3 / da3_iqn return forward_attn's per-block attention weights [1, heads, T, T] — [attns] on the local view,
        [local_attns, relative_attns] for the merged models and DA6 (da3_brain.py:47-56, customs/da3.py:167-176,
        customs/da6.py:128-165) — and da3_baseline a zero placeholder (da3_brain.py:93-123)."""
        pos = self._positions(state, False) if self.group.area_mask != "local" else None
        q = self.group.q_values(self._pack(state[self.view].unsqueeze(0)), pos=pos, canvas=self._canvas(state))
        action = int(q[0, 0].argmax().item())
        brain = self.config.brain
        if brain == "da3_baseline":
            t = int(self.config.visible_range) ** 2 + 1
            heads = int(self.config.model.get("num_heads") or 1)
            return action, [[torch.zeros((int(self.config.batch_size), heads, t, t), device=self.device)]]
        if brain in ("da3", "da6", "da3_iqn"):
            g, blocks = self.group, int(self.config.model.block_loop)
            local = [g.attention(1, stack=0, layer=l)[0] for l in range(blocks)]
            if g.lib.maexp_learner_attention_tokens(g.h, 1) > 0:  # DA6 / DA6_IQN: the relative transformer
                return action, [local, [g.attention(1, stack=1, layer=l)[0] for l in range(blocks)]]
            if g.area_mask == "merged":  # MergedDA3 returns the local maps twice (customs/da3.py:199-201)
                return action, [local, local]
            return action, [local]
        return action

    def learn(self, states_ind, actions_ind, rewards_ind, dones_ind, next_states_ind):
        batch = {"obs": self._pack(states_ind[self.view]), "next_obs": self._pack(next_states_ind[self.view]),
                 "actions": actions_ind.to(self.device).to(torch.int32).reshape(1, -1).contiguous(),
                 "rewards": rewards_ind.float().to(self.device).reshape(1, -1).contiguous(),
                 "dones": dones_ind.to(self.device).to(torch.uint8).reshape(1, -1).contiguous()}
        if self.group.area_mask != "local":
            batch["pos"], batch["next_pos"] = self._positions(states_ind, True), self._positions(next_states_ind, True)
            batch.update(self._canvas(states_ind))
            batch.update(self._canvas(next_states_ind, "next_"))
        loss = self.group.learn(batch)
        return {"total_loss": loss[0].detach().cpu()}  # always a dict

    def synchronize_network(self):
        self.group.synchronize()

    @property
    def network(self):
        return _StateDictView(self.group, _lib.BUF_ONLINE)

    @property
    def target_network(self):
        return _StateDictView(self.group, _lib.BUF_TARGET)


class _StateDictView:
    def __init__(self, group, which):
        self.group, self.which = group, which

    def state_dict(self):
        return self.group.state_dict(0, self.which)

    def load_state_dict(self, sd):
        self.group.load_state_dict(0, sd, self.which)


def generate_brain(config, obs_shape, act_size):
    """core/agents/brains/__init__.py:42-85 (dqn, da3_baseline, categorical_dqn, qr_dqn)."""
    learner_kind(config)  # raises ValueError on unknown keys, as the reference does
    return Brain(config=config, obs_shape=obs_shape, act_size=act_size)
