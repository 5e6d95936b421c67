"""Attribute-dict configuration with the reference's own key names.

The reference composes its configuration with Hydra/omegaconf (main.py:31,
configs/experiments/*.yaml).  Neither package exists in this image, and none of the five
BASELINE.json configurations is shipped as a YAML file, so the
build authors them here with exactly the keys the hot path reads.
"""
import copy


class Config(dict):
    """Minimal DictConfig look-alike: attribute get/set, nested dicts, ``dict(cfg)``."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = Config(v) if isinstance(v, dict) and not isinstance(v, Config) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return Config(copy.deepcopy(dict(self)))

    def merged_with(self, **kw):
        c = self.copy()
        for k, v in kw.items():
            c[k] = Config(v) if isinstance(v, dict) else v
        return c


# Free cells (centred coordinates, `init_xys`) that are off every wall of every supported map.
_INIT_XYS = [[4, 4], [-4, 4], [4, -4], [-4, -4], [6, 6], [-6, 6], [6, -6], [-6, -6],
             [4, 6], [-4, 6], [4, -6], [-4, -6]]

_DEFAULTS = dict(
    seed=921, gpu=0, world="default", environment="default", phase="training",
    map=dict(name="four_rooms", SIZE_X=24, SIZE_Y=24),
    num_agents=8, num_objects=25, type_objects=1, keep_objects_num=True,
    agent_tasks=["0"] * 8, init_xys=_INIT_XYS[:8], shuffle_init_xys=True,
    agents_color=["blue"] * 8, objects_color=["yellow"],
    agents_channel=False, objects_channel=False, destination_channel=False, reset_destination_period=0,
    visible_range=7, transparent_observation=False,
    agent_view_method="default", object_view_method="default", observation_area_mask="local",
    observation_noise=False, noise_threshold_ratio=1.0, flip_noise_probabilities=[0.2, 0.1, 0.05],
    std_distribution=[0.3, 0.2, 0.1], past_step=0,
    max_episode_length=200, max_epochs=5000, batch_size=32, populate_steps=1000, capacity=100000,
    buffer="default", synchronize_frequency=5, gamma=0.9,
    epsilon_initial=1.0, epsilon_end=0.05, epsilon_decay=0.999,
    agent_type="default", brain="da3_baseline", trainer="default", evaluator="default",
    validate_epochs=1000, pretrained_weight_path=False,
    model=dict(name="conv_mlp", hidden_channels=[32], output_channel=64, hidden_kernels=[2, 2],
               hidden_strides=[1, 1], conv_afs=["max_pool2d", "max_pool2d"], af_kernels=[2, 2],
               af_strides=[2, 2], hidden_layer_sizes=[128], fc_afs=["relu"]),
    opt=dict(name="rmsprop", learning_rate=1e-3, alpha=0.99, eps=1e-8),
)


def make_config(**overrides):
    """Default (north-star shaped) configuration with keyword overrides.

    Per-agent list keys are resized to ``num_agents`` when they are not given explicitly."""
    cfg = Config(copy.deepcopy(_DEFAULTS))
    for k, v in overrides.items():
        cfg[k] = Config(v) if isinstance(v, dict) else v
    n = cfg.num_agents
    if "agent_tasks" not in overrides:
        cfg.agent_tasks = ["0"] * n if cfg.type_objects == 1 else [str(i % cfg.type_objects) for i in range(n)]
    if "init_xys" not in overrides:
        cfg.init_xys = copy.deepcopy(_INIT_XYS[:n])
    if "agents_color" not in overrides:
        cfg.agents_color = ["blue"] * n
    if "objects_color" not in overrides:
        cfg.objects_color = ["yellow"] * cfg.type_objects
    return cfg


def baseline_config(index):
    """The five configurations BASELINE.json names."""
    if index == 1:
        return make_config(map=dict(name="simple", SIZE_X=20, SIZE_Y=20), num_agents=2, visible_range=5,
                           brain="dqn", model=dict(name="mlp", hidden_layer_sizes=[120], fc_afs=["relu"]))
    if index == 2:
        return make_config(map=dict(name="four_rooms", SIZE_X=24, SIZE_Y=24), num_agents=4,
                           agent_view_method="transition", past_step=4, observation_noise="flip",
                           environment="shared_mean_reward", brain="da3_baseline")
    if index == 3:
        return make_config(map=dict(name="three_rooms", SIZE_X=25, SIZE_Y=25), num_agents=8, batch_size=256,
                           brain="qr_dqn",
                           model=dict(name="qr_dqn", hidden_channels=[32], output_channel=64, hidden_kernels=[2, 2],
                                      hidden_strides=[1, 1], conv_afs=["max_pool2d", "max_pool2d"],
                                      af_kernels=[2, 2], af_strides=[2, 2], hidden_layer_sizes=[128],
                                      fc_afs=["relu"], num_quantiles=100))
    if index == 4:
        return make_config(map=dict(name="four_rectangle", SIZE_X=40, SIZE_Y=24), num_agents=8, type_objects=2,
                           agent_view_method="individual", object_view_method="object_types",
                           observation_area_mask="merged", agent_type="attention", brain="da3",
                           model=dict(name="da3", embed_dim=64, patch_size=1, relative_patch_size=5,
                                      local_patch_size=1, block_loop=1, num_heads=4, mlp_ratio=4.0,
                                      hidden_channels=[32], output_channel=64, hidden_kernels=[2, 2],
                                      hidden_strides=[1, 1], conv_afs=["max_pool2d", "max_pool2d"],
                                      af_kernels=[2, 2], af_strides=[2, 2], hidden_layer_sizes=[128],
                                      fc_afs=["relu"]))
    if index == 5:
        return make_config(map=dict(name="four_rooms", SIZE_X=24, SIZE_Y=24), num_agents=8, batch_size=256,
                           brain="categorical_dqn",
                           model=dict(name="categorical_dqn", hidden_channels=[32], output_channel=64,
                                      hidden_kernels=[2, 2], hidden_strides=[1, 1],
                                      conv_afs=["max_pool2d", "max_pool2d"], af_kernels=[2, 2], af_strides=[2, 2],
                                      hidden_layer_sizes=[128], fc_afs=["relu"], num_atoms=51, V_min=-10.0,
                                      V_max=10.0))
    raise ValueError(f"no BASELINE config {index}")
