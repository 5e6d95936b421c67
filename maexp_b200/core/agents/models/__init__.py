"""Network *definitions* (host side): layer shapes, state_dict keys and default initialisation.

The forward/backward arithmetic runs in the CUDA library (csrc/learner.cu); these torch modules
only exist so that parameters are created with PyTorch's default initialisers in the reference's
construction order and under the reference's state_dict keys
(core/agents/models/convolution.py:14-55, mlp.py:14-34, customs/conv_mlp.py:21-52,
customs/categorical_dqn.py:28-58, customs/qr_dqn.py:20-44), so `agent{i}.pth` files interchange.
"""
import torch
from torch import nn


class _Conv(nn.Module):
    def __init__(self, config, input_channel, output_channel):
        super().__init__()
        m = config.model
        self.conv_list = nn.ModuleList()
        for ch, k, s in zip(m.hidden_channels, m.hidden_kernels, m.hidden_strides):
            self.conv_list.append(nn.Conv2d(input_channel, ch, kernel_size=k, stride=s))
            input_channel = ch
        self.conv_post = nn.Conv2d(input_channel, output_channel, kernel_size=m.hidden_kernels[-1],
                                   stride=m.hidden_strides[-1])


class _MLP(nn.Module):
    def __init__(self, config, input_size, output_size):
        super().__init__()
        self.fc_list = nn.ModuleList()
        for size in config.model.hidden_layer_sizes:
            self.fc_list.append(nn.Linear(input_size, size))
            input_size = size
        self.fc_post = nn.Linear(input_size, output_size)


def extra_channels(config, family):
    """in_chans the reference's constructors add for the canvas inputs they concatenate on the relative view: ConvMLP and
    DA3 read objects_channel (conv_mlp.py:25-27, da3.py:31-33), IQN objects + agents (iqn.py:55-60), DA3_IQN all three
    (da3_iqn.py:96-104); DA6 / DA6_IQN put objects + agents on their relative stack (da6.py:32-37, da6_iqn.py:42-47)."""
    t = int(config.type_objects) if config.get("objects_channel") else 0
    n = int(config.num_agents) if config.get("agents_channel") else 0
    d = 1 if config.get("destination_channel") else 0
    return {"conv_mlp": t, "da3": t, "iqn": t + n, "da6": t + n, "da3_iqn": t + n + d}[family]


def conv_output_size(config, obs_shape):
    """Flattened size of Conv's output (conv k2 s1 -> max_pool2d(2,2) -> conv k2 s1)."""
    h2, w2 = (int(obs_shape[1]) - 1) // 2 - 1, (int(obs_shape[2]) - 1) // 2 - 1
    return int(config.model.output_channel) * h2 * w2


class _ConvHeads(nn.Module):
    def __init__(self, config, obs_shape, heads, extra=0):
        super().__init__()
        self.conv = _Conv(config, int(obs_shape[0]) + extra, config.model.output_channel)
        torch.randn(1, int(obs_shape[0]) + extra, *obs_shape[1:])  # the reference probes the conv with a random input here (conv_mlp.py:51)
        d = conv_output_size(config, obs_shape)
        for name, out in heads:
            setattr(self, name, _MLP(config, d, out))


class _IQN(nn.Module):
    """Parameter layout and construction order of customs/iqn.py:35-67."""

    def __init__(self, config, obs_shape, act_size):
        super().__init__()
        x = extra_channels(config, "iqn")
        self.conv = _Conv(config, int(obs_shape[0]) + x, config.model.output_channel)
        torch.randn(1, int(obs_shape[0]) + x, *obs_shape[1:])  # customs/iqn.py:143
        d = conv_output_size(config, obs_shape)
        self.state_embedder = _MLP(config, d, d)
        self.cosine_net = nn.Module()
        self.cosine_net.net = nn.Sequential(nn.Linear(int(config.model.num_cosines), d), nn.ReLU(inplace=False))
        self.fc_V = _MLP(config, d, 1)
        self.fc_A = _MLP(config, d, act_size)


class _MergedIQN(_IQN):
    """customs/iqn.py:165-201: IQN's conv / state_embedder stay as unused parameters; the heads are rebuilt on the
    concatenated local + relative embedding."""

    def __init__(self, config, obs_shape, act_size):
        super().__init__(config, obs_shape, act_size)
        v = int(config.visible_range)
        self.local_conv = _Conv(config, obs_shape[0], config.model.output_channel)
        self.relative_conv = _Conv(config, 1, config.model.output_channel)
        torch.randn(1, obs_shape[0], v, v)
        torch.randn(1, 1, *obs_shape[1:])
        dl, dr = conv_output_size(config, [obs_shape[0], v, v]), conv_output_size(config, [1, *obs_shape[1:]])
        self.local_state_embedder = _MLP(config, dl, dl)
        self.relative_state_embedder = _MLP(config, dr, dr)
        self.cosine_net = nn.Module()
        self.cosine_net.net = nn.Sequential(nn.Linear(int(config.model.num_cosines), dl + dr), nn.ReLU(inplace=False))
        self.fc_V = _MLP(config, dl + dr, 1)
        self.fc_A = _MLP(config, dl + dr, act_size)


class _MergedConvMLP(nn.Module):
    """customs/conv_mlp.py:69-98: the parent's `conv` stays as unused parameters; `mlp` is replaced."""

    def __init__(self, config, obs_shape, act_size):
        super().__init__()
        v = int(config.visible_range)
        x = extra_channels(config, "conv_mlp")
        self.conv = _Conv(config, int(obs_shape[0]) + x, config.model.output_channel)
        torch.randn(1, int(obs_shape[0]) + x, *obs_shape[1:])
        self.mlp = _MLP(config, conv_output_size(config, obs_shape), act_size)  # replaced below, as in the reference
        self.local_conv = _Conv(config, obs_shape[0], config.model.output_channel)
        self.relative_conv = _Conv(config, 1, config.model.output_channel)
        torch.randn(1, obs_shape[0], v, v)
        torch.randn(1, 1, *obs_shape[1:])
        d = conv_output_size(config, [obs_shape[0], v, v]) + conv_output_size(config, [1, *obs_shape[1:]])
        self.mlp = _MLP(config, d, act_size)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=False)  # vit.py:111 (qkv_bias=False)
        self.proj = nn.Linear(dim, dim)


class _VitMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    """vit.py:50-80 parameter layout."""

    def __init__(self, dim, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _VitMlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class _DA3(nn.Module):
    """Parameter layout and construction order of customs/da3.py:22-60."""

    FAMILY = "da3"

    def __init__(self, config, obs_shape, act_size):
        super().__init__()
        m = config.model
        tokens = (int(obs_shape[1]) // int(m.patch_size)) * (int(obs_shape[2]) // int(m.patch_size)) + 1
        self.patch_embed = _PatchEmbed(int(m.patch_size), int(obs_shape[0]) + extra_channels(config, self.FAMILY), int(m.embed_dim))
        self.saliency_vector = nn.Parameter(torch.zeros(1, 1, int(m.embed_dim)))
        self.pos_embed = nn.Parameter(torch.zeros(1, tokens, int(m.embed_dim)))
        self.blocks = nn.ModuleList([_Block(int(m.embed_dim), float(m.mlp_ratio)) for _ in range(int(m.block_loop))])
        self.norm = nn.LayerNorm(int(m.embed_dim))
        self.drl_head = nn.Linear(int(m.embed_dim), act_size)


class _MergedDA3(_DA3):
    """customs/da3.py:99-165: parent parameters (unused) + local ViT + relative conv / embedder + wide head."""

    def __init__(self, config, obs_shape, act_size):
        super().__init__(config, obs_shape, act_size)
        m, v = config.model, int(config.visible_range)
        self.local_patch_embed = _PatchEmbed(1, int(obs_shape[0]), int(m.embed_dim))
        self.local_saliency_vector = nn.Parameter(torch.zeros(1, 1, int(m.embed_dim)))
        self.local_pos_embed = nn.Parameter(torch.zeros(1, v * v + 1, int(m.embed_dim)))
        self.local_blocks = nn.ModuleList([_Block(int(m.embed_dim), float(m.mlp_ratio)) for _ in range(int(m.block_loop))])
        self.norm = nn.LayerNorm(int(m.embed_dim))
        self.relative_conv = _Conv(config, 1, m.output_channel)
        torch.randn(1, 1, *obs_shape[1:])
        d = conv_output_size(config, [1, *obs_shape[1:]])
        self.relative_state_embedder = _MLP(config, d, d)
        self.drl_head = nn.Linear(int(m.embed_dim) + d, act_size)


class _IQNHead(nn.Module):
    """customs/da3_iqn.py:26-40"""

    def __init__(self, config, d, act_size):
        super().__init__()
        self.cosine_net = nn.Module()
        self.cosine_net.net = nn.Sequential(nn.Linear(int(config.model.num_cosines), d), nn.ReLU(inplace=False))
        self.fc_V = _MLP(config, d, 1)
        self.fc_A = _MLP(config, d, act_size)


class _DA3IQN(_DA3):
    """customs/da3_iqn.py:87-145: DA3's transformer with the IQN head."""
    FAMILY = "da3_iqn"

    def __init__(self, config, obs_shape, act_size):
        super().__init__(config, obs_shape, act_size)
        self.drl_head = _IQNHead(config, int(config.model.embed_dim), act_size)


class _MergedDA3IQN(_MergedDA3):
    """customs/da3_iqn.py:213-265"""
    FAMILY = "da3_iqn"

    def __init__(self, config, obs_shape, act_size):
        super().__init__(config, obs_shape, act_size)
        d = conv_output_size(config, [1, *obs_shape[1:]])
        self.drl_head = _IQNHead(config, int(config.model.embed_dim) + d, act_size)


class _DA6IQN(_DA3):
    """customs/da6_iqn.py:24-96: DA3_IQN's (unused) parameters plus the two DA6 stages; patch sizes 5 and 1 are hard-coded."""
    FAMILY = "da3_iqn"

    def __init__(self, config, obs_shape, act_size):
        super().__init__(config, obs_shape, act_size)
        m, v, e = config.model, int(config.visible_range), int(config.model.embed_dim)
        self.relative_patch_embed = _PatchEmbed(5, 1 + extra_channels(config, "da6"), e)
        self.local_patch_embed = _PatchEmbed(1, int(obs_shape[0]), e)
        self.saliency_vector = nn.Parameter(torch.zeros(1, 1, e))
        self.relative_pos_embed = nn.Parameter(torch.zeros(1, (int(obs_shape[1]) // 5) * (int(obs_shape[2]) // 5) + 1, e))
        self.local_pos_embed = nn.Parameter(torch.zeros(1, v * v + 1, e))
        self.relative_blocks = nn.ModuleList([_Block(e, float(m.mlp_ratio)) for _ in range(int(m.block_loop))])
        self.local_blocks = nn.ModuleList([_Block(e, float(m.mlp_ratio)) for _ in range(int(m.block_loop))])
        self.norm = nn.LayerNorm(e)
        self.fc1 = nn.Linear(e, e)
        self.drl_head = _IQNHead(config, e, act_size)


class _DA6(nn.Module):
    """Parameter layout and construction order of customs/da6.py:22-96."""

    def __init__(self, config, obs_shape, act_size):
        super().__init__()
        m, v, e = config.model, int(config.visible_range), int(config.model.embed_dim)
        rp, lp = int(m.relative_patch_size), int(m.local_patch_size)
        self.relative_patch_embed = _PatchEmbed(rp, 1 + extra_channels(config, "da6"), e)
        self.local_patch_embed = _PatchEmbed(lp, int(obs_shape[0]), e)
        self.saliency_vector = nn.Parameter(torch.zeros(1, 1, e))
        self.relative_pos_embed = nn.Parameter(torch.zeros(1, (int(obs_shape[1]) // rp) * (int(obs_shape[2]) // rp) + 1, e))
        self.local_pos_embed = nn.Parameter(torch.zeros(1, (v // lp) * (v // lp) + 1, e))
        self.relative_blocks = nn.ModuleList([_Block(e, float(m.mlp_ratio)) for _ in range(int(m.block_loop))])
        self.local_blocks = nn.ModuleList([_Block(e, float(m.mlp_ratio)) for _ in range(int(m.block_loop))])
        self.norm = nn.LayerNorm(e)
        self.fc1 = nn.Linear(e, e)
        self.drl_head = nn.Linear(e, act_size)


def build_reference_modules(config, obs_shape, act_size):
    name = config.model.name
    if name == "mlp":  # the flattened observation feeds the MLP
        return _MLP(config, int(obs_shape[0]) * int(obs_shape[1]) * int(obs_shape[2]), act_size)
    if name in ("conv_mlp", "da3_baseline"):
        if config.observation_area_mask == "merged":
            return _MergedConvMLP(config, obs_shape, act_size)
        return _ConvHeads(config, obs_shape, [("mlp", act_size)], extra=extra_channels(config, "conv_mlp"))
    if name == "categorical_dqn":
        z = int(config.model.num_atoms)
        return _ConvHeads(config, obs_shape, [("fc_V", z), ("fc_A", act_size * z)])
    if name == "qr_dqn":
        z = int(config.model.num_quantiles)
        return _ConvHeads(config, obs_shape, [("fc_V", z), ("fc_A", act_size * z)])
    if name == "iqn":
        if config.observation_area_mask == "merged":
            return _MergedIQN(config, obs_shape, act_size)
        return _IQN(config, obs_shape, act_size)
    if name == "da6":
        return _DA6(config, obs_shape, act_size)
    if name == "da6_iqn":
        return _DA6IQN(config, obs_shape, act_size)
    if name == "da3_iqn":
        if config.observation_area_mask == "merged":
            return _MergedDA3IQN(config, obs_shape, act_size)
        return _DA3IQN(config, obs_shape, act_size)
    if name == "da3":
        if config.observation_area_mask == "merged":
            return _MergedDA3(config, obs_shape, act_size)
        return _DA3(config, obs_shape, act_size)
    raise ValueError(name)
