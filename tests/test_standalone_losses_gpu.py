"""Stand-alone loss / optimizer entry points (maexp_loss_dqn / _c51 / _qr / _iqn, maexp_clip_rmsprop_step) on caller-owned
network outputs: a maintainer's own torch network (here the torch restatement with the live-reference fixture's weights) feeds
its outputs to the CUDA loss kernel, back-propagates the returned output gradient and hands the flat parameter gradient to the
fused clip + RMSprop kernel.  Two such updates must reproduce what the LIVE reference's brain.learn did on the same batch:
its losses (tests/golden/learner_*.npz `losses`, 1e-5 relative) and its weights after the second update (`final`).
References: da3_brain.py:69-87, categorical_dqn_brain.py:51-87, qr_dqn_brain.py:15-51,91-117, iqn_brain.py:15-50,73-113,
optimizers/__init__.py:28-33."""
import ctypes as C

import pytest
import torch

from tests.test_learner_golden import IQN_KINDS, load, make_oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def ptr(t):
    return C.c_void_p(t.data_ptr())


def flat(tensors):
    return torch.cat([t.reshape(-1) for t in tensors])


@pytest.mark.parametrize("kind", ["dqn_conv_mlp", "c51", "qr", "iqn", "da3"])
def test_standalone_loss_and_optimizer_reproduce_the_reference_update(kind):
    from maexp_b200 import _lib
    lib = _lib.load()
    fix = load(kind)
    o, b = make_oracle(kind, fix), fix["batch"]
    B, A = b["obs"].shape[0], 4
    names = list(o.sd)
    params = flat([o.sd[k].detach() for k in names]).cuda().view(1, -1).contiguous()
    square_avg = torch.zeros_like(params)
    actions = b["actions"].to(torch.int32).view(1, B).contiguous().cuda()
    rewards = b["rewards"].float().view(1, B).contiguous().cuda()
    dones = b["dones"].to(torch.uint8).view(1, B).contiguous().cuda()
    loss = torch.zeros(1, device="cuda")
    norm = torch.zeros(1, device="cuda")
    for u, want in enumerate(fix["losses"]):
        if kind in IQN_KINDS:
            taus, dash, sel = fix["taus"][u]
            out_s = o.quantiles(o.sd, b["obs"], taus)                                 # [B, Nq, A]
            with torch.no_grad():
                out_n, out_t = o.quantiles(o.sd, b["next_obs"], sel), o.quantiles(o.target, b["next_obs"], dash)
        else:
            out_s = o.net(o.sd, b["obs"])                                             # [B, A] or [B, A, Z]
            with torch.no_grad():
                out_n, out_t = o.net(o.sd, b["next_obs"]), o.net(o.target, b["next_obs"])
        dev = [t.detach().contiguous().cuda() for t in (out_s, out_n, out_t)]
        d_out = torch.full_like(dev[0], float("nan"))
        if kind in ("dqn_conv_mlp", "da3"):
            _lib.check(lib.maexp_loss_dqn(ptr(dev[0]), ptr(dev[2]), ptr(actions), ptr(rewards), ptr(dones), fix["gamma"], 1, B, A,
                                          ptr(loss), ptr(d_out), None))
        elif kind == "c51":
            support = o.support.cuda()
            _lib.check(lib.maexp_loss_c51(ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(actions), ptr(rewards), ptr(support),
                                          fix["gamma"], o.v_min, o.v_max, o.delta_z, 1, B, A, o.Z, ptr(loss), ptr(d_out), None))
        elif kind == "qr":
            tau_hats = o.tau_hats.view(-1).contiguous().cuda()
            _lib.check(lib.maexp_loss_qr(ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(actions), ptr(rewards), ptr(tau_hats),
                                         fix["gamma"], 1, B, A, o.Z, ptr(loss), ptr(d_out), None))
        else:
            t = taus.view(1, B, -1).contiguous().cuda()
            _lib.check(lib.maexp_loss_iqn(ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(t), ptr(actions), ptr(rewards),
                                          fix["gamma"], 1, B, A, taus.shape[1], dash.shape[1], ptr(loss), ptr(d_out), None))
        torch.cuda.synchronize()
        assert abs(float(loss) - want) <= RTOL * abs(want), (u, float(loss), want)
        # the output gradient against autograd through the restated loss
        o.opt.zero_grad()
        o.taus = fix["taus"][u] if kind in IQN_KINDS else None
        ref_loss = o.loss(b["obs"], b["next_obs"], b["actions"], b["rewards"], b["dones"])
        ref_loss.backward()
        ref_grads = flat([o.sd[k].grad if o.sd[k].grad is not None else torch.zeros_like(o.sd[k]) for k in names])
        o.opt.zero_grad()
        assert not torch.isnan(d_out).any()
        out_s.backward(d_out.cpu())
        grads = flat([o.sd[k].grad if o.sd[k].grad is not None else torch.zeros_like(o.sd[k]) for k in names])
        assert float((grads - ref_grads).norm()) <= 1e-5 * float(ref_grads.norm())
        # clip_grad_norm_ + RMSprop on the flat buffers
        g = grads.cuda().view(1, -1).contiguous()
        _lib.check(lib.maexp_clip_rmsprop_step(ptr(params), ptr(g), ptr(square_avg), params.shape[1], 1, 1.0, o.max_norm, 1e-3, 0.99,
                                               1e-8, None, ptr(norm), None))
        torch.cuda.synchronize()
        assert abs(float(norm) - float(ref_grads.norm())) <= 1e-5 * float(ref_grads.norm())
        host, off = params[0].cpu(), 0
        with torch.no_grad():
            for k in names:
                n = o.sd[k].numel()
                o.sd[k].copy_(host[off:off + n].view_as(o.sd[k]))
                off += n
    bad = total = 0
    for k, v in fix["final"].items():
        bad += int(((o.sd[k].detach() - v).abs() > 1e-4 + 1e-4 * v.abs()).sum())
        total += v.numel()
    assert bad <= total * 1e-3, (bad, total)


def test_standalone_entry_points_reject_bad_arguments():
    from maexp_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(64, device="cuda")
    i = torch.zeros(4, dtype=torch.int32, device="cuda")
    with pytest.raises(_lib.MaexpError):
        _lib.check(lib.maexp_loss_c51(ptr(x), ptr(x), ptr(x), ptr(i), ptr(x), ptr(x), 0.9, -10.0, 10.0, 0.4, 1, 4, 4, 200,
                                      ptr(x), ptr(x), None))  # more than 128 atoms
    with pytest.raises(_lib.MaexpError):
        _lib.check(lib.maexp_loss_dqn(ptr(x), None, ptr(i), ptr(x), ptr(i), 0.9, 1, 4, 4, ptr(x), ptr(x), None))
    with pytest.raises(_lib.MaexpError):
        _lib.check(lib.maexp_clip_rmsprop_step(ptr(x), ptr(x), ptr(x), 0, 1, 1.0, 0.1, 1e-3, 0.99, 1e-8, None, None, None))
