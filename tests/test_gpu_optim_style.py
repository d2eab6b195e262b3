"""FusedAdam (+ folded gradient clipping) against torch.optim.Adam + torch.nn.utils.clip_grad_norm_ -- the optimizer the
reference trainers configure (cc_train_hypernet.py:110-122,405)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.3).cuda()) for s in shapes]


@pytest.mark.parametrize("clip", [None, 5.0, 0.05])
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(clip, wd):
    import hypernet_image_captioning_b200 as C
    shapes = [(257, 33), (1000,), (3,), (64, 64), (1,), (90000, 13)]
    ref_p, our_p = _params(shapes, 1), _params(shapes, 1)
    ref = torch.optim.Adam(ref_p, lr=3e-3, weight_decay=wd)
    our = C.FusedAdam(our_p, lr=3e-3, weight_decay=wd, max_grad_norm=clip)
    g = torch.Generator().manual_seed(2)
    for it in range(4):
        grads = [(torch.randn(*s, generator=g) * (0.5 if it else 3.0)).cuda() for s in shapes]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        if clip is not None:
            n_ref = torch.nn.utils.clip_grad_norm_(ref_p, clip)
        ref.step()
        our.step()
        if clip is not None:
            assert abs(float(our.last_grad_norm) - float(n_ref)) <= 2e-6 * float(n_ref)
            for q, gr in zip(our_p, grads):
                assert torch.equal(q.grad, gr)          # the stored gradient is not modified
        # with weight decay a few elements have clip*g + wd*p ~ 0: there m/(sqrt(v)+eps) is ill-conditioned and the last
        # bit of the global norm (fp32 in torch, double here) moves the update by ~0.3 % of lr
        tol = 2e-6 if wd == 0.0 else 2e-5
        for p, q in zip(ref_p, our_p):
            assert (p - q).abs().max().item() <= tol * max(1.0, p.abs().max().item()), (it, tuple(p.shape))
    for p, q in zip(ref_p, our_p):
        sr, so = ref.state[p], our.state[q]
        assert int(sr["step"]) == int(so["step"])
        for k in ("exp_avg", "exp_avg_sq"):      # elementwise ulp-level differences relative to the tensor's scale
            assert (sr[k] - so[k]).abs().max().item() <= tol * sr[k].abs().max().item(), k


def test_fused_adam_state_dict_roundtrip_and_lr_scheduler():
    """state_dict()/load_state_dict() and ReduceLROnPlateau (cc_train_hypernet.py:121) work as with torch's Adam."""
    import hypernet_image_captioning_b200 as C
    p = _params([(50, 7)], 3)
    opt = C.FusedAdam(p, lr=1e-2, max_grad_norm=5.0)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, cooldown=2, factor=0.5, patience=0)
    p[0].grad = torch.ones_like(p[0])
    opt.step()
    sched.step(1.0); sched.step(2.0)
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-3)
    import copy
    sd = copy.deepcopy(opt.state_dict())      # a checkpoint (torch.save / torch.load) owns its tensors
    q = _params([(50, 7)], 3)
    opt2 = C.FusedAdam(q, lr=1e-2, max_grad_norm=5.0)
    opt2.load_state_dict(sd)
    q[0].data.copy_(p[0].data)
    p[0].grad = torch.full_like(p[0], 0.5); q[0].grad = torch.full_like(q[0], 0.5)
    opt.step(); opt2.step()
    assert torch.equal(p[0], q[0])


def test_fused_adam_trains_the_attention_hypernet_like_torch_adam():
    """Three optimizer steps of the flow-mode caption loss on the tiny attention model: same losses as torch Adam + clip."""
    import hypernet_image_captioning_b200 as C
    from hypernet_image_captioning_b200.synth import synth_captions
    losses = []
    for fused in (False, True):
        torch.manual_seed(0)
        with torch.device("cuda"):
            m = C.HyperNetAttention(16, 12, 20, 50, None)
        g = torch.Generator().manual_seed(5)
        feats = torch.randn(6, 49, 2048, generator=g).cuda()
        caps = synth_captions(6, 7, 50, g).cuda()
        params = [p for n, p in m.named_parameters() if not n.startswith("captioner.gru.")]
        opt = C.FusedAdam(params, lr=1e-2, max_grad_norm=5.0) if fused else torch.optim.Adam(params, lr=1e-2)
        ls = []
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            cap = m.forward(m.captioner.embed.weight[4:5].detach())
            loss, _, _ = cap.forward_loss(feats, caps, 0.0, ignore_index=0)
            loss.backward()
            if not fused:
                torch.nn.utils.clip_grad_norm_(params, 5.0)
            opt.step()
            ls.append(float(loss.detach()))
        losses.append(ls)
    assert losses[0][0] == pytest.approx(losses[1][0], rel=1e-6)
    assert losses[0][2] == pytest.approx(losses[1][2], rel=1e-4)
    assert losses[1][2] < losses[1][0]


@pytest.mark.parametrize("kind,nin", [("one hot", None), ("embedding", None), ("histogram", 61), ("JSD", 2)])
def test_domain_embedding_front_ends_match_torch_modules(kind, nin):
    """cc_train_hypernet.py:86-106,137-149: the four ways a domain name becomes the hypernet input; values and gradients
    against the same torch modules (nn.Embedding / nn.Sequential(Linear, LeakyReLU, ...))."""
    import hypernet_image_captioning_b200 as C
    domains = ["news\n", "sport", "food", "travel"]
    g = torch.Generator().manual_seed(11)
    vectors = {d.replace("\n", ""): torch.randn(nin, generator=g) for d in domains} if nin else None
    torch.manual_seed(1)
    m = C.DomainEmbedding(kind, domains, hyper_emb=10, in_features=nin, vectors=vectors).cuda()
    assert m.dict_domain == {"news": 0, "sport": 1, "food": 2, "travel": 3}
    out = m("food")
    assert out.shape == (len(domains) if kind == "one hot" else 10,) and out.is_cuda
    if kind == "one hot":
        assert out.tolist() == [0.0, 0.0, 1.0, 0.0]
        return
    w = torch.randn(10, generator=g).cuda()
    (out * w).sum().backward()
    if kind == "embedding":
        ref = m.embed.weight.detach()[2]
        assert torch.equal(out.detach(), ref)
        assert torch.allclose(m.embed.weight.grad[2], w) and float(m.embed.weight.grad[[0, 1, 3]].abs().max()) == 0.0
        return
    import copy
    ref_m = copy.deepcopy(m.embed)
    for p_ in ref_m.parameters():
        p_.grad = None
    ref = ref_m(vectors["food"].cuda())
    (ref * w).sum().backward()
    assert torch.allclose(out.detach(), ref.detach(), rtol=1e-5, atol=1e-6)
    for (n1, p1), (_, p2) in zip(m.embed.named_parameters(), ref_m.named_parameters()):
        assert torch.allclose(p1.grad, p2.grad, rtol=1e-4, atol=1e-6), n1


@pytest.mark.parametrize("variant", ["pooled", "attention"])
def test_lowrank_head_gradients_give_the_same_training_trajectory(variant):
    """head_grad_mode = "lowrank": the large head matrices keep dW2 = dtheta^T a as (dtheta, a); FusedAdam forms it on the
    fly and takes its norm from Gram matrices.  Parameters after 3 clipped Adam steps must equal the dense path."""
    import hypernet_image_captioning_b200 as C
    from hypernet_image_captioning_b200.synth import synth_captions
    results = []
    for mode in ("materialize", "lowrank"):
        torch.manual_seed(0)
        g = torch.Generator().manual_seed(5)
        with torch.device("cuda"):
            m = C.HyperNetPooled(16, 24, 90, None) if variant == "pooled" else C.HyperNetAttention(16, 12, 20, 50, None)
        m.head_grad_mode = mode
        m.lowrank_min_numel = 2000            # tiny model: make its larger head matrices qualify
        V = 90 if variant == "pooled" else 50
        caps = synth_captions(6, 7, V, g).cuda()
        if variant == "pooled":
            pooled, h0 = torch.relu(torch.randn(6, 2048, generator=g)).cuda(), torch.rand(6, 24, generator=g).cuda()
        else:
            feats = torch.randn(6, 49, 2048, generator=g).cuda()
        params = [p for n, p in m.named_parameters() if not (n.startswith("captioner.gru.") or n.startswith("captioner.lstm_cell."))]
        opt = C.FusedAdam(params, lr=1e-2, max_grad_norm=0.5)
        n_lowrank, norms = 0, []
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            cap = m.forward(m.captioner.embed.weight[4:5].detach())
            if variant == "pooled":
                loss = cap.forward_loss(m.image_encoder(pooled), caps, h0=h0)[0]
            else:
                loss = cap.forward_loss(feats, caps, 0.0, ignore_index=0)[0]
            loss.backward()
            n_lowrank += sum(getattr(p, "grad_lowrank", None) is not None for p in params)
            opt.step()
            norms.append(float(opt.last_grad_norm))
            assert all(getattr(p, "grad_lowrank", None) is None for p in params)      # consumed by the step
        assert (n_lowrank > 0) == (mode == "lowrank")
        results.append(({n: p.detach().clone() for n, p in m.named_parameters()}, norms))
    (pa, na), (pb, nb) = results
    for x, y in zip(na, nb):
        assert abs(x - y) <= 1e-5 * x
    for n in pa:
        if n.startswith("captioner.gru.") or n.startswith("captioner.lstm_cell."):
            continue          # the last generated weights (a function of the parameters, not trained)
        if n.endswith("attention.v_a.bias"):
            continue          # its gradient is analytically zero (softmax shift invariance): Adam normalises rounding noise
        # updates are lr * O(1) = 1e-2 per step: 5e-5 is 0.5 % of one update.  The two runs are not bit-reproducible (the dA
        # reductions use atomics, so their summation order varies run to run) and Adam's m / sqrt(v) amplifies last-bit
        # gradient differences of near-cancelling elements: observed spread 0 ... 1.3e-5 between identical runs.
        assert (pa[n] - pb[n]).abs().max().item() <= 5e-5 * max(1.0, pa[n].abs().max().item()), n


def test_train_init_regression_loop_runs_on_the_streaming_kernels():
    """train_init.py:70-123 applies ``hn_base(style_embed)`` and ``hn_heads[i](base)`` itself (plain module calls).  Those
    modules are RowsLinear layers here, so the unedited loop runs the weight-streaming kernels; loss and gradients equal the
    torch reference of the same loop."""
    import torch
    import hypernet_image_captioning_b200 as C
    from golden_util import grad_close
    torch.manual_seed(0)
    m = C.HyperNetAttention(16, 12, 20, 60, None).cuda()
    ref = {k: v.detach().clone().requires_grad_(True) for k, v in m.named_parameters() if k.startswith("hn_")}
    g = torch.Generator().manual_seed(2)
    style_embed = torch.randn(1, 12, generator=g).cuda()
    targets = [torch.randn(*p.shape, generator=g).cuda() * 0.1 for p in m.captioner.gru.parameters()]
    n0 = C._cabi.launches()
    crit = torch.nn.MSELoss()
    base = m.hn_base(style_embed)                                   # the reference's own lines (train_init.py:78, 90-91)
    loss = 0.
    for i, W in enumerate(targets):
        loss = loss + crit(m.hn_heads[i](base).flatten(), W.flatten())
    loss.backward()
    assert C._cabi.launches() - n0 >= 10                           # 2 base + 8 head layers forward on our kernels
    lrelu = torch.nn.functional.leaky_relu
    b = lrelu(lrelu(style_embed @ ref["hn_base.0.weight"].t() + ref["hn_base.0.bias"]) @ ref["hn_base.2.weight"].t() + ref["hn_base.2.bias"])
    rl = 0.
    for i, W in enumerate(targets):
        a = lrelu(b @ ref[f"hn_heads.{i}.0.weight"].t() + ref[f"hn_heads.{i}.0.bias"])
        rl = rl + crit((a @ ref[f"hn_heads.{i}.2.weight"].t() + ref[f"hn_heads.{i}.2.bias"]).flatten(), W.flatten())
    rl.backward()
    assert abs(loss.item() - rl.item()) < 1e-5 * abs(rl.item())
    for k, v in m.named_parameters():
        if k.startswith("hn_"):
            assert grad_close(v.grad, ref[k].grad, 1e-3), k
