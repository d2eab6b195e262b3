"""The host-side mirror keeps the reference's constructor / forward signatures and the dropin/ shims
resolve the reference's import names."""
import importlib
import inspect
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.name != "self"]


def test_signatures_match_reference():
    import hypernet_image_captioning_b200 as C
    E = inspect.Parameter.empty
    # hypernet_attention.py:33
    assert _params(C.HyperNetAttention.__init__) == [
        ("feature_size", E), ("embed_size", E), ("hidden_size", E), ("vocab_size", E), ("vocab", E), ("num_layers", 1),
        ("lr", 1e-6), ("mixup", False), ("alpha", 0.3), ("cc", False), ("hyper_emb", 10)]
    assert _params(C.HyperNetAttention.forward) == [("x", E)]
    # hypernet.py:27
    assert _params(C.HyperNetPooled.__init__) == [
        ("embed_size", E), ("hidden_size", E), ("vocab_size", E), ("vocab", E), ("num_layers", 1), ("type", "gru"),
        ("lr", 1e-6)]
    # models/decoderlstm.py:12,49,122
    assert _params(C.AttentionGru.__init__) == [
        ("num_features", E), ("feature_out", E), ("embedding_dim", E), ("hidden_dim", E), ("vocab_size", E),
        ("num_layers", 1), ("p", 0.0)]
    assert _params(C.AttentionGru.forward)[:3] == [("features", E), ("captions", E), ("sample_prob", 0.0)]
    assert _params(C.AttentionGru.init_hidden) == [("features", E)]
    # later.py:363,389,459
    assert _params(C.DecoderGRU.__init__) == [
        ("embed_size", E), ("hidden_size", E), ("vocab_size", E), ("num_layers", 1), ("dropout", False)]
    assert _params(C.DecoderGRU.forward)[:3] == [("features", E), ("captions", E), ("teacher_forcing", True)]
    assert _params(C.DecoderGRU.infer)[:2] == [("features", E), ("max_len", 50)]
    # models/attention.py:9
    assert _params(C.BahdanauAttention.__init__) == [("num_features", E), ("hidden_dim", E), ("output_dim", 1)]


def test_attributes_read_by_the_reference_trainers_exist():
    """cc_train_hypernet.py:110-120,131,151 read these attributes off the hypernet."""
    import hypernet_image_captioning_b200 as C
    m = C.HyperNetAttention(16, 12, 20, 50, None, cc=True, hyper_emb=10)
    for attr in ("hn_heads", "hn_base", "image_encoder", "captioner"):
        assert hasattr(m, attr)
    for attr in ("feature_fc", "embed", "fc", "attention", "init_h", "gru"):
        assert hasattr(m.captioner, attr)
    assert m.hparams["embed_size"] == 12 and m.teacher_forcing_proba == 0.0
    opt, sch = m.configure_optimizers()
    assert len(opt) == 1 and sch[0]["monitor"] == "val_loss with TF"


def test_dropin_shims_resolve_reference_import_names():
    d = os.path.join(ROOT, "dropin")
    sys.path.insert(0, d)
    try:
        for mod in ("hypernet_attention", "hypernet", "models.decoderlstm", "models.attention"):
            sys.modules.pop(mod, None)
        sys.modules.pop("models", None)
        import hypernet_image_captioning_b200 as C
        assert importlib.import_module("hypernet_attention").HyperNet is C.HyperNetAttention
        assert importlib.import_module("hypernet").HyperNet is C.HyperNetPooled
        dl = importlib.import_module("models.decoderlstm")
        assert dl.AttentionGru is C.AttentionGru and dl.DecoderGRU is C.DecoderGRU
        assert importlib.import_module("models.attention").BahdanauAttention is C.BahdanauAttention
    finally:
        sys.path.remove(d)
        for mod in ("hypernet_attention", "hypernet", "models.decoderlstm", "models.attention", "models"):
            sys.modules.pop(mod, None)


def test_theta_split_node_equals_autograd_slicing():
    """Fn.ThetaSplitFn (the generated-parameter views of utils.py:24-69 as ONE autograd node, incl. the offset-0 aliasing of
    extra cells, utils.py:45) == plain slicing + autograd: same views, same d(theta).  Pure torch: runs without a GPU."""
    import torch
    from hypernet_image_captioning_b200 import functional as Fn
    torch.manual_seed(0)
    for shapes in ((((6, 4), (6, 2), (6,), (6,)),),
                   (((6, 4), (6, 2), (6,), (6,)), ((6, 2), (6, 2), (6,), (6,))),
                   (((8, 3), (8, 2), (8,), (8,)), ((8, 2), (8, 2), (8,), (8,)), ((8, 2), (8, 2), (8,), (8,)))):
        n = sum(torch.Size(s).numel() for s in shapes[0]) + 5            # theta may be longer than what the cells use
        th = torch.randn(1, n, requires_grad=True)
        outs = Fn.ThetaSplitFn.apply(th, shapes)
        ws = [torch.randn_like(o) for o in outs]
        sum((o * w).sum() for o, w in zip(outs, ws)).backward()
        g1, th.grad = th.grad.clone(), None
        ref = []
        for cell in shapes:
            a = 0
            for shp in cell:
                k = torch.Size(shp).numel()
                ref.append(th[0][a:a + k].reshape(shp))
                a += k
        sum((o * w).sum() for o, w in zip(ref, ws)).backward()
        assert all(torch.equal(a, b) for a, b in zip(outs, ref))
        assert torch.allclose(g1, th.grad, rtol=0, atol=1e-6)
        # a missing gradient (an unused generated tensor) counts as zero
        th.grad = None
        outs = Fn.ThetaSplitFn.apply(th, shapes)
        (outs[0] * ws[0]).sum().backward()
        want = torch.zeros_like(th)
        want[0, :ws[0].numel()] = ws[0].reshape(-1)
        assert torch.allclose(th.grad, want)


def test_slice_readiness_marks_never_claim_cpu_tensors():
    """streams.wait_ranges (the partial wait of the fused decoder node for the generated W_ih / b_ih slices) must report
    'not covered' -- so the caller falls back to the full join -- for anything it holds no marks for."""
    import torch
    from hypernet_image_captioning_b200 import streams
    streams.clear_ranges()
    assert streams.wait_ranges((torch.zeros(4),)) is False
    streams.mark_range(torch.zeros(4))            # CPU tensors are not marked
    assert streams.wait_ranges((torch.zeros(4),)) is False
