"""Data-parallel plumbing on CPU (gloo, world_size 2): the d(theta) all-reduce hook and the flat shared-gradient bucket
reproduce the single-process global-batch gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _toy(theta_src, W, xb, scale):
    """theta = theta_src**2 (stands in for the replicated hypernet); per-rank loss = scale * sum((xb @ W) * theta)."""
    from hypernet_image_captioning_b200.parallel import allreduce_grad
    theta = theta_src * theta_src
    theta = allreduce_grad(theta)
    return scale * ((xb @ W) * theta).sum()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hypernet_image_captioning_b200.parallel import allreduce_shared_grads
    g = torch.Generator().manual_seed(0)
    theta_src = torch.randn(5, generator=g).requires_grad_(True)
    W = torch.nn.Parameter(torch.randn(3, 5, generator=g))
    x = torch.randn(8, 3, generator=g)
    xb = x[rank * 4:(rank + 1) * 4]
    loss = _toy(theta_src, W, xb, 1.0 / world)
    loss.backward()
    cnt = torch.tensor([float(xb.shape[0])])
    allreduce_shared_grads([W], extra=cnt)
    torch.save({"theta_grad": theta_src.grad, "W_grad": W.grad, "cnt": cnt}, os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_dp_gradients_match_single_process(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    theta_src = torch.randn(5, generator=g).requires_grad_(True)
    W = torch.nn.Parameter(torch.randn(3, 5, generator=g))
    x = torch.randn(8, 3, generator=g)
    theta = theta_src * theta_src
    loss = (1.0 / world) * ((x @ W) * theta).sum()
    loss.backward()
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        # the hypernet-side gradient is identical on every rank after the d(theta) all-reduce ...
        assert torch.allclose(d["theta_grad"], theta_src.grad, atol=1e-6)
        # ... and the shared-parameter bucket sums the per-rank contributions
        assert torch.allclose(d["W_grad"], W.grad, atol=1e-6)
        assert d["cnt"].item() == 8.0


def test_shared_parameter_selection():
    import hypernet_image_captioning_b200 as C
    from hypernet_image_captioning_b200.parallel import shared_parameters
    m = C.HyperNetPooled(8, 6, 50, None)
    names = {n for n, p in m.named_parameters() if any(p is q for q in shared_parameters(m))}
    assert all(not n.startswith("hn_") for n in names)
    assert {"captioner.embed.weight", "captioner.fc_out.weight", "image_encoder.fc.weight"} <= names


def _worker_masked(rank, world, port, out):
    """Masked mean loss (ignore_index = 0) with different numbers of valid tokens per rank."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hypernet_image_captioning_b200.parallel import allreduce_shared_grads, loss_weight
    g = torch.Generator().manual_seed(0)
    W = torch.nn.Parameter(torch.randn(6, 7, generator=g))
    x = torch.randn(8, 5, 6, generator=g)
    caps = torch.randint(1, 7, (8, 5), generator=g)
    caps[0:4, 3:] = 0                       # rank 0 holds the heavily padded rows
    caps[5, 4] = 0
    xb, cb = x[rank * 4:(rank + 1) * 4], caps[rank * 4:(rank + 1) * 4]
    loss = torch.nn.functional.cross_entropy((xb @ W).reshape(-1, 7), cb.reshape(-1), ignore_index=0)
    w = loss_weight(cb, 0)
    (loss * w).backward()
    allreduce_shared_grads([W])
    torch.save({"W_grad": W.grad, "w": w}, os.path.join(out, f"m{rank}.pt"))
    dist.destroy_process_group()


def test_dp_masked_loss_weight_reproduces_global_mean(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_masked, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    W = torch.nn.Parameter(torch.randn(6, 7, generator=g))
    x = torch.randn(8, 5, 6, generator=g)
    caps = torch.randint(1, 7, (8, 5), generator=g)
    caps[0:4, 3:] = 0
    caps[5, 4] = 0
    torch.nn.functional.cross_entropy((x @ W).reshape(-1, 7), caps.reshape(-1), ignore_index=0).backward()
    ws = []
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"m{r}.pt"))
        assert torch.allclose(d["W_grad"], W.grad, atol=1e-6)      # == gradient of the global-batch masked mean
        ws.append(float(d["w"]))
    n0, n1 = int((caps[:4] != 0).sum()), int((caps[4:] != 0).sum())
    assert ws[0] == pytest.approx(n0 / (n0 + n1)) and ws[1] == pytest.approx(n1 / (n0 + n1)) and n0 != n1


def _worker_overlap(rank, world, port, out):
    """enable_overlap: early gradients are reduced from inside the backward (flush in the d(theta) hook), the late one
    (a parameter that also feeds the hypernet input, like captioner.embed.weight) by allreduce_shared_grads."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hypernet_image_captioning_b200 import parallel
    g = torch.Generator().manual_seed(0)
    emb = torch.nn.Parameter(torch.randn(6, 5, generator=g))      # row 2 is the "style" (hypernet input)
    W = torch.nn.Parameter(torch.randn(3, 5, generator=g))
    hn = torch.nn.Parameter(torch.randn(5, generator=g))          # replicated hypernet parameter: never all-reduced
    x = torch.randn(8, 3, generator=g)
    idx = torch.randint(0, 6, (8,), generator=g)
    xb, ib = x[rank * 4:(rank + 1) * 4], idx[rank * 4:(rank + 1) * 4]
    assert parallel.enable_overlap([emb, W])
    flushed = {}
    for it in range(2):                                            # two steps: per-step state must reset
        emb.grad = W.grad = hn.grad = None
        x = emb[2:3]
        leaf = parallel.route_style_grad(x) if it == 1 else None          # step 1: routed (style row detached), step 0: scaled
        style = leaf if leaf is not None else parallel.ScaleGradFn.apply(x, 1.0 / world)
        theta = parallel.allreduce_grad((style * hn).reshape(-1) ** 2)
        loss = (1.0 / world) * (((xb @ W) + emb[ib]) * theta).sum()
        loss.backward()
        flushed[it] = sorted(id(p) == id(W) for p in parallel._ov.params if id(p) in parallel._ov.done)
        parallel.allreduce_shared_grads([emb, W])
    torch.save({"emb": emb.grad, "W": W.grad, "hn": hn.grad, "flushed": flushed}, os.path.join(out, f"o{rank}.pt"))
    parallel.disable_overlap()
    dist.destroy_process_group()


def test_overlapped_bucket_matches_single_process(tmp_path):
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_overlap, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    emb = torch.nn.Parameter(torch.randn(6, 5, generator=g))
    W = torch.nn.Parameter(torch.randn(3, 5, generator=g))
    hn = torch.nn.Parameter(torch.randn(5, generator=g))
    x = torch.randn(8, 3, generator=g)
    idx = torch.randint(0, 6, (8,), generator=g)
    theta = (emb[2:3] * hn).reshape(-1) ** 2
    ((1.0 / world) * (((x @ W) + emb[idx]) * theta).sum()).backward()
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"o{r}.pt"))
        assert torch.allclose(d["W"], W.grad, atol=1e-6)
        assert torch.allclose(d["emb"], emb.grad, atol=1e-6)      # incl. the style row: hypernet-path gradient counted once
        assert torch.allclose(d["hn"], hn.grad, atol=1e-6)
        # step 0: W went early (inside backward), emb late; step 1 (style row routed): both early
        assert d["flushed"][0] == [True] and d["flushed"][1] == [False, True]
