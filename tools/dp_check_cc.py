"""Multi-GPU parity check of the domain-conditioned attention variant (BASELINE configs[3]; run under torchrun on >= 2 GPUs):
N-rank data-parallel gradients -- d(theta) [G, theta] all-reduced in front of the head backward, shared gradients in the
overlapped bucket, per-rank loss weighted by its share of the non-pad tokens -- equal the single-process gradients on the
concatenated global batch, for one domain per batch (cc_train_hypernet.py:134-153) and for G domains in one batch (grouped
kernels).  No oracle involved: the single-process path is what tests/test_gpu_fullsize.py pins against the oracle.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_check_cc.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypernet_image_captioning_b200 as C  # noqa: E402
from hypernet_image_captioning_b200 import parallel  # noqa: E402
from hypernet_image_captioning_b200.synth import synth_captions  # noqa: E402


def step(model, styles, feats, caps, groups, G, weight):
    model.zero_grad(set_to_none=True)
    if G == 1:
        captioner = model.forward(styles[0])                                  # 1-D one-hot, cc_train_hypernet.py:141-143
        logits, _ = captioner(feats, caps, 0.0)
    else:
        captioner = model.forward_grouped(styles)
        logits, _ = captioner(feats, caps, 0.0, groups=groups)
    loss = C.cross_entropy(logits, caps, 0)
    (loss * weight).backward()
    return loss.detach()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Fo = E = H = 200
    V, T, Bl, he = 1500, 6, 16, 5
    B = Bl * world
    g = torch.Generator().manual_seed(31)
    feats = torch.randn(B, 49, 2048, generator=g).to(dev)
    caps = synth_captions(B, T, V, g)
    caps[1::3, T // 2:] = 0                   # ragged padding: the ranks hold different numbers of non-pad tokens
    caps = caps.to(dev)
    eye = torch.eye(he, device=dev)

    def make():
        torch.manual_seed(0)
        with torch.device(dev):
            return C.HyperNetAttention(Fo, E, H, V, None, cc=True, hyper_emb=he)

    lines, worst = [], 0.0
    for G in (1, 3):
        groups = torch.arange(B) % G                      # host tensor (loader data)
        styles = eye[:G].contiguous()
        ref = make()
        step(ref, styles, feats, caps, groups, G, 1.0)
        ref_grads = {k: v.grad.clone() for k, v in ref.named_parameters() if v.grad is not None}
        del ref
        for overlap in (False, True):
            m = make()
            m.dp_enabled = True
            shared = parallel.shared_parameters(m)
            if overlap:
                parallel.enable_overlap(shared)
            sl = slice(rank * Bl, (rank + 1) * Bl)
            for _ in range(2):                            # twice: the per-step state of the overlap must reset
                w = parallel.loss_weight(caps[sl], 0)
                step(m, styles, feats[sl], caps[sl], groups[sl], G, w)
                parallel.allreduce_shared_grads(shared)
            parallel.disable_overlap()
            case = 0.0
            for k, v in m.named_parameters():
                if k.startswith("captioner.gru.") or k not in ref_grads:
                    continue
                d = (v.grad - ref_grads[k]).abs().max().item()
                s = ref_grads[k].abs().max().item()
                rel = d / s if s > 0 else d
                case = max(case, rel if d > 1e-7 else 0.0)
            worst = max(worst, case)
            lines.append(f"  G={G} overlap={int(overlap)}: max relative gradient difference {case:.3e}")
            del m
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        theta = 3 * H * (E + Fo) + 3 * H * H + 6 * H
        print(f"dp_check_cc world={world} (F=E=H={H}, V={V}, B={Bl}/rank, T={T}, he={he}; d(theta) all-reduce = [G, {theta}] floats)")
        print("\n".join(lines))
        print(f"max over ranks and cases = {t.item():.3e}")
        assert t.item() < 1e-3
        print("dp_check_cc OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
