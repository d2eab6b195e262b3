"""Multi-GPU parity check (run under torchrun on >= 2 GPUs): N-rank data-parallel gradients of the pooled-variant step
equal the single-process gradients on the concatenated global batch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypernet_image_captioning_b200 as C  # noqa: E402
from hypernet_image_captioning_b200 import parallel  # noqa: E402
from oracle import caption_hn_oracle as O  # noqa: E402  (input generators + parameter init only)


def run(model, style, pooled, caps, h0, scale):
    model.zero_grad(set_to_none=True)
    cap = model.forward(style)
    logits = cap(model.image_encoder(pooled), caps, True, h0=h0)
    loss = C.cross_entropy(logits, caps, None)
    (loss * scale).backward()
    return loss.detach()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    E, H, V, T, Bl = 24, 30, 311, 9, 6
    p = O.init_params_pooled(2048, E, H, V, seed=5)
    g = torch.Generator().manual_seed(99)
    B = Bl * world
    pooled = torch.relu(torch.randn(B, 2048, generator=g)).to(dev)
    caps = O.synth_captions(B, T, V, g).to(dev)
    style_plain = torch.randn(1, E, generator=g).to(dev)
    h0 = torch.rand(B, H, generator=g).to(dev)

    def make():
        m = C.HyperNetPooled(E, H, V, None)
        sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
        return m.to(dev)

    worst = 0.0
    # style = a plain tensor, and style = a row of the embedding matrix (the Flickr idiom, hypernet_attention.py:139-142:
    # that parameter then receives a second, replicated, gradient contribution through the hypernet);
    # shared-gradient bucket reduced after the backward, and overlapped with the head backward (parallel.enable_overlap)
    for style_from_embed in (False, True):
        for overlap in (False, True):
            ref = make()
            run(ref, ref.captioner.embed.weight[4:5] if style_from_embed else style_plain, pooled, caps, h0, 1.0)
            ref_grads = {k: v.grad.clone() for k, v in ref.named_parameters() if v.grad is not None}
            m = make()
            m.dp_enabled = True
            shared = parallel.shared_parameters(m)
            if overlap:
                parallel.enable_overlap(shared)
            sl = slice(rank * Bl, (rank + 1) * Bl)
            for _ in range(2):                                      # twice: the per-step state of the overlap must reset
                run(m, m.captioner.embed.weight[4:5] if style_from_embed else style_plain, pooled[sl], caps[sl], h0[sl],
                    1.0 / world)
                parallel.allreduce_shared_grads(shared)
            parallel.disable_overlap()
            for k, v in m.named_parameters():
                if k.startswith("captioner.lstm_cell."):
                    continue
                d = (v.grad - ref_grads[k]).abs().max().item()
                s = ref_grads[k].abs().max().item()
                rel = d / s if s > 0 else d
                worst = max(worst, rel if d > 1e-7 else 0.0)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dp_check world={world}: max relative gradient difference vs single-process global batch = {t.item():.3e}")
        assert t.item() < 1e-3
        print("dp_check OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
