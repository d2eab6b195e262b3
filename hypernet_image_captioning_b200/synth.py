"""Seeded synthetic inputs of the reference's shapes; used by bench.py / tools (no oracle dependency)."""
import torch


def synth_captions(B: int, T: int, V: int, gen: torch.Generator) -> torch.Tensor:
    """caps[b,0]=<s>=1, body ~ U{7..V-1}, caps[b,L-1]=</s>=2, 0-padded; L ~ clip(round(N(12.5,4)),4,T); row 0 is full."""
    L = torch.clamp(torch.round(torch.normal(12.5, 4.0, (B,), generator=gen)), 4, T).long()
    L[0] = T
    caps = torch.randint(7, V, (B, T), generator=gen)
    caps[:, 0] = 1
    idx = torch.arange(T).unsqueeze(0)
    caps[idx == (L - 1).unsqueeze(1)] = 2
    caps[idx >= L.unsqueeze(1)] = 0
    return caps
