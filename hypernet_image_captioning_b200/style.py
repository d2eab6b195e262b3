"""Style / domain embedding front-ends of the trainers: what turns a domain name into the vector the
hypernetwork consumes.  Mirrors cc_train_hypernet.py:63-106 (construction) and :137-149 (lookup in training_step):

  kind            reference                                              output
  "one hot"       F.one_hot(ids, #domains)[i].float()        (:86-89,141-144)   [he] 1-D, he = #domains (the cc=True input)
  "embedding"     nn.Embedding(#domains, hyper_emb)(i)        (:90-92,138-140)   [hyper_emb] 1-D
  "histogram"     Linear(V+1,4he) LReLU Linear(4he,he) LReLU  (:93-100,146-149)  [he] 1-D  (histograme / log / tfidf vectors)
  "JSD"           Linear(n_tsne,he) LReLU                     (:101-106)         [he] 1-D

The per-domain histogram / t-SNE vectors themselves are CPU preprocessing of the caption files (utils.tfidf_hist,
get_hist_embedding, get_jsd_tsne) and are passed in as a dict -- out of scope here.  The Linear layers run on the
weight-streaming kernels through functional.rows_linear, so the caption loss back-propagates into them in flow mode.
"""
import torch
import torch.nn as nn

from . import functional as Fn


class DomainEmbedding(nn.Module):
    def __init__(self, kind, domains, hyper_emb=10, in_features=None, vectors=None):
        """domains: list of domain names (index = position).  vectors: dict name -> 1-D tensor for "histogram" / "JSD"
        (in_features = len(vocab)+1 or n_tsne)."""
        super().__init__()
        if kind not in ("one hot", "embedding", "histogram", "JSD"):
            raise ValueError(f"unknown embedding kind {kind!r}")
        self.kind = kind
        self.dict_domain = {d.replace("\n", ""): i for i, d in enumerate(domains)}     # cc_train_hypernet.py:81-82
        self.vectors = vectors
        if kind == "one hot":
            self.hyper_emb = len(self.dict_domain)
            self.embed = None
        elif kind == "embedding":
            self.hyper_emb = hyper_emb
            self.embed = nn.Embedding(len(self.dict_domain), hyper_emb)
        elif kind == "histogram":
            self.hyper_emb = hyper_emb
            self.embed = nn.Sequential(nn.Linear(in_features, hyper_emb * 4), nn.LeakyReLU(),
                                       nn.Linear(hyper_emb * 4, hyper_emb), nn.LeakyReLU())
        else:
            self.hyper_emb = hyper_emb
            self.embed = nn.Sequential(nn.Linear(in_features, hyper_emb), nn.LeakyReLU())

    def forward(self, domain):
        """domain: a domain name.  Returns the 1-D style vector on the module's device."""
        dev = next(self.parameters()).device if self.embed is not None else torch.device("cuda")
        if self.kind == "one hot":
            out = torch.zeros(self.hyper_emb, device=dev)
            out[self.dict_domain[domain]] = 1.0
            return out
        if self.kind == "embedding":
            return self.embed.weight[self.dict_domain[domain]]
        x = self.vectors[domain].to(device=dev, dtype=torch.float32).reshape(1, -1)
        layers = [m for m in self.embed if isinstance(m, nn.Linear)]
        for lin in layers:                                   # every Linear of both MLPs is followed by LeakyReLU(0.01)
            x = Fn.rows_linear(x, lin.weight, lin.bias, leaky=True)
        return x.reshape(-1)
