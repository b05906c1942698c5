"""TEST INFRASTRUCTURE ONLY (oracle shim) -- never imported by the product path.

Stand-in for the un-installed third-party ``torch_scatter`` so the UNMODIFIED reference
(/root/reference) can be imported on CPU in the build container (shim 1).
Semantics restated from the torch_scatter 2.x docs (reference pins "torch-scatter for
torch-1.12.0+cu116", README.md:28): rows = dim_size or index.max()+1; ``sum``; ``mean`` =
sum / clamp(count, 1) (floor division for integer dtypes); ``max`` returns values only and
leaves EMPTY segments at 0.  Call sites: models/motionnet.py:159-160,
models/pillar_encoder.py:116,120, models/tpointnet.py:227-284, models/alignnet.py:133-134.
"""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None
    index = index.long()
    n = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    tail = src.shape[1:]
    idx = index.view((-1,) + (1,) * len(tail)).expand_as(src)
    if reduce in ("sum", "add"):
        return torch.zeros((n,) + tail, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
    if reduce == "mean":
        s = torch.zeros((n,) + tail, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
        cnt = torch.zeros(n, dtype=torch.long, device=src.device).scatter_add_(0, index, torch.ones_like(index))
        cnt = cnt.clamp(min=1).view((-1,) + (1,) * len(tail))
        if src.is_floating_point():
            return s / cnt.to(src.dtype)
        return torch.div(s, cnt, rounding_mode="floor")
    if reduce == "max":
        o = torch.zeros((n,) + tail, dtype=src.dtype, device=src.device)
        o.scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
        return o
    raise ValueError(reduce)
