"""CUDA-graph capture of a cast-heavy forward.

Every dmxq kernel is enqueued on the caller's current stream, never synchronises and never
allocates outside torch's caching allocator, so a whole BASIC-mode forward (hundreds of tiny
cast launches between the GEMMs) can be captured once and replayed: the per-cast python / launch
overhead (~30 us each, the dominant cost at small batch) disappears.

    fwd = dmx_compressor_b200.graph.capture(model, example_ids)   # static shapes
    logits = fwd(ids)                                             # copy-in, replay, result view
"""
from __future__ import annotations

import torch


def _drop_weight_caches(model) -> None:
    for m in model.modules():
        if getattr(m, "_wcache", None) is not None:
            m._wcache = None
        m.__dict__.pop("_dmxq_wcache", None)  # (the plugin's caches on the reference's own DmxModule objects)
        m.__dict__.pop("_dmxq_bcache", None)
        m.__dict__.pop("_bcache", None)


class Captured:
    """A captured forward.  What the graph reads besides ``static_in`` is FROZEN at capture time: the parameters and buffers
    of ``model`` (kept alive by the reference held here) and -- with ``elide_casts`` -- the cast weights computed during the
    capture; after a weight update, ``configure()`` or ``fold_weight_and_bias()``, capture again.

    With ``elide_casts`` the cast memo and the weight caches are cleared right before the capture: a memo hit on a tensor
    computed during warm-up would otherwise leave that cast OUT of the graph (the graph would keep reading the warm-up
    result, freed when the elision context ends), so every cast the forward needs is recorded; and they are cleared again
    afterwards, so nothing outside the graph keeps pointing into the graph's private memory pool."""

    def __init__(self, model, *example, warmup: int = 3, elide_casts: bool = False):
        from . import elide

        self.model = model
        self.static_in = [e.clone() for e in example]
        self.elide = elide_casts
        ctx = elide.enabled if elide_casts else _null
        with torch.no_grad(), ctx():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    model(*self.static_in)
            torch.cuda.current_stream().wait_stream(s)
            if elide_casts:
                elide.reset()
                _drop_weight_caches(model)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                # (a deferred output cast -- elide.Lazy -- must run INSIDE the graph: materialised later, eagerly, it would be
                # computed once from the first replay's data and then returned for every input)
                self.static_out = elide.materialise(model(*self.static_in))
            if elide_casts:
                elide.reset()
                _drop_weight_caches(model)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src)
        self.graph.replay()
        return self.static_out


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def capture(model, *example, **kw) -> Captured:
    return Captured(model, *example, **kw)
