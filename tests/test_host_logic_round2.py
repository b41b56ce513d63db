"""Host-side logic added in round 2 that needs no GPU: the per-model cache and its structural-edit
invalidation (reference behaviour: flashy/distrib.py:205-210 walks the module on every call), and the bucket
bookkeeping of the backward-overlap hooks (in-order launches, accumulation passes, held-back buckets)."""
import torch
from torch import nn

from flashy_b200 import distrib as D


def _model():
    return nn.Sequential(nn.Linear(8, 16), nn.BatchNorm1d(16), nn.Linear(16, 4))


def test_model_cache_survives_unrelated_modules_and_follows_edits():
    model = _model()
    first = D._model_entry(model)
    assert [id(p) for p in first.params] == [id(p) for p in model.parameters()]
    assert len(first.float_buffers) == 2 and len(first.buffers) == 3          # num_batches_tracked is int64
    nn.Linear(3, 3)                                                           # some OTHER module is built ...
    assert D._model_entry(model) is first                                     # ... this model's cache is kept
    model[2] = nn.Linear(16, 5)                                               # structural edit of THIS model
    second = D._model_entry(model)
    assert second is not first
    assert [id(p) for p in second.params] == [id(p) for p in model.parameters()]
    model.register_buffer("extra", torch.zeros(3))
    third = D._model_entry(model)
    assert third is not second and len(third.float_buffers) == 3
    model.register_parameter("scale", nn.Parameter(torch.ones(1)))
    assert len(D._model_entry(model).params) == len(second.params) + 1


class _Recorder(D._Overlap):
    """The hook bookkeeping without a communicator: launches are recorded instead of sent."""

    def __init__(self, entry, cap):
        self.sent = []
        super().__init__(None, None, entry, cap)

    def _launch(self, tag, tensors):
        self.sent.append((tag, len(tensors)))
        self.pending = True

    def _join(self):
        self.sent.append(("join", 0))


def _overlap(n_layers=6, width=32, cap=3 * 32 * 32 * 4):
    model = nn.Sequential(*[nn.Linear(width, width) for _ in range(n_layers)])
    entry = D._ModelLists(model)
    import os
    os.environ["FLASHY_B200_OVERLAP_TAIL_KB"] = "5"                            # tail = the first layer (4.1 KB)
    try:
        ov = _Recorder(entry, cap)
    finally:
        del os.environ["FLASHY_B200_OVERLAP_TAIL_KB"]
    return model, ov


def test_overlap_buckets_partition_the_parameters():
    model, ov = _overlap()
    try:
        flat = [i for b in ov.buckets for i in b]
        assert sorted(flat) == list(range(len(ov.params)))                    # every parameter exactly once
        assert ov.buckets[-1] == [1, 0]                                       # tail: first registered layer (weight, bias)
        assert ov.buckets[0][0] == len(ov.params) - 1                         # first bucket starts at the last parameter
        assert ov.n_hook == len(ov.buckets) - 1 >= 2
    finally:
        ov.remove()


def test_overlap_launches_in_order_once_per_backward_pass():
    model, ov = _overlap()
    try:
        x = torch.randn(4, 32)
        model(x).sum().backward()                                             # the real autograd hooks drive it
        hook = [t for t, _ in ov.sent]
        assert hook == [("hook", k) for k in range(ov.n_hook)]                # strictly in bucket order
        model(x).sum().backward()                                             # accumulation: every pass sends again
        assert [t for t, _ in ov.sent] == [("hook", k) for k in range(ov.n_hook)] * 2
        ov.finish([torch.zeros(3)])
        assert ov.sent[-2:] == [(("tail", 3), 3), ("join", 0)] and not ov.pending   # 2 tail gradients + 1 buffer, then the join
        assert ov.next == 0 and ov.launched == 0 and ov.left == [len(b) for b in ov.buckets]
    finally:
        ov.remove()


def test_overlap_holds_back_buckets_with_a_missing_gradient():
    model, ov = _overlap()
    try:
        model.zero_grad(set_to_none=True)
        x = torch.randn(4, 32)
        # the LAST layer is skipped: the first bucket (in arrival order) never completes, nothing may be sent early
        nn.Sequential(*list(model)[:-1])(x).sum().backward()
        assert ov.sent == []
        ov.finish([])
        assert ov.sent[-1] == ("join", 0)
        tags = [t for t, _ in ov.sent[:-1]]
        assert tags[-1][0] == "tail" and all(t[0] == "late" for t in tags[:-1])
        sent_params = sum(n for _, n in ov.sent)
        assert sent_params == len(ov.params) - 2                              # everything except the skipped layer
    finally:
        ov.remove()


def test_signaller_coverage_formula_matches_a_brute_force_model():
    """fx_fuse.cu, signal role: reduce warp w owns units w, w + R, ... of the chunk-major unit sequence and reports
    how many it has finished; chunks are complete up to min_w floor((done_w * R + w) / units_per_chunk).  Checked
    against an explicit simulation for every reachable progress state of small configurations."""
    import itertools
    R = 4
    for chunks, upc in ((1, 1), (5, 1), (3, 2), (4, 4), (2, 8), (7, 3)):
        units = chunks * upc
        owned = [len(range(w, units, R)) for w in range(R)]
        for done in itertools.product(*[range(n + 1) for n in owned]):
            finished = set()
            for w in range(R):
                finished.update(list(range(w, units, R))[:done[w]])
            brute = 0
            while brute < chunks and all(u in finished for u in range(brute * upc, (brute + 1) * upc)):
                brute += 1
            formula = min(min((done[w] * R + w) // upc, chunks) for w in range(R))
            assert formula == brute, (chunks, upc, done, formula, brute)
