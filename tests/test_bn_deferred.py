"""Host logic of the deferred BatchNorm running-statistics update (mono_vifi_b200/bn_act.py): a pass that runs
concurrently with another pass over the same modules records its batch statistics with momentum 1 into private buffers;
apply_deferred must then leave the module buffers exactly where two sequential nn.BatchNorm2d calls leave them
(train.py:943-954 calls the pose networks twice per step)."""
import torch
import torch.nn.functional as F

from mono_vifi_b200 import bn_act


def _batch_stats(x, C):
    """what the CUDA kernel writes in deferred mode: momentum 1 into zeroed buffers = (mean, unbiased variance)"""
    tmp = torch.zeros(2 * C)
    F.batch_norm(x, tmp[:C], tmp[C:], None, None, True, 1.0, 1e-5)
    return tmp


def test_deferred_updates_equal_sequential_updates():
    torch.manual_seed(0)
    C = 8
    x1, x2 = torch.randn(4, C, 5, 7) * 2 + 1, torch.randn(4, C, 5, 7) * 0.5 - 3
    seq = torch.nn.BatchNorm2d(C).train()
    seq(x1)
    seq(x2)
    par = torch.nn.BatchNorm2d(C).train()
    par(x1)                                     # first pass updates in place
    with bn_act.deferred_running_stats() as entries:
        assert bn_act._deferred is entries
        entries.append((par, _batch_stats(x2, C), float(par.momentum)))   # second pass, deferred
    assert bn_act._deferred is None
    bn_act.apply_deferred(entries)
    assert entries == []
    assert torch.allclose(par.running_mean, seq.running_mean, rtol=1e-6, atol=1e-7)
    assert torch.allclose(par.running_var, seq.running_var, rtol=1e-6, atol=1e-7)
    assert int(par.num_batches_tracked) == int(seq.num_batches_tracked) == 2


def test_deferred_contexts_nest_and_restore():
    with bn_act.deferred_running_stats() as outer:
        with bn_act.deferred_running_stats() as inner:
            assert bn_act._deferred is inner
        assert bn_act._deferred is outer
    assert bn_act._deferred is None
    bn_act.apply_deferred(None)
    bn_act.apply_deferred([])
