"""Host logic of the drop-in modules that needs no GPU: the per-instance caches of device state (weight-shadow buffers, native
model struct, fp16 status word with its stream / event) are transient -- copy.deepcopy(model) and torch.save(model) work on a
model that has already run, and the copy rebuilds them on demand; precision names are validated."""
import copy
import io

import pytest
import torch

from molclr_b200 import GCN, GINet
from molclr_b200.ginet import PRECISIONS, _check_precision
from molclr_b200.ginet_finetune import GINet as FineTuneGINet


def _used(m):
    """Put un-copyable stand-ins where a forward on the GPU leaves its caches."""
    m.__dict__["_native_struct"] = ("key", object(), [lambda: 0])
    m.__dict__["_fp16_status"] = {"dev": torch.zeros(1), "ev": (lambda: 0)}
    m.__dict__["_gemm_weight_specs"] = {1: [lambda: 0]}
    m._rounded._launch = lambda: 0
    return m


@pytest.mark.parametrize("make", [lambda: GINet(2, 32, 16), lambda: GCN(2, 32, 16), lambda: FineTuneGINet("classification", 2, 32, 16)])
def test_device_caches_are_not_copied_or_pickled(make):
    m = _used(make())
    m.precision, m.deterministic = "tf32x3", True
    c = copy.deepcopy(m)
    assert not any(k in c.__dict__ for k in m._TRANSIENT) and c._rounded._launch is None
    assert (c.precision, c.deterministic) == ("tf32x3", True)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    assert "_native_struct" in m.__dict__                      # the original keeps its caches
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert list(r.state_dict().keys()) == list(m.state_dict().keys()) and r.precision == "tf32x3"


def test_precision_names():
    m = GINet(2, 32, 16)
    assert m.precision in PRECISIONS
    assert [_check_precision(type("M", (), {"precision": p})()) for p in ("fp16x3", "tf32x3", "tf32")] == [2, 1, 0]
    m.precision = "bf16"
    with pytest.raises(ValueError):
        _check_precision(m)
