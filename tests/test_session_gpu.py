"""`Model.inference_session`: frozen-weight inference from / to pinned host buffers (H2D + one graph replay + D2H)."""
import pytest
import torch

from oracle import torch_port as tp
from tests.helpers import assert_close, build_model, cases, golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_session_equals_forward_and_tracks_staleness():
    c = cases("forward")["cfg1_shape"]
    g = golden("cfg1_shape")
    m = build_model(c, DEV).eval()
    x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
    xh = x.pin_memory()
    out = torch.empty(c["B"], c["H"], c["N"]).pin_memory()
    sess = m.inference_session(c["B"])
    sess(xh, out)
    torch.cuda.synchronize()
    assert_close(out, g["forecast"], msg="session forecast vs reference golden")
    with torch.no_grad():
        f, _ = m(x.to(DEV))
    assert torch.equal(out, f.reshape(out.shape).cpu())          # same captured graph, same bits
    # another batch size through the plain forward re-allocates the model's workspace: the session keeps its own alive
    x2, _ = tp.synthetic_batch(3, c["N"], c["W"], c["H"], seed=7)
    with torch.no_grad():
        m(x2.to(DEV))
        m(x2.to(DEV))
    out2 = torch.empty_like(out)
    sess(xh, out2)
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    # a parameter update makes the session stale; refresh() picks the new weights up
    assert not sess.stale()
    with torch.no_grad():
        m.fc[0].weight.mul_(1.5)
    assert sess.stale()
    sess.refresh()
    assert not sess.stale()
    sess(xh, out2)
    torch.cuda.synchronize()
    with torch.no_grad():
        f3, _ = m(x.to(DEV))
    assert torch.equal(out2, f3.reshape(out2.shape).cpu())
    assert not torch.equal(out2, out)
    # device-resident use: no output buffer -> the static device forecast
    d = sess(x.to(DEV))
    assert torch.equal(d.cpu(), out2)


def test_session_needs_eval_mode():
    c = cases("forward")["tiny_taps"]
    m = build_model(c, DEV).train()
    with pytest.raises(RuntimeError):
        m.inference_session(c["B"])
