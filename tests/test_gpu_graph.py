"""The whole loss step -- forward and backward -- captured as one CUDA graph and replayed must give the eager step's loss
and feature gradients bit for bit (validated on a B200 in round 2, profiles/r2/r2a_staged_tests_n1.log)."""
import pytest
import torch

from conftest import has_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["clip", "siglip"])
@pytest.mark.parametrize("n,d", [(1024, 512), (4096, 768)])
def test_captured_step_replays_the_eager_step(kind, n, d):
    if not has_b200():
        pytest.fail("needs a B200 (sm_100a)")
    import mrclip_b200
    from mrclip_b200 import ClipLoss, SigLipLoss
    mrclip_b200.set_engine(None)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n + d)
    img = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).bfloat16().to(dev).requires_grad_(True)
    txt = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).bfloat16().to(dev).requires_grad_(True)
    scale = torch.tensor(14.285714 if kind == "clip" else 10.0, device=dev, requires_grad=True)
    bias = torch.tensor(-10.0, device=dev, requires_grad=True)
    mod = ClipLoss() if kind == "clip" else SigLipLoss()
    params = [img, txt, scale] + ([bias] if kind == "siglip" else [])

    def step():
        for p in params:
            p.grad = None
        loss = mod(img, txt, scale) if kind == "clip" else mod(img, txt, scale, bias)
        loss.backward()
        return loss

    eager_loss = step().detach().clone()
    eager = [p.grad.detach().clone() for p in params]
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side, capture_error_mode="thread_local"):
        g_loss = step()
    static = [p.grad for p in params]
    # new inputs through the static tensors: the replay must follow them
    with torch.no_grad():
        img.copy_(torch.roll(img, 1, 0))
    gr.replay()
    torch.cuda.synchronize()
    moved = g_loss.detach().clone()
    with torch.no_grad():
        img.copy_(torch.roll(img, -1, 0))
    gr.replay()
    torch.cuda.synchronize()
    assert not torch.equal(moved, eager_loss)
    assert torch.equal(g_loss.detach(), eager_loss)
    for a, b in zip(static, eager):
        if a.dim() == 2:
            assert torch.equal(a, b)                       # feature gradients: fixed reduction order
        else:
            torch.testing.assert_close(a, b, rtol=1e-5, atol=0)   # scalar gradients go through float atomics
