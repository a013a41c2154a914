"""GPU: the host-buffer pipeline (per-sample H2D | compute | D2H on three streams) returns exactly
what the plain autograd call returns."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fac_host_pipeline_matches_direct_call():
    from ebfi_be_b200 import kernelconv2d
    from ebfi_be_b200.host_pipeline import HostPipeline
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, C, K, H, W = 3, 8, 5, 40, 48          # 8 channels: exercises the 4-way channel split
    x = torch.randn(B, C, H + K - 1, W + K - 1).pin_memory()
    ker = torch.randn(B, C * K * K, H, W).pin_memory()
    go = torch.randn(B, C, H, W).pin_memory()
    out, gi, gk = (torch.empty_like(t).pin_memory() for t in (go, x, ker))
    pipe = HostPipeline(dev)
    for _ in range(2):      # second pass reuses the device buffers
        pipe.fac_forward_backward(x, ker, go, K, out, gi, gk)
    xg, kg = x.to(dev).requires_grad_(), ker.to(dev).requires_grad_()
    o = kernelconv2d.KernelConv2DFunction.apply(xg, kg, K)
    o.backward(go.to(dev))
    assert torch.equal(out, o.detach().cpu()) and torch.equal(gi, xg.grad.cpu()) and torch.equal(gk, kg.grad.cpu())


def test_dcn_host_pipeline_matches_direct_call():
    from ebfi_be_b200 import dcn_v2
    from ebfi_be_b200.host_pipeline import HostPipeline
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    B, C, H, W, dg = 2, 64, 24, 24, 8
    host = [torch.randn(B, C, H, W), 2 * torch.randn(B, 2 * dg * 9, H, W), torch.sigmoid(torch.randn(B, dg * 9, H, W)),
            torch.randn(64, C, 3, 3) / 24, torch.randn(64), torch.randn(B, 64, H, W)]
    host = [t.pin_memory() for t in host]
    names = ["out", "grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"]
    res = {n: torch.empty_like(t).pin_memory() for n, t in zip(names, [host[5]] + host[:5])}
    s_out, keep = HostPipeline(dev).dcn_forward_backward(*host, 1, 1, 1, dg, res)
    s_out.synchronize()
    leaves = [t.to(dev).requires_grad_() for t in host[:5]]
    o = dcn_v2.dcn_v2_conv(*leaves, 1, 1, 1, dg)
    o.backward(host[5].to(dev))
    assert torch.equal(res["out"], o.detach().cpu())
    for n, l in zip(names[2:], leaves[1:]):                 # deterministic gradients: bit equality
        assert torch.equal(res[n], l.grad.cpu()), n
    assert torch.allclose(res["grad_input"], leaves[0].grad.cpu(), rtol=0, atol=1e-4 * float(leaves[0].grad.abs().max()))
