"""Host-buffer entry points: run the alignment ops on tensors that live in (pinned) host memory.

The reference's operators only take CUDA tensors (`KernelConv2D.py:38-39` raises on CPU input);
data that starts on the host — e.g. what a DataLoader hands over — has to cross PCIe both ways.
For FAC that is 1.8 GB in and 1.8 GB out per step at the benchmark shape, ~25x the kernel time,
so the copies are what matters: these helpers split the batch into samples and run
H2D(sample i+1) | forward+backward(sample i) | D2H(sample i-1) on three streams, which keeps both
PCIe directions busy at once instead of serialising them.

All compute goes through the public autograd Functions (`KernelConv2DFunction`, `dcn_v2_conv`).
"""
import torch

from .dcn_v2 import dcn_v2_conv
from .kernelconv2d import KernelConv2DFunction


class HostPipeline:
    """Three CUDA streams + reusable per-sample device buffers for one device."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._bufs = {}
        self._dcn_busy = []          # events of the previous dcn_forward_backward still using the cached buffers

    def _buf(self, key, like):
        b = self._bufs.get(key)
        if b is None or b.shape != like.shape or b.dtype != like.dtype:
            b = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            self._bufs[key] = b
        return b

    def fac_forward_backward(self, input_pad, kernel, grad_output, K, out, grad_input, grad_kernel,
                             channel_splits=4):
        """input_pad / kernel / grad_output: pinned host tensors (B, ...); out / grad_input / grad_kernel:
        pinned host tensors that receive the results. Returns after every result is on the host.
        The op is independent per (sample, channel), so each sample is further cut into
        `channel_splits` channel blocks (free views of the contiguous tensors): more, smaller pipeline
        items shorten the un-overlapped first H2D / last D2H."""
        C = input_pad.shape[1]
        ns = channel_splits if (channel_splits > 1 and C % channel_splits == 0) else 1
        if ns > 1:
            def cut(t, per_c):
                return t.view(t.shape[0] * ns, (C // ns) * per_c, *t.shape[2:])
            input_pad, grad_output, out, grad_input = (cut(t, 1) for t in (input_pad, grad_output, out, grad_input))
            kernel, grad_kernel = cut(kernel, K * K), cut(grad_kernel, K * K)
        B = input_pad.shape[0]
        comp = torch.cuda.current_stream(self.device)
        keep = []
        for i in range(B):
            slot = i & 1                                     # two sets of input buffers: copy i+1 while i computes
            with torch.cuda.stream(self.s_in):
                # the buffers of this slot were last read by the compute of sample i-2
                if i >= 2:
                    self.s_in.wait_event(keep[i - 2][0])
                x = self._buf(("x", slot), input_pad[i:i + 1]); x.copy_(input_pad[i:i + 1], non_blocking=True)
                k = self._buf(("k", slot), kernel[i:i + 1]); k.copy_(kernel[i:i + 1], non_blocking=True)
                g = self._buf(("g", slot), grad_output[i:i + 1]); g.copy_(grad_output[i:i + 1], non_blocking=True)
                ready = torch.cuda.Event(); ready.record(self.s_in)
            comp.wait_event(ready)
            xr, kr = x.detach().requires_grad_(), k.detach().requires_grad_()
            o = KernelConv2DFunction.apply(xr, kr, K)
            o.backward(g)
            done = torch.cuda.Event(); done.record(comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                out[i:i + 1].copy_(o.detach(), non_blocking=True)
                grad_input[i:i + 1].copy_(xr.grad, non_blocking=True)
                grad_kernel[i:i + 1].copy_(kr.grad, non_blocking=True)
            keep.append((done, o, xr, kr))                   # keep device results alive until the D2H has run
        self.s_out.synchronize()
        return out, grad_input, grad_kernel

    def dcn_forward_backward(self, input, offset, mask, weight, bias, grad_output, stride, padding, dilation,
                             deformable_groups, results):
        """Host tensors in, `results` = dict of pinned host tensors for out, grad_input, grad_offset, grad_mask,
        grad_weight, grad_bias. Single shot (the DCN tensors are 20x smaller than FAC's)."""
        comp = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.s_in):
            # the cached device buffers are still read by the previous call's kernels / D2H copies until its `done`
            # and copy-out events: order this call's H2D after them
            for e in self._dcn_busy:
                self.s_in.wait_event(e)
            dev = [self._buf(("d", j), t) for j, t in enumerate((input, offset, mask, weight, bias, grad_output))]
            for d_, h_ in zip(dev, (input, offset, mask, weight, bias, grad_output)):
                d_.copy_(h_, non_blocking=True)
            ready = torch.cuda.Event(); ready.record(self.s_in)
        comp.wait_event(ready)
        leaves = [t.detach().requires_grad_() for t in dev[:5]]
        o = dcn_v2_conv(*leaves, stride, padding, dilation, deformable_groups)
        o.backward(dev[5])
        done = torch.cuda.Event(); done.record(comp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            for name, t in zip(("out", "grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"),
                               [o.detach()] + [l.grad for l in leaves]):
                results[name].copy_(t, non_blocking=True)
            copied = torch.cuda.Event(); copied.record(self.s_out)
        self._dcn_busy = [done, copied]
        keep = (o, leaves)
        return self.s_out, keep                              # caller synchronises (lets FAC overlap the D2H)
