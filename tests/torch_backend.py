"""Test-only arithmetic for grasp_b200.fused.FusedLlama: every primitive in plain torch, and every backward
taken from autograd (NOT from the formulas the CUDA kernels implement), so that running FusedLlama on it checks
the orchestration -- which tensors are saved, where gradients are summed, where the backward stops --
independently of the kernels.  Never imported by the product."""
import torch


class TorchBackend:
    def prep(self, x):
        return x

    def wprep(self, w, tag=None):
        return w.detach()

    def drop_weights(self, tag=None):
        pass

    def mm_nt(self, x, w, out=None, beta=0.0):
        y = x @ w.t()
        if out is None:
            return y
        out.mul_(beta).add_(y)
        return out

    def mm_nn(self, dy, w, out=None, beta=0.0):
        y = dy @ w
        if out is None:
            return y
        out.mul_(beta).add_(y)
        return out

    def harvest(self, dy, x, G):
        g = dy.t() @ x
        return g if G is None else G.add_(g)

    @staticmethod
    def _rms(x, w, eps):
        return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))

    def rmsnorm_fwd(self, x, w, eps):
        self._eps = eps
        rstd = torch.rsqrt(x.pow(2).mean(-1) + eps)
        return self._rms(x, w, eps), rstd

    def rmsnorm_fwd_op(self, x, w, eps, keep_y):
        y, rstd = self.rmsnorm_fwd(x, w, eps)
        return (y if keep_y else None), rstd, y

    def swiglu_fwd_op(self, g, u, keep_h):
        h = self.swiglu_fwd(g, u)
        return (h if keep_h else None), h

    def swiglu_bwd_op(self, dh, g, u, keep_grads):
        dg, du = self.swiglu_bwd(dh, g, u)
        return (dg if keep_grads else None), (du if keep_grads else None), dg, du

    def rmsnorm_bwd(self, dy, x, w, rstd, add=None):
        xl = x.detach().requires_grad_(True)
        with torch.enable_grad():
            y = self._rms(xl, w.detach(), self._eps)
        (dx,) = torch.autograd.grad(y, xl, dy)
        return dx if add is None else dx + add

    @staticmethod
    def _rope(x, seq, heads, hd, cos, sin):
        t = x.shape[0]
        x4 = x.view(t // seq, seq, heads, hd)
        c, s = cos.unsqueeze(2), sin.unsqueeze(2)
        rot = torch.cat((-x4[..., hd // 2:], x4[..., :hd // 2]), dim=-1)
        return (x4 * c + rot * s).reshape(t, heads * hd)

    def rope_(self, x, seq, heads, hd, cos, sin, inverse=False):
        if not inverse:
            x.copy_(self._rope(x.clone(), seq, heads, hd, cos, sin))
            return x
        xl = torch.zeros_like(x).requires_grad_(True)
        with torch.enable_grad():
            y = self._rope(xl, seq, heads, hd, cos, sin)
        (dx,) = torch.autograd.grad(y, xl, x.clone())
        x.copy_(dx)
        return x

    def swiglu_fwd(self, g, u):
        return torch.nn.functional.silu(g) * u

    def swiglu_bwd(self, dh, g, u):
        gl, ul = g.detach().clone().requires_grad_(True), u.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            h = torch.nn.functional.silu(gl) * ul
        dg, du = torch.autograd.grad(h, (gl, ul), dh)
        g.copy_(dg)
        u.copy_(du)
        return g, u

    def ce_loss_bwd_(self, logits, labels, coef):
        ll = logits.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            per = torch.nn.functional.cross_entropy(ll, labels, reduction="none", ignore_index=-100)
            loss = per * coef
            total = loss.sum()
        (dl,) = torch.autograd.grad(total, ll)
        logits.copy_(dl)
        return loss.detach()
