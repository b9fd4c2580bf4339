"""Second, independent restatement of the reference graph on torch-CPU with AUTOGRAD.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Two uses:

* fp64: cross-validates the hand-derived backward of oracle/vae_assoc_oracle.py
  (tests/test_oracle_grads.py) -- the graph is written forward-only, following
  vae_assoc.py:78-119,163-304,306-371, and differentiated by torch.
* fp32, all host threads: the "CPU restatement (torch), not TensorFlow" baseline
  timed by bench.py at the reference's op granularity:
  separate matmul / add / activation ops, autograd backward, TF-formula Adam,
  one host read of the scalar cost per step (vae_assoc.py:383-386).
"""
import itertools

import numpy as np
import torch
import torch.nn.functional as F

from . import vae_assoc_oracle as vo


def _same_pad(n, k, s):
    o = -(-n // s)
    tot = max((o - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def tf_conv2d(x, w, s, padding):
    """x NHWC, w [kh,kw,cin,cout] (TensorFlow layout) -> NHWC."""
    xn = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        pb, pa = _same_pad(x.shape[1], w.shape[0], s)
        xn = F.pad(xn, (pb, pa, pb, pa))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=s)
    return y.permute(0, 2, 3, 1)


def tf_conv2d_transpose(y, w, s, padding):
    """y NHWC [B,h,h,in_depth], w [kh,kw,out_depth,in_depth] -> NHWC [B,H,H,out_depth] (deconv.py:107)."""
    k = w.shape[0]
    full = F.conv_transpose2d(y.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=s)
    if padding == "SAME":
        H = y.shape[1] * s
        pb, _ = _same_pad(H, k, s)
        full = full[:, :, pb:pb + H, pb:pb + H]
    return full.permute(0, 2, 3, 1)


class TorchAssocVAE(object):
    def __init__(self, archs, binary, transfer_fct, weights, assoc_lambda, learning_rate, batch_size,
                 params, dtype=torch.float64):
        self.archs, self.binary, self.weights = archs, list(binary), list(weights)
        self.f = {"relu": torch.relu, "softplus": F.softplus}[transfer_fct]
        self.lam, self.lr, self.batch_size = assoc_lambda, learning_rate, batch_size
        self.n_z = archs[0]["n_z"]
        self.dtype = dtype
        self.params = [[torch.tensor(np.asarray(p), dtype=dtype, requires_grad=True) for p in ps] for ps in params]
        self.flat = [p for ps in self.params for p in ps]
        self.m = [torch.zeros_like(p) for p in self.flat]
        self.v = [torch.zeros_like(p) for p in self.flat]
        self.t = 0

    def encode(self, m, x):
        na, P, f = self.archs[m], self.params[m], self.f
        if na["hidden_conv"]:
            s0 = vo.conv_geometry(na)[0]
            h = tf_conv2d(x.reshape(-1, s0, s0, 1), P[0], 2, "SAME")
            h = tf_conv2d(h, P[1], 2, "SAME")
            h = tf_conv2d(h, P[2], 1, "VALID")
            h2 = h.reshape(h.shape[0], -1)
            return torch.add(torch.matmul(h2, P[3]), P[4]), torch.add(torch.matmul(h2, P[5]), P[6])
        h1 = f(torch.add(torch.matmul(x, P[0]), P[1]))
        h2 = f(torch.add(torch.matmul(h1, P[2]), P[3]))
        return torch.add(torch.matmul(h2, P[4]), P[5]), torch.add(torch.matmul(h2, P[6]), P[7])

    def decode(self, m, z):
        na, P, f = self.archs[m], self.params[m], self.f
        if na["hidden_conv"]:
            h = torch.sigmoid(tf_conv2d_transpose(z.reshape(-1, 1, 1, self.n_z), P[7], 1, "VALID") + P[8])
            h = torch.sigmoid(tf_conv2d_transpose(h, P[9], 1, "VALID") + P[10])
            h = torch.sigmoid(tf_conv2d_transpose(h, P[11], 2, "SAME") + P[12])
            h = torch.sigmoid(tf_conv2d_transpose(h, P[13], 2, "SAME") + P[14])
            g2 = h.reshape(h.shape[0], -1)
            return torch.sigmoid(torch.add(torch.matmul(g2, P[15]), P[16]))
        g1 = f(torch.add(torch.matmul(z, P[8]), P[9]))
        g2 = f(torch.add(torch.matmul(g1, P[10]), P[11]))
        a = torch.add(torch.matmul(g2, P[12]), P[13])
        return torch.sigmoid(a) if self.binary[m] else a

    def cost(self, X, eps):
        M = len(self.archs)
        mus, lvs, costs = [], [], []
        for m in range(M):
            x = X[m]
            mu, lv = self.encode(m, x)
            z = torch.add(mu, torch.mul(torch.sqrt(torch.exp(lv)), eps))
            xh = self.decode(m, z)
            if self.binary[m]:
                rec = -torch.sum(x * torch.log(1e-3 + xh) + (1 - x) * torch.log(1e-3 + 1 - xh), 1)
            else:
                rec = 0.5 * torch.sum((x - xh) ** 2)
            lat = -0.5 * torch.sum(1 + lv - mu ** 2 - torch.exp(lv), 1)
            costs.append(torch.mean(rec + lat) * self.weights[m])
            mus.append(mu); lvs.append(lv)
        assoc = []
        for i, j in itertools.combinations(range(M), 2):
            a = torch.sum(0.5 * (lvs[j].sum(1) - lvs[i].sum(1) - self.n_z + torch.exp(lvs[i] - lvs[j]).sum(1)
                                 + ((mus[j] - mus[i]) ** 2 * torch.exp(-lvs[j])).sum(1)))
            a = a + torch.sum(0.5 * (lvs[i].sum(1) - lvs[j].sum(1) - self.n_z + torch.exp(lvs[j] - lvs[i]).sum(1)
                                     + ((mus[i] - mus[j]) ** 2 * torch.exp(-lvs[i])).sum(1)))
            assoc.append(a)
        c = sum(costs)
        if assoc:
            c = c + self.lam * sum(assoc)
        return c

    def cost_and_grads(self, X, eps):
        X = [torch.as_tensor(np.asarray(x), dtype=self.dtype) for x in X]
        eps = torch.as_tensor(np.asarray(eps), dtype=self.dtype)
        for p in self.flat:
            p.grad = None
        c = self.cost(X, eps)
        c.backward()
        grads, k = [], 0
        for ps in self.params:
            grads.append([self.flat[k + i].grad.detach().numpy().copy() for i in range(len(ps))])
            k += len(ps)
        return float(c.detach()), grads

    @torch.no_grad()
    def _adam(self):
        self.t += 1
        b1, b2 = vo.ADAM_BETA1, vo.ADAM_BETA2
        lr_t = self.lr * np.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
        for p, m, v in zip(self.flat, self.m, self.v):
            g = p.grad
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            p.addcdiv_(m, v.sqrt().add_(vo.ADAM_EPS), value=-lr_t)

    def partial_fit(self, X, eps=None):
        """One step at the reference's granularity; returns a python float (host sync, vae_assoc.py:383-386)."""
        if eps is None:
            eps = torch.randn(self.batch_size, self.n_z, dtype=self.dtype)
        for p in self.flat:
            p.grad = None
        c = self.cost(X, eps)
        c.backward()
        self._adam()
        return float(c.detach())
