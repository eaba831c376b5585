"""torch-CPU float32 restatement of the training step ("fast mode": batched over contexts, same math as oracle/npe.py).

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).  Used by bench.py's cpu_baseline and
`--impl reference` legs: the reference itself (Lua/Torch7) cannot run here, so this port -- one feval_train +
optimizer step per call (main.lua:247-269,301-302) with torch's multi-threaded CPU kernels -- stands in for
`th main.lua` on the box's host cores.  Gradient scaling follows npe.lua:301-320.
"""
import numpy as np
import torch

from . import npe


class TorchNPE:
    def __init__(self, params, threads=None):
        if threads:
            torch.set_num_threads(threads)
        ents, n = npe.param_layout()
        self.flat = torch.tensor(np.asarray(params, np.float32))
        self.flat.requires_grad_(True)
        self.ents = ents
        self.opt_state = None

    def views(self):
        return {n: self.flat[o:o + int(np.prod(s))].reshape(s) for n, o, s in self.ents}

    def forward(self, this_past, ctx, mask):
        p = self.views()
        m = mask[:, :, None]
        et = torch.relu((this_past[:, None, :] * m) @ p["enc.this.W"].T)
        ec = torch.relu((ctx * m) @ p["enc.ctx.W"].T)
        h = torch.cat([et, ec], 2)
        for l in range(1, 6):
            h = torch.relu(h @ p[f"pair.L{l}.W"].T)
        u = torch.cat([h.sum(1), this_past], 1)
        for i in range(1, 6):
            u = u @ p[f"dec.L{i}.W"].T + p[f"dec.L{i}.b"]
            if i < 5:
                u = torch.relu(u)
        return u

    def train_step(self, this, context, y, lr=3e-4, opt="adam"):
        """One feval_train + optimizer step on a reference batch tuple (numpy float32 arrays)."""
        B = this.shape[0]
        mask = torch.from_numpy(npe.neighbor_mask(this, context, check_focus=False))
        tp = torch.from_numpy(this.reshape(B, -1)); cx = torch.from_numpy(context.reshape(B, context.shape[1], -1))
        yy = torch.from_numpy(y.reshape(B, -1))
        if self.flat.grad is not None:
            self.flat.grad = None
        u = self.forward(tp, cx, mask)
        sv = ((u[:, 2:4] - yy[:, 2:4]) ** 2).sum(); sw = ((u[:, 5] - yy[:, 5]) ** 2).sum()
        obj = sv / (2 * B) + sw / B
        obj.backward()
        loss = float((sv + sw).detach() / (3 * B))
        g = self.flat.grad
        with torch.no_grad():
            if opt == "adam":
                if self.opt_state is None:
                    self.opt_state = {"t": 0, "m": torch.zeros_like(g), "v": torch.zeros_like(g)}
                s = self.opt_state
                s["t"] += 1
                s["m"].mul_(0.9).add_(g, alpha=0.1)
                s["v"].mul_(0.999).addcmul_(g, g, value=0.001)
                step = lr * np.sqrt(1 - 0.999 ** s["t"]) / (1 - 0.9 ** s["t"])
                self.flat.addcdiv_(s["m"], s["v"].sqrt().add_(1e-8), value=-step)
            else:
                if self.opt_state is None:
                    self.opt_state = {"m": torch.zeros_like(g)}
                s = self.opt_state
                s["m"].mul_(0.99).addcmul_(g, g, value=0.01)
                self.flat.addcdiv_(g, s["m"].sqrt().add_(1e-8), value=-lr)
        return loss
