"""Fused dense AdamW over flat parameter / gradient buffers (csrc/optim.cu).

The reference uses torch.optim.AdamW or apex FusedAdam with betas (0.9, 0.99), eps 1e-15, per-group weight decay
(train_permuto_sdf.py:293-304). Here every parameter of every group is re-homed into ONE flat fp32 buffer (and its
gradient into one flat gradient buffer, so autograd accumulates straight into it): a step is one streaming kernel
per group (28 B/param, gradient zeroing folded in), and data-parallel training needs a single all-reduce of
`flat_grad`. `param_groups` keeps the torch layout (lr / weight_decay / name can be edited by schedulers)."""
import torch

from ._lib import call


class FusedAdamW:
    def __init__(self, groups, betas=(0.9, 0.99), eps=1e-15, lr=1e-3, weight_decay=0.0):
        self.betas, self.eps = betas, eps
        self.param_groups = []
        total = 0
        layout = []
        for g in groups:
            ps = [p for p in g["params"] if p.requires_grad]
            n = sum(p.numel() for p in ps)
            n_pad = (n + 3) // 4 * 4                     # keep every group 16-byte aligned
            layout.append((ps, total, n, n_pad))
            total += n_pad
        dev = layout[0][0][0].device
        self.flat_param = torch.zeros(total, device=dev)
        self.flat_grad = torch.zeros(total, device=dev)
        self.exp_avg = torch.zeros(total, device=dev)
        self.exp_avg_sq = torch.zeros(total, device=dev)
        self.step_count = 0
        self.device_step = False        # True: the step count is kept in self.step_dev (device) for CUDA-graph replay
        self.step_dev = None
        for g, (ps, off, n, n_pad) in zip(groups, layout):
            o = off
            for p in ps:
                k = p.numel()
                self.flat_param[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.flat_param[o:o + k].view_as(p)
                p.grad = self.flat_grad[o:o + k].view_as(p)
                o += k
            self.param_groups.append({"params": ps, "lr": g.get("lr", lr), "weight_decay": g.get("weight_decay", weight_decay),
                                      "name": g.get("name", ""), "_off": off, "_n": n_pad})

    def zero_grad(self, set_to_none=False):
        # gradients are zeroed inside the step kernel; an explicit call is only needed before the first backward
        if self.step_count == 0:
            self.flat_grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        self.step_count += 1
        b1, b2 = self.betas
        step_dev = None
        if self.device_step:            # CUDA-graph mode: the step counter lives (and is incremented) on the device
            if self.step_dev is None:
                self.step_dev = torch.full((1,), self.step_count - 1, dtype=torch.int32, device=self.flat_param.device)
            self.step_dev.add_(1)
            step_dev = self.step_dev
        for g in self.param_groups:
            off, n = g["_off"], g["_n"]
            if n == 0:
                continue
            call("psdf_adamw_step", n, self.flat_param[off:off + n], self.flat_grad[off:off + n], self.exp_avg[off:off + n],
                 self.exp_avg_sq[off:off + n], float(g["lr"]), b1, b2, self.eps, float(g["weight_decay"]), self.step_count, step_dev,
                 float(grad_scale), 1)

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}
