"""Fused dense AdamW over flat parameter / gradient buffers (csrc/optim.cu).

The reference uses torch.optim.AdamW or apex FusedAdam with betas (0.9, 0.99), eps 1e-15, per-group weight decay
(train_permuto_sdf.py:293-304). Here every parameter of every group is re-homed into ONE flat fp32 buffer (and its
gradient into one flat gradient buffer, so autograd accumulates straight into it): a step is one streaming kernel
per group (28 B/param, gradient zeroing folded in), and data-parallel training needs a single all-reduce of
`flat_grad`. `param_groups` keeps the torch layout (lr / weight_decay / name can be edited by schedulers: in CUDA-graph mode
the kernel reads both from a small device array that `sync_hyper()` refreshes before every replay, so an edit takes effect
without re-capturing).

Gradient semantics: the step kernel leaves `flat_grad` zeroed; `zero_grad()` really zeroes whenever a backward may have run
without a following step (same behaviour as torch.optim.AdamW + zero_grad in the reference loop)."""
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
        self._clean = True              # flat_grad is known to be all zero (fresh, or just swept by the step kernel)
        self.hyper_dev = None           # [groups, 2] device copy of (lr, weight_decay), read by the kernel in device_step mode
        self._hyper_host = None
        self._hyper_cached = None
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
        # the step kernel zeroes the gradients it consumed; a backward that was not followed by a step leaves them dirty
        if not self._clean:
            self.flat_grad.zero_()
        self._clean = False             # the caller is about to run a backward

    def check_aliasing(self):
        """parameters must still be views into flat_param (a later model.to() / .cuda() re-homes p.data and breaks the step)"""
        for g in self.param_groups:
            o = g["_off"]
            for p in g["params"]:
                if p.data_ptr() != self.flat_param.data_ptr() + 4 * o or p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                    raise RuntimeError("FusedAdamW: parameter storage was re-assigned after the optimizer was built "
                                       "(model.to()/cuda() after construction?); rebuild the optimizer")
                o += p.numel()

    def sync_hyper(self):
        """device copy of (lr, weight_decay) per group; call outside graph capture, before every replay (cheap: host compare,
        one 8 B x groups copy only when a scheduler changed something)"""
        cur = tuple((float(g["lr"]), float(g["weight_decay"])) for g in self.param_groups)
        if self.hyper_dev is None:
            self.hyper_dev = torch.zeros(len(cur), 2, device=self.flat_param.device)
            self._hyper_host = torch.zeros(len(cur), 2).pin_memory() if torch.cuda.is_available() else torch.zeros(len(cur), 2)
        if cur != self._hyper_cached:
            self._hyper_host.copy_(torch.tensor(cur, dtype=torch.float32).view(-1, 2))
            self.hyper_dev.copy_(self._hyper_host, non_blocking=True)
            self._hyper_cached = cur

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
            if self.hyper_dev is None:
                self.sync_hyper()
        for gi, g in enumerate(self.param_groups):
            off, n = g["_off"], g["_n"]
            if n == 0:
                continue
            hyper = self.hyper_dev[gi] if self.device_step else None
            call("psdf_adamw_step", n, self.flat_param[off:off + n], self.flat_grad[off:off + n], self.exp_avg[off:off + n],
                 self.exp_avg_sq[off:off + n], float(g["lr"]), b1, b2, self.eps, float(g["weight_decay"]), self.step_count, step_dev,
                 hyper, float(grad_scale), 1)
        self._clean = True

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        """inverse of state_dict (moments, step count, per-group lr / weight_decay)"""
        if sd["exp_avg"].numel() != self.exp_avg.numel():
            raise RuntimeError("FusedAdamW.load_state_dict: flat size mismatch")
        self.exp_avg.copy_(sd["exp_avg"].to(self.exp_avg.device))
        self.exp_avg_sq.copy_(sd["exp_avg_sq"].to(self.exp_avg.device))
        self.step_count = int(sd["step"])
        if self.step_dev is not None:
            self.step_dev.fill_(self.step_count)
        for g, sg in zip(self.param_groups, sd["groups"]):
            g["lr"], g["weight_decay"] = sg["lr"], sg["weight_decay"]
        self._hyper_cached = None
