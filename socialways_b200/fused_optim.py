"""Adam on one flat buffer, fused with the gradient all-reduce of the sharded step.

`FlatAdam` is a drop-in for the two `torch.optim.Adam` objects of train.py:381,385 (same hyper-parameters, same update
rule, `state_dict()` in torch's layout so the checkpoints of train.py:653-663 interoperate):

* the parameters are re-homed as views into ONE flat fp32 buffer, their `.grad`s as views into a flat gradient buffer
  (autograd accumulates into existing `.grad`s in place, so the views survive backward());
* `step()` is ONE kernel (`sw_adam_flat`), or -- with a process group -- ONE kernel that also sums the gradients of
  all ranks with peer loads over NVLink from symmetric memory (`sw_allreduce_adam`): no NCCL call, nothing that
  cannot be captured into a CUDA graph;
* `zero_grad()` is one memset; gradients are never set to None (that would drop the views).
"""
import torch

from . import _lib
from .ops import _stream, sm_count


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, group=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise _lib.SocialWaysCudaError("FlatAdam needs fp32 CUDA parameters on one device (no CPU path)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.device = dev
        self.n = sum(p.numel() for p in self.params)
        self.n_pad = (self.n + 31) // 32 * 32
        self.group = group
        self.world, self.rank = 1, 0
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.flat_p = torch.zeros(self.n_pad, device=dev)
        self.exp_avg = torch.zeros(self.n_pad, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_pad, device=dev)
        self.step_t = torch.zeros(2, device=dev)            # [step count, scratch word of the kernel]
        self._symm = None
        if self.world > 1:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            # [gradient n_pad floats | ready[world] | done[world]] in peer-mapped memory
            self._buf = symm_mem.empty(self.n_pad + 2 * self.world + 32, dtype=torch.float32, device=dev)
            self._buf.zero_()
            self._symm = symm_mem.rendezvous(self._buf, group.group_name if hasattr(group, "group_name") else group)
            self.flat_g = self._buf[:self.n_pad]
            self._peer_ptrs = torch.tensor([int(x) for x in self._symm.buffer_ptrs], dtype=torch.int64, device=dev)
            self._seq = torch.zeros(4, dtype=torch.int32, device=dev)     # sequence, CTA counter, status, pad
            torch.cuda.synchronize(dev)
            dist.barrier(group)                       # every rank's flags are zero before anyone signals
        else:
            self.flat_g = torch.zeros(self.n_pad, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                p.grad = self.flat_g[off:off + k].view(p.shape)
                off += k

    # ---- torch.optim.Optimizer surface used by train.py ----
    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p, g in zip(self.params, self._grad_views()):
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g                                # somebody dropped the view (zero_grad(set_to_none=True)): restore it

    def _grad_views(self):
        off = 0
        for p in self.params:
            k = p.numel()
            yield self.flat_g[off:off + k].view(p.shape)
            off += k

    def step(self):
        for p, g in zip(self.params, self._grad_views()):
            if p.grad is None:                            # unused parameter (torch's Adam skips it; a zero gradient leaves it unchanged)
                g.zero_()
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)                           # autograd replaced the tensor: fold it back (eager mode only)
                p.grad = g
        b1, b2 = self.betas
        with torch.cuda.device(self.device):
            if self.world == 1:
                code = _lib.lib().sw_adam_flat(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                               self.exp_avg_sq.data_ptr(), self.step_t.data_ptr(), self.n, self.lr, b1, b2,
                                               self.eps, sm_count(self.device), _stream())
                _lib.check(code, "sw_adam_flat")
            else:
                code = _lib.lib().sw_allreduce_adam(self._peer_ptrs.data_ptr(), self.rank, self.world, self.n, self.n_pad,
                                                    self.flat_p.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                                    self.step_t.data_ptr(), self._seq.data_ptr(), self.lr, b1, b2, self.eps,
                                                    _stream())
                _lib.check(code, "sw_allreduce_adam")
        # the kernels write the parameters through raw pointers: advance the version counters the packed-weight caches key on
        torch.autograd.graph.increment_version(self.params)

    def check_status(self):
        """Host-side read of the all-reduce kernel's status word (synchronises): raises if a peer missed a step."""
        if self.world > 1:
            code = int(self._seq[2].item())
            if code:
                raise _lib.SocialWaysCudaError(
                    f"sw_allreduce_adam: rank {self.rank} gave up waiting for a peer (status {code}); the step was "
                    + ("not applied" if code == 1 else "applied, a peer never acknowledged"))

    def state_dict(self):
        state, off = {}, 0
        for i, p in enumerate(self.params):
            k = p.numel()
            state[i] = {"step": self.step_t[0].detach().clone().reshape(()).cpu(),
                        "exp_avg": self.exp_avg[off:off + k].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape).clone()}
            off += k
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"])
        off, step = 0, 0.0
        with torch.no_grad():
            for i, p in enumerate(self.params):
                k = p.numel()
                st = sd["state"].get(i)
                if st is not None:
                    self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    step = max(step, float(st["step"]))
                off += k
            self.step_t[0] = step
