"""One optimiser step of EdgeLine-YOLO through the product's training path (BASELINE.json configs[4]).

Mirrors what the reference's trainer does per batch (engine/trainer.py:217-273 DDP wrap with find_unused_parameters=True -- the
DGQP heads get no gradient on the uniH path; :380-395 autocast forward + loss, `scaler.scale(loss).backward()`, optimizer
step; build_optimizer :760-818 SGD momentum 0.937 nesterov): forward under autocast with the CUDA DWT / merge / gated-residual /
attention kernels under autograd, `v8DetectionLoss` (el_tal_assign + el_dfl_fwd/bwd + BCE), backward through the CUDA backward
kernels, DDP's bucketed NCCL all-reduce of the gradients, SGD step.  One process per GPU.
"""
from __future__ import annotations

import contextlib

import torch

from . import modules as M
from .detection_loss import v8DetectionLoss
from .model import EdgeLineYOLO


class TrainStep:
    def __init__(self, scale: str = "s", nc: int = 80, device="cuda", world: int = 1, local_rank: int = 0, amp: bool = True,
                 channels_last: bool = True, seed: int = 0, lr: float = 0.01):
        torch.manual_seed(seed)
        self.device, self.world, self.amp, self.channels_last = torch.device(device), world, amp, channels_last
        model = EdgeLineYOLO(scale, nc)
        with torch.no_grad():
            for m in model.modules():
                if isinstance(m, M._WaveletEnhancer):
                    m.gamma.fill_(0.5)  # the reference's init (0) switches the wavelet branch off
        model.model[-1].bias_init()
        model = model.to(self.device).train()
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
        self.model = model
        self.criterion = v8DetectionLoss(model)
        self.net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], find_unused_parameters=True) if world > 1 else model
        self.opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.937, nesterov=True, weight_decay=5e-4)
        self.n_params = sum(p.numel() for p in model.parameters())

    def synth_batch(self, batch: int, imgsz: int, rank: int = 0, boxes_per_image: int = 8):
        """Seeded synthetic batch of SURVEY 8(d) config 5: 8 boxes per image, cls ~ U{0..nc-1}, centre ~ U(.1,.9), size ~ U(.05,.4)."""
        dev, nc = self.device, self.model.nc
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        x = torch.rand(batch, 3, imgsz, imgsz, device=dev, generator=g)
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        n = batch * boxes_per_image
        cls = torch.randint(0, nc, (n, 1), device=dev, generator=g).float()
        cxy = 0.1 + 0.8 * torch.rand(n, 2, device=dev, generator=g)
        wh = 0.05 + 0.35 * torch.rand(n, 2, device=dev, generator=g)
        targets = {"batch_idx": torch.arange(batch, device=dev).repeat_interleave(boxes_per_image).float(), "cls": cls, "bboxes": torch.cat([cxy, wh], 1)}
        return x, targets

    def step(self, x, targets, sync_grads: bool = True, events=None):
        """forward + loss + backward + optimiser step.  `sync_grads=False` skips DDP's all-reduce (measurement aid: the difference to a
        normal step is the exposed cost of the collective).  `events`: optional 4 CUDA events recorded at the phase boundaries."""
        ctx = self.net.no_sync() if (self.world > 1 and not sync_grads) else contextlib.nullcontext()
        with ctx:
            if events:
                events[0].record()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                feats = self.net(x)
            if events:
                events[1].record()
            loss, items = self.criterion([f.float() for f in feats], targets)
            if events:
                events[2].record()
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
        self.opt.step()
        if events:
            events[3].record()
        return items
