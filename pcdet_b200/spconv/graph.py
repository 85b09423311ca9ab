"""CUDA-graph capture of an UNMODIFIED spconv module tree through the static-shape mode (SparseConvTensor.n_dev).

The reference drives its backbones through `spconv.SparseSequential` modules that synchronise with the host at every
rulebook build (spconv v1.0 `indiceNum.to(cpu)`).  With capacity-sized tensors and device-side row
counts nothing in the forward pass needs the host, so the whole pass is one graph:

    runner = GraphedSparseModule(net, capacity=B * 40000, channels=4, spatial_shape=[41, 1600, 1408], batch_size=B)
    out = runner(voxel_features, coordinates)          # copies the batch into the static buffers and replays
    runner.check_overflow()                             # optional, synchronises: a level outgrew its capacity

`out` is what `net(...)` returned at capture time (tensors keep their addresses; rows past the level's count are padding).
pcdet_b200/pipeline.py (SECOND) and pcdet_b200/parta2.py (Part-A^2) are hand-scheduled pipelines on the same kernels; this
is the drop-in route for any other module tree built from pcdet_b200.spconv."""
from __future__ import annotations

from typing import Sequence

import torch

from .tensor import SparseConvTensor


class GraphedSparseModule:
    def __init__(self, net: torch.nn.Module, capacity: int, channels: int, spatial_shape: Sequence[int], batch_size: int,
                 dtype=torch.float32, device="cuda", warmup: int = 2):
        self.net = net.eval()
        self.dev = torch.device(device)
        self.capacity, self.batch_size, self.spatial_shape = int(capacity), int(batch_size), [int(s) for s in spatial_shape]
        self.features = torch.zeros((self.capacity, channels), dtype=dtype, device=self.dev)
        self.indices = torch.zeros((self.capacity, 4), dtype=torch.int32, device=self.dev)
        self.n_dev = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.graph = None
        self.out = None
        self._overflow = []
        self._warmup = warmup

    def _forward(self):
        x = SparseConvTensor(self.features, self.indices, self.spatial_shape, self.batch_size, n_dev=self.n_dev)
        with torch.no_grad():
            out = self.net(x)
        self._overflow = x.indice_dict.get("__overflow__", [])
        return out

    def _load(self, features: torch.Tensor, indices: torch.Tensor):
        n = features.shape[0]
        assert n <= self.capacity and indices.shape[0] == n, f"{n} rows exceed the capacity {self.capacity}"
        self.features[:n].copy_(features, non_blocking=True)
        self.indices[:n].copy_(indices, non_blocking=True)
        self.n_dev.fill_(n)

    def capture(self, features: torch.Tensor, indices: torch.Tensor):
        """Warm up on this batch (weight casts, workspaces) and capture; called by the first __call__ if not before."""
        self._load(features, indices)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self._warmup):
                self._forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._forward()
        self.graph.replay()             # capturing does not execute: run the graph once so that `out` holds this batch
        return self.out

    def __call__(self, features: torch.Tensor, indices: torch.Tensor):
        if self.graph is None:
            return self.capture(features, indices)
        self._load(features, indices)
        self.graph.replay()
        return self.out

    def check_overflow(self):
        """Synchronises.  Raises if a strided level held more sites than its capacity (ops.CAPACITY_GROWTH) in the last run."""
        if self._overflow and int(torch.cat(self._overflow).sum()) != 0:
            from .._lib import PcdbError
            raise PcdbError("a sparse level exceeded its static-shape capacity: sites were dropped; raise `capacity` or "
                            "pcdet_b200.spconv.ops.CAPACITY_GROWTH")
