"""SECOND backbone training step, device resident and capturable (BASELINE config 5).

What tools/train_utils/train_utils.py:12-60 does per iteration for the backbone -- forward in train mode, backward,
clip_grad_norm_, optimizer step, with DistributedDataParallel's gradient all-reduce (tools/train.py:119-122) -- as one fixed
launch sequence:

    rulebooks (pcdb_rulebook_chain, device-side counts, + the input-stationary maps of the strided layers)
    12 x [ weight image from the fp32 parameter -> conv (tcgen05) -> BatchNorm(train)+ReLU ]          bf16 activations
    dense() -> loss head (the caller's; RPN head and losses are out of scope) -> gradient of the dense map
    12 x [ BatchNorm+ReLU backward -> weight gradient (tcgen05, fp32, straight into the flat gradient buffer)
           -> input gradient (the forward kernel over the rulebook read the other way round) ]
    NCCL all-reduce of the flat gradient buffer in buckets, each issued as soon as its layers are done so that it
    overlaps the rest of the backward pass; gradient-norm clip; Adam (torch's fused, capturable kernel).

Buffers are capacities, counts stay on the device, nothing synchronises the host: the whole step is captured in one CUDA
graph (`capture()` / `replay()`), NCCL collectives included.  The module API (spconv.SparseSequential in train mode with
bf16 features) runs the same kernels through autograd, one host sync per strided rulebook.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from . import functional as F
from .backbone import BACKBONE8X_LAYERS, BackBone8x


def mean_square_loss(dense: torch.Tensor):
    """Stand-in for the detection head: loss = mean(dense^2); returns (loss, d loss / d dense)."""
    d = dense.float()
    return d.square().mean(), (d * (2.0 / d.numel())).to(dense.dtype)


class BackboneTrainStep:
    def __init__(self, net: BackBone8x, batch_size: int, sparse_shape: Sequence[int], max_voxels: int,
                 level_capacity: Optional[Sequence[int]] = None, lr: float = 1e-4, betas=(0.9, 0.99), weight_decay: float = 0.0,
                 grad_norm_clip: Optional[float] = 10.0, process_group=None, n_buckets: int = 3, sync_bn: bool = False,
                 loss_fn: Callable = mean_square_loss, device="cuda"):
        """sync_bn: BatchNorm statistics over every rank's rows (the reference's multi-GPU scripts all pass --sync_bn,
        tools/train.py:94-95): one all-reduce of 2C + 1 doubles per layer forward and of 2C backward, inside the step."""
        self.net, self.dev, self.B = net, torch.device(device), int(batch_size)
        self.shape = [int(v) for v in sparse_shape]
        n1 = int(max_voxels)
        caps = list(level_capacity) if level_capacity else [n1, int(n1 * 1.6), n1, n1 // 2, n1 // 2]
        self.caps = [max(int(c), 64) for c in caps]
        self.loss_fn, self.pg, self.clip = loss_fn, process_group, grad_norm_clip
        self.bn_pg = process_group if (sync_bn and process_group is not None) else None
        self.world = 1 if process_group is None else torch.distributed.get_world_size(process_group)
        level_of_key = {"subm1": 0, "spconv2": 1, "subm2": 1, "spconv3": 2, "subm3": 2, "spconv4": 3, "subm4": 3, "spconv_down2": 4}
        self.layers = []
        level = 0
        for (stem, kind, _ci, _co, ks, st, pd, key), (_s, conv, bn) in zip(BACKBONE8X_LAYERS, net.conv_modules()):
            assert conv.bias is None and type(bn) in (torch.nn.BatchNorm1d, torch.nn.SyncBatchNorm) and bn.momentum is not None
            out_level = level_of_key[key]
            self.layers.append(dict(stem=stem, kind=kind, key=key, conv=conv, bn=bn, ks=list(conv.kernel_size), st=list(conv.stride),
                                    pd=list(conv.padding), level_in=level, level_out=out_level,
                                    c_in=conv.in_channels, c_out=conv.out_channels))
            level = out_level
        self.convs = [dict(ksize=l["ks"], stride=l["st"], padding=l["pd"]) for l in self.layers if l["kind"] != "subm"]
        self.subm_ksizes = [next((l["ks"] for l in self.layers if l["kind"] == "subm" and l["level_in"] == lv), None) for lv in range(5)]
        # one flat fp32 gradient buffer, parameters' .grad are views of it, in BACKWARD order (so a bucket is a prefix slice
        # that is complete early)
        params: List[torch.nn.Parameter] = []
        for l in reversed(self.layers):
            params += [l["conv"].weight, l["bn"].weight, l["bn"].bias]
        self.params = params
        self.flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=self.dev)
        at = 0
        for p in params:
            p.grad = self.flat_grad[at:at + p.numel()].view_as(p)
            at += p.numel()
        # bucket boundaries (in elements) after whole layers
        per_layer = [sum(p.numel() for p in (l["conv"].weight, l["bn"].weight, l["bn"].bias)) for l in reversed(self.layers)]
        total, bounds, acc = sum(per_layer), [], 0
        for i, n in enumerate(per_layer):
            acc += n
            if len(bounds) < n_buckets - 1 and acc >= total * (len(bounds) + 1) / n_buckets:
                bounds.append((i, acc))
        bounds.append((len(per_layer) - 1, total))
        self.bucket_after_layer = {i: end for i, end in bounds}       # index in backward order -> end offset of the bucket
        self.opt = torch.optim.Adam(params, lr=lr, betas=betas, weight_decay=weight_decay, fused=True, capturable=True)
        i32 = dict(dtype=torch.int32, device=self.dev)
        self.feats = torch.zeros((self.caps[0], 16), dtype=torch.bfloat16, device=self.dev)      # 4 point features, zero-padded
        self.coords = torch.zeros((self.caps[0], 4), **i32)
        self.n0 = torch.zeros((1,), **i32)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.level_counts = None
        self.graph = None
        # the weight gradient of a layer needs grad_y only, nothing downstream needs it before the all-reduce: it trails the
        # backward chain on its own stream, beside the next layer's BatchNorm backward and input-gradient convolution
        self.wg_stream = torch.cuda.Stream(device=self.dev)
        self.rb_stream = torch.cuda.Stream(device=self.dev)
        self._rb_forked, self._maps_ready, self._inv_ready = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._gy_ready = [torch.cuda.Event() for _ in self.layers]
        self._wg_done = [torch.cuda.Event() for _ in self.layers]

    # ------------------------------------------------------------------------------------------
    def set_input(self, features: torch.Tensor, coords: torch.Tensor):
        """features (n, <=16) any float dtype, coords (n, 4) int32 [b,z,y,x]; device-to-device copies into the step's buffers."""
        n = features.shape[0]
        assert n <= self.caps[0], f"{n} voxels exceed the capacity {self.caps[0]}"
        self.feats[:n, :features.shape[1]].copy_(features)
        self.coords[:n].copy_(coords)
        self.n0.fill_(n)

    def _forward_backward(self):
        B, caps = self.B, self.caps
        # Rulebooks: the occupancy of every level and level 0's SubM map on this stream (all that conv_input / conv1 wait for);
        # the numbering, the other maps and the input-stationary maps of the strided layers (needed by the BACKWARD pass only)
        # on the rulebook stream, beside the first convolutions.
        main = torch.cuda.current_stream()
        rb = F.rulebook_chain(self.coords, self.n0, B, self.shape, self.convs, self.subm_ksizes, caps=caps, phase=1 | 4)
        counts = [self.n0] + [c[:1] for c in rb["counts"][1:]]
        self.level_counts = rb["counts"]
        self._rb_forked.record(main)
        self.rb_stream.wait_event(self._rb_forked)
        maps, inv = {}, {}
        with torch.cuda.stream(self.rb_stream):
            rb["run"](2)
            self._maps_ready.record(self.rb_stream)
            strided_i = 0
            for l in self.layers:
                if l["key"] in maps:
                    continue
                if l["kind"] == "subm":
                    maps[l["key"]] = rb["nbr_subm"][l["level_in"]]
                else:
                    strided_i += 1
                    maps[l["key"]] = rb["nbr_conv"][strided_i]
                    inv[l["key"]] = F.rulebook_invert(maps[l["key"]], caps[l["level_out"]], caps[l["level_in"]],
                                                      n_out_dev=counts[l["level_out"]], n_in_dev=counts[l["level_in"]])
            self._inv_ready.record(self.rb_stream)
        waited_maps = False
        # ---- forward --------------------------------------------------------------------------------------------------
        x, saved = self.feats, []
        for l in self.layers:
            conv, bn = l["conv"], l["bn"]
            if l["level_out"] > 0 and not waited_maps:       # first layer that needs more than level 0's SubM map
                main.wait_event(self._maps_ready)
                waited_maps = True
            K = maps[l["key"]].shape[0]
            w = conv.weight.detach().view(K, l["c_in"], l["c_out"])
            if l["c_in"] < 16:
                w = torch.nn.functional.pad(w, (0, 0, 0, 16 - l["c_in"]))
            c_in = w.shape[1]
            cnt, cap = counts[l["level_out"]], caps[l["level_out"]]
            y = F.sparse_conv_fwd(x, None, maps[l["key"]], cap, n_out_dev=cnt, weight_packed=F.pack_conv_weights(w.contiguous()),
                                  weight_shape=(K, c_in, l["c_out"]))
            bn.num_batches_tracked.add_(1)
            res = F.bn_train_fwd(y, bn.weight.detach(), bn.bias.detach(), bn.eps, bn.momentum, bn.running_mean,
                                 bn.running_var, relu=True, n_dev=cnt, process_group=self.bn_pg)
            out, stats = res[0], res[1]
            saved.append((x, w, y, out, stats, res[2] if self.bn_pg is not None else None))
            x = out
        last = self.layers[-1]
        shape_out = rb["shapes"][last["level_out"]]
        dense = F.to_dense(x, rb["coords"][last["level_out"]], shape_out, B, n_dev=counts[last["level_out"]], n=caps[last["level_out"]])
        loss, grad_dense = self.loss_fn(dense)
        self.loss.copy_(loss.detach())
        # ---- backward -------------------------------------------------------------------------------------------------
        g = F.from_dense(grad_dense.contiguous(), rb["coords"][last["level_out"]], caps[last["level_out"]],
                         counts[last["level_out"]], out_dtype=torch.bfloat16)
        main.wait_event(self._inv_ready)
        works, start, keep, reduced_upto = [], 0, [], 0
        for bi, (l, (x_in, w, y, out, stats, sums)) in enumerate(zip(reversed(self.layers), reversed(saved))):
            conv, bn = l["conv"], l["bn"]
            K, c_in, c_out = w.shape
            cnt_out, cap_out = counts[l["level_out"]], caps[l["level_out"]]
            gy, _, _ = F.bn_train_bwd(g, out, y, bn.weight.detach(), stats, relu=True, n_dev=cnt_out,
                                      grad_gamma=bn.weight.grad, grad_beta=bn.bias.grad, process_group=self.bn_pg, fwd_sums=sums)
            gw = conv.weight.grad.view(K, l["c_in"], c_out)
            keep.append(gy)                         # read on the other stream: its memory must not be handed out again in this step
            self._gy_ready[bi].record(main)
            self.wg_stream.wait_event(self._gy_ready[bi])
            with torch.cuda.stream(self.wg_stream):
                if l["c_in"] == c_in:
                    F.sparse_conv_wgrad(x_in, gy, maps[l["key"]], cap_out, n_out_dev=cnt_out, out=gw)
                else:
                    gw.copy_(F.sparse_conv_wgrad(x_in, gy, maps[l["key"]], cap_out, n_out_dev=cnt_out)[:, :l["c_in"]])
                self._wg_done[bi].record(self.wg_stream)
            if bi + 1 < len(self.layers):          # conv_input's own input needs no gradient
                centred = l["kind"] == "subm"
                nb = maps[l["key"]] if centred else inv[l["key"]]
                cap_in, cnt_in = caps[l["level_in"]], counts[l["level_in"]]
                if c_out <= 64:
                    wpt = F.pack_conv_weights(w.contiguous(), transpose=True, flip=centred)
                    g = F.sparse_conv_fwd(gy, None, nb, cap_in, n_out_dev=cnt_in, weight_packed=wpt, weight_shape=(K, c_out, c_in))
                else:       # 128 gradient channels (conv_out, K = 3): two 64-channel pseudo-offsets per offset
                    assert c_out == 128 and 2 * K <= 27 and not centred
                    w2 = w.view(K, c_in, 2, 64).permute(0, 2, 1, 3).reshape(2 * K, c_in, 64).contiguous()
                    nb2 = torch.stack([torch.where(nb >= 0, nb * 2, nb), torch.where(nb >= 0, nb * 2 + 1, nb)], 1)
                    g = F.sparse_conv_fwd(gy.view(-1, 64), None, nb2.view(2 * K, -1).contiguous(), cap_in, n_out_dev=cnt_in,
                                          weight_packed=F.pack_conv_weights(w2, transpose=True), weight_shape=(2 * K, 64, c_in))
            if self.pg is not None and bi in self.bucket_after_layer:
                end = self.bucket_after_layer[bi]
                for j in range(reduced_upto, bi + 1):                 # the bucket's weight gradients must be complete
                    torch.cuda.current_stream().wait_event(self._wg_done[j])
                reduced_upto = bi + 1
                works.append(torch.distributed.all_reduce(self.flat_grad[start:end], op=torch.distributed.ReduceOp.AVG,
                                                          group=self.pg, async_op=True))
                start = end
        torch.cuda.current_stream().wait_stream(self.wg_stream)
        for w_ in works:
            w_.wait()
        del keep
        if self.clip is not None:
            norm = torch.linalg.vector_norm(self.flat_grad)
            self.grad_norm.copy_(norm)
            self.flat_grad.mul_(torch.clamp(self.clip / (norm + 1e-6), max=1.0))

    def step(self):
        """One eager step on the current stream (no host synchronisation either)."""
        self._forward_backward()
        self.opt.step()
        return self.loss

    def capture(self, warmup: int = 3):
        """Capture step() into a CUDA graph (after `warmup` eager steps on a side stream); replay() runs it."""
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step()
        return self

    def replay(self):
        self.graph.replay()
        return self.loss
