"""Torch-tensor front end of the C ABI: allocates outputs/workspaces with torch (device memory and
streams are plumbing), then calls libpcdet_b200.so on torch's CURRENT stream.

Every function here requires CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import BF16, EPI_RELU, F32, PACK_FLIP, PACK_TRANSPOSE, WEIGHT_PACKED, check, f32xN, i32x3, lib, ptr

_workspaces = {}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes: int, device, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer per (device, tag).  Buffers with different tags never alias, so a
    rulebook build and the NMS of another stream can be in flight together."""
    key = (torch.device(device).index, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype} (float32 or bfloat16)")


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.PcdbError("pcdet_b200 ops take CUDA tensors only (no CPU fallback)")


def _triple(v):
    if isinstance(v, (int, np.integer)):
        return [int(v)] * 3
    v = [int(x) for x in v]
    assert len(v) == 3
    return v


# ----------------------------------------------------------------------------------------------
# voxelisation + VFE
# ----------------------------------------------------------------------------------------------
def grid_size(voxel_size, point_cloud_range):
    """spconv VoxelGenerator: round((range[3:] - range[:3]) / voxel_size) in float32 -> int64 (x,y,z)."""
    r = np.asarray(point_cloud_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


def voxelize(points: torch.Tensor, frame_offsets: torch.Tensor, batch_size: int, voxel_size, point_cloud_range,
             max_num_points: int, max_voxels: int, overflow_break: bool = True, want_voxels: bool = True,
             want_mean: bool = False, mean_dtype=torch.float32, mean_stride: Optional[int] = None,
             want_point_idx: bool = False, capacity: Optional[int] = None):
    """Batched point->voxel assignment (+ optional fused mean VFE).

    points (N,C) f32 cuda, frames concatenated; frame_offsets (B+1) i32 cuda.
    Returns a dict of capacity-sized tensors plus `voxel_offsets` (B+1) i32 on the device; nothing is
    copied to the host.  Rows beyond voxel_offsets[B] are undefined.
    """
    _require_cuda(points, frame_offsets)
    assert points.dtype == torch.float32 and points.dim() == 2 and points.is_contiguous()
    assert frame_offsets.dtype == torch.int32 and frame_offsets.numel() == batch_size + 1
    n, c = points.shape
    dev = points.device
    grid = grid_size(voxel_size, point_cloud_range)
    cap = min(n, batch_size * max_voxels) if capacity is None else int(capacity)
    cap = max(cap, 1)
    out = {}
    out["voxels"] = torch.empty((cap, max_num_points, c), dtype=torch.float32, device=dev) if want_voxels else None
    out["coordinates"] = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    out["num_points"] = torch.empty((cap,), dtype=torch.int32, device=dev)
    ms = int(mean_stride or c)
    out["mean"] = torch.empty((cap, ms), dtype=mean_dtype, device=dev) if want_mean else None
    out["point_idx"] = torch.empty((cap, max_num_points), dtype=torch.int32, device=dev) if want_point_idx else None
    out["voxel_offsets"] = torch.empty((batch_size + 1,), dtype=torch.int32, device=dev)
    L = lib()
    nbytes = L.pcdb_voxelize_workspace_bytes(n, batch_size, max_num_points, max_voxels)
    ws = workspace(nbytes, dev, "voxelize")
    check(L.pcdb_voxelize(
        ptr(points), n, c, ptr(frame_offsets), batch_size, f32xN(np.asarray(voxel_size, np.float32)),
        f32xN(np.asarray(point_cloud_range, np.float32)), i32x3(grid), max_num_points, max_voxels,
        int(overflow_break), ptr(out["voxels"]), ptr(out["coordinates"]), ptr(out["num_points"]),
        ptr(out["mean"]), BF16 if mean_dtype == torch.bfloat16 else F32, ms, ptr(out["point_idx"]),
        ptr(out["voxel_offsets"]), ptr(ws), ws.numel(), _stream()), "pcdb_voxelize")
    return out


def vfe_mean(voxels: torch.Tensor, num_points: torch.Tensor, out_dtype=torch.float32,
             out_stride: Optional[int] = None) -> torch.Tensor:
    """MeanVoxelFeatureExtractor.forward (pcdet/models/vfe/vfe_utils.py:26-34)."""
    _require_cuda(voxels, num_points)
    voxels = voxels.contiguous().float()
    num_points = num_points.contiguous().int()
    v, p, c = voxels.shape
    ms = int(out_stride or c)
    out = torch.empty((v, ms), dtype=out_dtype, device=voxels.device)
    check(lib().pcdb_vfe_mean(ptr(voxels), ptr(num_points), v, p, c, ptr(out),
                              BF16 if out_dtype == torch.bfloat16 else F32, ms, _stream()), "pcdb_vfe_mean")
    return out


def pillar_vfe(voxels: torch.Tensor, num_points: torch.Tensor, coords: torch.Tensor, weight: torch.Tensor,
               scale: Optional[torch.Tensor], shift: Optional[torch.Tensor], voxel_size, center_offset,
               with_distance: bool = False, want_features: bool = True, canvas_shape=None, batch_size: int = 1,
               n_dev: Optional[torch.Tensor] = None):
    """PillarFeatureNetOld2 (one PFN layer, eval-mode BN as scale/shift) [+ PointPillarsScatter].

    voxels (N,P,C) f32, num_points (N) i32, coords (N,4) i32 [b,z,y,x], weight (F, C+6[+1]) f32.
    center_offset = voxel_size/2 + range_min per axis (vfe_utils.py:162-164).
    Returns (features (N,F) or None, canvas (B, F*nz, ny, nx) or None)."""
    _require_cuda(voxels, num_points, coords, weight)
    assert voxels.dtype == torch.float32 and voxels.dim() == 3 and voxels.is_contiguous()
    assert coords.dtype == torch.int32 and coords.shape[1] == 4 and coords.is_contiguous()
    num_points = num_points.to(torch.int32).contiguous()
    weight = weight.contiguous().float()
    n, p, c = voxels.shape
    f = weight.shape[0]
    dev = voxels.device
    feats = torch.empty((n, f), dtype=torch.float32, device=dev) if want_features else None
    canvas = None
    if canvas_shape is not None:
        nz, ny, nx = (int(v) for v in canvas_shape)
        canvas = torch.empty((batch_size, f * nz, ny, nx), dtype=torch.float32, device=dev)
    check(lib().pcdb_pillar_vfe(ptr(voxels), ptr(num_points), ptr(coords), n, ptr(n_dev), p, c, f32xN(voxel_size),
                                f32xN(center_offset), int(with_distance), ptr(weight), f,
                                ptr(scale.contiguous().float()) if scale is not None else None,
                                ptr(shift.contiguous().float()) if shift is not None else None, ptr(feats), ptr(canvas),
                                batch_size, i32x3(canvas_shape) if canvas_shape is not None else None, _stream()),
          "pcdb_pillar_vfe")
    return feats, canvas


# ----------------------------------------------------------------------------------------------
# rulebook
# ----------------------------------------------------------------------------------------------
def conv_output_size(in_shape, ksize, stride, padding, dilation):
    """spconv.ops.get_conv_output_size."""
    return [int((i + 2 * p - d * (k - 1) - 1) // s + 1)
            for i, k, s, p, d in zip(in_shape, ksize, stride, padding, dilation)]


def rulebook_subm(indices: torch.Tensor, batch_size: int, spatial_shape: Sequence[int], ksize=3, dilation=1,
                  n_dev: Optional[torch.Tensor] = None, site_table=None) -> torch.Tensor:
    """Neighbour map (K, N) i32 of a submanifold convolution (stride 1, padding k/2).

    site_table: the `site_table` entry of the rulebook_conv(..., keep_table=True) result whose out_indices are
    `indices` -- its hash table is looked up instead of building a new one (pcdb_rulebook_subm_reuse)."""
    _require_cuda(indices)
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == 4 and indices.is_contiguous()
    n = indices.shape[0]
    ks, dl = _triple(ksize), _triple(dilation)
    K = ks[0] * ks[1] * ks[2]
    nbr = torch.empty((K, max(n, 1)), dtype=torch.int32, device=indices.device)
    L = lib()
    if site_table is not None:
        tws, t_in_cap, t_k, t_out_cap = site_table
        check(L.pcdb_rulebook_subm_reuse(ptr(indices), n, ptr(n_dev), batch_size, i32x3(spatial_shape), i32x3(ks),
                                         i32x3(dl), ptr(nbr), nbr.shape[1], ptr(tws), t_in_cap, t_k, t_out_cap, 0,
                                         _stream()), "pcdb_rulebook_subm_reuse")
        return nbr
    ws = workspace(L.pcdb_rulebook_workspace_bytes(n, K, n), indices.device, "rulebook")
    check(L.pcdb_rulebook_subm(ptr(indices), n, ptr(n_dev), batch_size, i32x3(spatial_shape), i32x3(ks), i32x3(dl),
                               ptr(nbr), nbr.shape[1], ptr(ws), ws.numel(), _stream()), "pcdb_rulebook_subm")
    return nbr


def max_outputs_per_input(ksize, stride, dilation):
    """Upper bound on distinct output sites one input site can reach."""
    m = 1
    for k, s, d in zip(ksize, stride, dilation):
        m *= min(k, -(-((k - 1) * d + 1) // s))
    return m


def rulebook_conv(indices: torch.Tensor, batch_size: int, spatial_shape: Sequence[int], ksize, stride, padding,
                  dilation=1, n_dev: Optional[torch.Tensor] = None, out_capacity: Optional[int] = None,
                  want_inverse: bool = True, keep_table: bool = False):
    """Regular sparse convolution rulebook.

    Returns dict(out_indices (cap,4) i32, n_out (2,) i32 device [count, overflow flag], nbr (K,cap),
    nbr_inv (K,N) or None, out_shape [z,y,x]).  Nothing is copied to the host."""
    _require_cuda(indices)
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == 4 and indices.is_contiguous()
    n = indices.shape[0]
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    K = ks[0] * ks[1] * ks[2]
    out_shape = conv_output_size(spatial_shape, ks, st, pd, dl)
    if out_capacity is None:
        vol = batch_size * out_shape[0] * out_shape[1] * out_shape[2]
        out_capacity = min(n * max_outputs_per_input(ks, st, dl), vol)
    cap = max(int(out_capacity), 1)
    dev = indices.device
    out_indices = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    n_out = torch.empty((2,), dtype=torch.int32, device=dev)
    nbr = torch.empty((K, cap), dtype=torch.int32, device=dev)
    nbr_inv = torch.empty((K, max(n, 1)), dtype=torch.int32, device=dev) if want_inverse else None
    L = lib()
    nbytes = L.pcdb_rulebook_workspace_bytes(n, K, cap)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev) if keep_table else workspace(nbytes, dev, "rulebook")
    check(L.pcdb_rulebook_conv(ptr(indices), n, ptr(n_dev), batch_size, i32x3(spatial_shape), i32x3(out_shape),
                               i32x3(ks), i32x3(st), i32x3(pd), i32x3(dl), ptr(out_indices), cap, ptr(n_out),
                               ptr(nbr), cap, ptr(nbr_inv), nbr_inv.shape[1] if want_inverse else 0, ptr(ws),
                               ws.numel(), _stream()), "pcdb_rulebook_conv")
    return dict(out_indices=out_indices, n_out=n_out, nbr=nbr, nbr_inv=nbr_inv, out_shape=out_shape,
                site_table=(ws, n, K, cap) if keep_table else None)


def rulebook_chain(indices: torch.Tensor, n_dev: Optional[torch.Tensor], batch_size: int, spatial_shape: Sequence[int],
                   convs: Sequence[dict], subm_ksizes: Sequence, caps: Optional[Sequence[int]] = None, phase: int = 7):
    """Every rulebook of a strided backbone in four launches (pcdb_rulebook_chain, csrc/rulebook_chain.cu).

    indices (n0, 4) int32 [b,z,y,x] = level 0; convs = [{"ksize":, "stride":, "padding":}, ...] applied one after the
    other (level l = conv l of level l-1); subm_ksizes[l] = kernel size of the SubM map wanted at level l, or None.
    Returns a dict with per-level lists: shapes, caps, coords (level 0 = indices), counts ([count, overflow] device
    tensors; level 0: n_dev), nbr_conv (None at level 0) and nbr_subm.  Rows of levels >= 1 are in ascending (b,z,y,x) order.
    phase (a mask, see include/pcdet_b200.h): 7 = everything; 5 = the occupancy of every level and level 0's SubM map only --
    the dict's "run" entry then issues further phases (e.g. run(2) on another stream, beside the first convolutions)."""
    _require_cuda(indices)
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == 4 and indices.is_contiguous()
    dev = indices.device
    n_levels = len(convs) + 1
    assert len(subm_ksizes) == n_levels
    shapes = [[int(v) for v in spatial_shape]]
    for c in convs:
        shapes.append(conv_output_size(shapes[-1], _triple(c["ksize"]), _triple(c["stride"]), _triple(c["padding"]), [1, 1, 1]))
    n0 = indices.shape[0]
    if caps is None:
        caps = [n0]
        for c in convs:
            caps.append(min(caps[-1] * max_outputs_per_input(_triple(c["ksize"]), _triple(c["stride"]), [1, 1, 1]),
                            batch_size * int(np.prod(shapes[len(caps)]))))
    caps = [max(int(c), 1) for c in caps]
    if n_dev is None:
        n_dev = torch.tensor([n0], dtype=torch.int32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    coords = [indices] + [torch.empty((caps[l], 4), **i32) for l in range(1, n_levels)]
    counts = [n_dev] + [torch.zeros((2,), **i32) for _ in range(1, n_levels)]
    nbr_conv = [None] + [torch.empty((int(np.prod(_triple(c["ksize"]))), caps[l + 1]), **i32) for l, c in enumerate(convs)]
    nbr_subm = [None if k is None else torch.empty((int(np.prod(_triple(k))), caps[l]), **i32) for l, k in enumerate(subm_ksizes)]
    L = lib()
    caps_a = (C.c_int32 * n_levels)(*caps)
    flat = lambda rows: (C.c_int32 * (3 * len(rows)))(*[int(v) for r in rows for v in r])
    nbytes = L.pcdb_rulebook_chain_workspace_bytes(batch_size, n_levels, flat(shapes), caps_a)
    if nbytes == 0:
        raise _lib.PcdbError("pcdb_rulebook_chain: a level exceeds the cell-index limits (2^32 cells at level 0, 2^31 above)")
    ws = workspace(nbytes, dev, "rulebook_chain")
    ptrs = lambda ts: (C.c_void_p * n_levels)(*[None if t is None else t.data_ptr() for t in ts])
    a_shapes, a_coords, a_counts, a_conv, a_subm = flat(shapes), ptrs(coords), ptrs(counts), ptrs(nbr_conv), ptrs(nbr_subm)
    a_ks = flat([_triple(c["ksize"]) for c in convs]) if convs else None
    a_st = flat([_triple(c["stride"]) for c in convs]) if convs else None
    a_pd = flat([_triple(c["padding"]) for c in convs]) if convs else None
    a_sk = flat([[0, 0, 0] if k is None else _triple(k) for k in subm_ksizes])

    def run(ph: int):
        check(L.pcdb_rulebook_chain(ptr(indices), ptr(n_dev), batch_size, n_levels, a_shapes, a_ks, a_st, a_pd, caps_a, a_coords,
                                    a_counts, a_conv, a_sk, a_subm, None, ptr(ws), ws.numel(), 0, ph, _stream()),
              "pcdb_rulebook_chain")

    run(phase)
    return dict(shapes=shapes, caps=caps, coords=coords, counts=counts, nbr_conv=nbr_conv, nbr_subm=nbr_subm, run=run)


# ----------------------------------------------------------------------------------------------
# sparse convolution
# ----------------------------------------------------------------------------------------------
def tc_eligible(dtype, c_in: int, c_out: int, kernel_volume: int) -> bool:
    """Shapes the tcgen05 kernel takes (csrc/sparse_conv_tc.cu)."""
    return dtype == torch.bfloat16 and c_in in (16, 32, 64) and c_out in (16, 32, 64, 128) and kernel_volume <= 27


def pack_conv_weights(weight: torch.Tensor, transpose: bool = False, flip: bool = False) -> torch.Tensor:
    """Tensor-core operand image (uint8 buffer) of a bf16 or fp32 (K, Cin, Cout) weight; cache it per layer.
    transpose / flip: the image of the layer's INPUT-GRADIENT convolution instead (Cout -> Cin channels, W[k]^T, with
    `flip` the offsets reversed: the rulebook of a centred submanifold convolution read the other way round)."""
    _require_cuda(weight)
    assert weight.dtype in (torch.bfloat16, torch.float32) and weight.is_contiguous() and weight.dim() == 3
    K, c_in, c_out = weight.shape
    if transpose:
        c_in, c_out = c_out, c_in
    L = lib()
    nbytes = L.pcdb_conv_packed_weight_bytes(K, c_in, c_out)
    assert nbytes > 0, f"no tensor-core kernel for K={K} c_in={c_in} c_out={c_out}"
    packed = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device)
    flags = (PACK_TRANSPOSE if transpose else 0) | (PACK_FLIP if flip else 0)
    check(L.pcdb_pack_conv_weights_ex(ptr(weight), _dt(weight), K, c_in, c_out, flags, ptr(packed), _stream()),
          "pcdb_pack_conv_weights_ex")
    return packed


def sparse_conv_fwd(features: torch.Tensor, weight: Optional[torch.Tensor], nbr: torch.Tensor, n_out: int,
                    n_out_dev: Optional[torch.Tensor] = None, scale=None, shift=None, bias=None, relu: bool = False,
                    algo: int = 0, out: Optional[torch.Tensor] = None,
                    weight_packed: Optional[torch.Tensor] = None, weight_shape=None,
                    residual: Optional[torch.Tensor] = None, residual_post: bool = False) -> torch.Tensor:
    """out[o] = epilogue(sum_k features[nbr[k,o]] @ weight[k]).  weight (K, Cin, Cout), same dtype as features.
    weight_packed: optional cached pack_conv_weights(weight) for the tensor-core kernels (made on the fly otherwise);
    with it `weight` may be None and weight_shape = (K, Cin, Cout) names the layer.
    residual (n_out, Cout) bf16, tensor-core path only: added to the sum before the epilogue (a partial result over other
    input channels), or behind scale / shift and before the ReLU with residual_post (a shortcut connection).
    algo: 0 auto, 1 FMA-pipe kernel, 2 tcgen05 + TMA gather, 3 tcgen05 + cp.async gather."""
    _require_cuda(features, weight, nbr)
    assert features.is_contiguous() and nbr.is_contiguous() and nbr.dtype == torch.int32
    if weight is None:
        assert weight_packed is not None and weight_shape is not None
        K, c_in, c_out = (int(v) for v in weight_shape)
    else:
        assert weight.is_contiguous() and weight.dtype == features.dtype
        K, c_in, c_out = weight.shape
    assert features.shape[1] == c_in and nbr.shape[0] == K and nbr.shape[1] >= n_out
    if out is None:
        out = torch.empty((n_out, c_out), dtype=features.dtype, device=features.device)
    for v in (scale, shift, bias):
        assert v is None or (v.dtype == torch.float32 and v.is_cuda and v.numel() == c_out)
    flags = EPI_RELU if relu else 0
    w = weight
    if algo != 1 and tc_eligible(features.dtype, c_in, c_out, K):
        w = weight_packed if weight_packed is not None else pack_conv_weights(weight)
        flags |= WEIGHT_PACKED
    assert w is not None, "packed weights were given for a shape the tensor-core kernels do not take"
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.is_contiguous() and residual.shape[0] >= n_out and residual.shape[1] == c_out
        flags |= _lib.EPI_RESIDUAL_POST if residual_post else 0
    check(lib().pcdb_sparse_conv_fwd_ex(ptr(features), features.shape[0], ptr(w), ptr(nbr), nbr.shape[1], K, n_out,
                                        ptr(n_out_dev), c_in, c_out, _dt(features), ptr(scale), ptr(shift), ptr(bias),
                                        ptr(residual), flags, ptr(out), algo, _stream()), "pcdb_sparse_conv_fwd")
    return out


def sparse_conv_bwd(features: torch.Tensor, weight: torch.Tensor, grad_out: torch.Tensor, nbr: torch.Tensor,
                    n_out: int, need_input_grad: bool = True, need_weight_grad: bool = True):
    """fp32 backward of sparse_conv_fwd without epilogue: (grad_features, grad_weight)."""
    _require_cuda(features, weight, grad_out, nbr)
    features, weight, grad_out = features.contiguous().float(), weight.contiguous().float(), grad_out.contiguous().float()
    K, c_in, c_out = weight.shape
    gf = torch.zeros_like(features) if need_input_grad else None
    gw = torch.zeros_like(weight) if need_weight_grad else None
    check(lib().pcdb_sparse_conv_bwd(ptr(features), ptr(weight), ptr(grad_out), ptr(nbr), nbr.shape[1], K,
                                     features.shape[0], n_out, c_in, c_out, ptr(gf), ptr(gw), _stream()),
          "pcdb_sparse_conv_bwd")
    return gf, gw


def sparse_conv_wgrad(features: torch.Tensor, grad_out: torch.Tensor, nbr: torch.Tensor, n_out: int,
                      n_out_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                      accumulate: bool = False) -> torch.Tensor:
    """Weight gradient on the tensor cores: out[k] (+)= features[nbr[k]]^T @ grad_out, fp32 (K, Cin, Cout).
    features (n_in, Cin), grad_out (>= n_out, Cout) bf16 (csrc/sparse_conv_wgrad_tc.cu)."""
    _require_cuda(features, grad_out, nbr)
    assert features.dtype == torch.bfloat16 and grad_out.dtype == torch.bfloat16 and nbr.dtype == torch.int32
    assert features.is_contiguous() and grad_out.is_contiguous() and nbr.is_contiguous()
    K, c_in, c_out = nbr.shape[0], features.shape[1], grad_out.shape[1]
    assert nbr.shape[1] >= n_out and grad_out.shape[0] >= n_out
    if out is None:
        assert not accumulate
        out = torch.empty((K, c_in, c_out), dtype=torch.float32, device=features.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (K, c_in, c_out)
    L = lib()
    nbytes = L.pcdb_sparse_conv_wgrad_workspace_bytes(K, n_out, c_in, c_out)
    assert nbytes > 0, f"no tensor-core weight-gradient kernel for K={K} c_in={c_in} c_out={c_out}"
    ws = workspace(nbytes, features.device, "wgrad")
    check(L.pcdb_sparse_conv_wgrad(ptr(features), features.shape[0], ptr(grad_out), ptr(nbr), nbr.shape[1], K, n_out,
                                   ptr(n_out_dev), c_in, c_out, ptr(out), int(accumulate), ptr(ws), ws.numel(), _stream()),
          "pcdb_sparse_conv_wgrad")
    return out


def bn_train_fwd(y: torch.Tensor, gamma, beta, eps: float, momentum: float, running_mean=None, running_var=None,
                 relu: bool = True, conv_partials: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None,
                 process_group=None):
    """Train-mode BatchNorm1d (+ ReLU) over the rows of y (n, C): returns (out, stats (4, C) = mean, 1/std, scale, shift);
    running statistics are updated in place (csrc/bn_train.cu).  process_group: SyncBatchNorm -- the statistics are those of
    every rank's rows (one all-reduce of 2C + 1 doubles); the third return value then holds the global sums for bn_train_bwd."""
    _require_cuda(y)
    assert y.is_contiguous() and y.dim() == 2
    n, c = y.shape
    out = torch.empty_like(y)
    stats = torch.empty((4, c), dtype=torch.float32, device=y.device)
    L = lib()
    ws = workspace(L.pcdb_bn_train_workspace_bytes(), y.device, "bn")
    flags = EPI_RELU if relu else 0
    if process_group is not None:
        sums = torch.empty((2 * c + 1,), dtype=torch.float64, device=y.device)
        check(L.pcdb_bn_train_sums(ptr(y), n, ptr(n_dev), c, _dt(y), ptr(sums), ptr(ws), ws.numel(), _stream()), "pcdb_bn_train_sums")
        torch.distributed.all_reduce(sums, group=process_group)
        check(L.pcdb_bn_train_fwd_from_sums(ptr(y), n, ptr(n_dev), c, _dt(y), ptr(sums), ptr(gamma), ptr(beta), float(eps),
                                            float(momentum), ptr(running_mean), ptr(running_var), flags, ptr(out), ptr(stats),
                                            _stream()), "pcdb_bn_train_fwd_from_sums")
        return out, stats, sums
    n_part = 0 if conv_partials is None else conv_partials.shape[0]
    check(L.pcdb_bn_train_fwd(ptr(y), n, ptr(n_dev), c, _dt(y), ptr(gamma), ptr(beta), float(eps), float(momentum),
                              ptr(running_mean), ptr(running_var), flags, ptr(out), ptr(stats),
                              ptr(conv_partials), n_part, ptr(ws), ws.numel(), _stream()), "pcdb_bn_train_fwd")
    return out, stats


def bn_train_bwd(grad_out: torch.Tensor, out: torch.Tensor, y: torch.Tensor, gamma, stats: torch.Tensor, relu: bool = True,
                 n_dev: Optional[torch.Tensor] = None, grad_gamma: Optional[torch.Tensor] = None,
                 grad_beta: Optional[torch.Tensor] = None, grad_y: Optional[torch.Tensor] = None,
                 process_group=None, fwd_sums: Optional[torch.Tensor] = None):
    """Backward of bn_train_fwd: (grad_y like y, grad_gamma (C) fp32, grad_beta (C) fp32); the optional output tensors are
    overwritten (e.g. views of a flat gradient buffer).  process_group + fwd_sums (third result of the synchronised
    forward): SyncBatchNorm backward -- grad_gamma / grad_beta from this rank's rows, grad_y from every rank's sums."""
    _require_cuda(grad_out, y, stats)
    assert grad_out.is_contiguous() and y.is_contiguous() and grad_out.dtype == y.dtype and grad_out.shape == y.shape
    n, c = y.shape
    grad_y = torch.empty_like(y) if grad_y is None else grad_y
    gg = torch.empty((c,), dtype=torch.float32, device=y.device) if grad_gamma is None else grad_gamma
    gb = torch.empty((c,), dtype=torch.float32, device=y.device) if grad_beta is None else grad_beta
    assert gg.dtype == torch.float32 and gb.dtype == torch.float32 and gg.is_contiguous() and gb.is_contiguous()
    L = lib()
    ws = workspace(L.pcdb_bn_train_workspace_bytes(), y.device, "bn")
    flags = EPI_RELU if relu else 0
    if process_group is not None:
        assert fwd_sums is not None and fwd_sums.dtype == torch.float64
        local = torch.empty((2 * c,), dtype=torch.float64, device=y.device)
        check(L.pcdb_bn_train_bwd_sums(ptr(grad_out), ptr(out), ptr(y), n, ptr(n_dev), c, _dt(y), ptr(stats), flags, ptr(local),
                                       ptr(ws), ws.numel(), _stream()), "pcdb_bn_train_bwd_sums")
        glob = local.clone()
        torch.distributed.all_reduce(glob, group=process_group)
        check(L.pcdb_bn_train_bwd_from_sums(ptr(grad_out), ptr(out), ptr(y), n, ptr(n_dev), c, _dt(y), ptr(gamma), ptr(stats),
                                            ptr(local), ptr(glob), ptr(fwd_sums), flags, ptr(grad_y), ptr(gg), ptr(gb), 0, ptr(ws),
                                            ws.numel(), _stream()), "pcdb_bn_train_bwd_from_sums")
        return grad_y, gg, gb
    check(L.pcdb_bn_train_bwd(ptr(grad_out), ptr(out), ptr(y), n, ptr(n_dev), c, _dt(y), ptr(gamma), ptr(stats),
                              flags, ptr(grad_y), ptr(gg), ptr(gb), 0, ptr(ws), ws.numel(), _stream()),
          "pcdb_bn_train_bwd")
    return grad_y, gg, gb


def sparse_maxpool_fwd(features: torch.Tensor, nbr: torch.Tensor, n_out: int,
                       n_out_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """spconv indice_maxpool forward: out[o] = max(0, max_k features[nbr[k, o]])."""
    _require_cuda(features, nbr)
    assert features.is_contiguous() and nbr.is_contiguous() and nbr.dtype == torch.int32
    out = torch.empty((n_out, features.shape[1]), dtype=features.dtype, device=features.device)
    check(lib().pcdb_sparse_maxpool_fwd(ptr(features), ptr(nbr), nbr.shape[1], nbr.shape[0], n_out, ptr(n_out_dev),
                                        features.shape[1], _dt(features), ptr(out), _stream()), "pcdb_sparse_maxpool_fwd")
    return out


class _SparseMaxPool(torch.autograd.Function):
    """indice_maxpool with spconv's backward (the gradient goes to every input equal to the pooled maximum)."""

    @staticmethod
    def forward(ctx, features, nbr, n_out):
        out = sparse_maxpool_fwd(features, nbr, n_out)
        ctx.save_for_backward(features, out, nbr)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, out, nbr = ctx.saved_tensors
        assert features.dtype == torch.float32, "training runs in fp32"
        grad_in = torch.zeros_like(features)
        check(lib().pcdb_sparse_maxpool_bwd(ptr(features), ptr(out), ptr(grad_out.contiguous()), ptr(nbr), nbr.shape[1], nbr.shape[0],
                                            out.shape[0], features.shape[1], ptr(grad_in), _stream()), "pcdb_sparse_maxpool_bwd")
        return grad_in, None, None


def sparse_maxpool(features: torch.Tensor, nbr: torch.Tensor, n_out: int) -> torch.Tensor:
    """sparse_maxpool_fwd that takes part in autograd."""
    if torch.is_grad_enabled() and features.requires_grad:
        return _SparseMaxPool.apply(features, nbr, n_out)
    return sparse_maxpool_fwd(features, nbr, n_out)


class _ToDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, indices, spatial_shape, batch_size):
        ctx.save_for_backward(indices)
        return to_dense(features.contiguous(), indices, spatial_shape, batch_size)

    @staticmethod
    def backward(ctx, grad):
        (indices,) = ctx.saved_tensors
        i = indices.long()
        return grad[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]].contiguous(), None, None, None


def to_dense_autograd(features, indices, spatial_shape, batch_size):
    """to_dense with a backward (gather of the dense gradient at the active sites) for training."""
    return _ToDense.apply(features, indices, spatial_shape, batch_size)


def to_dense(features: torch.Tensor, indices: torch.Tensor, spatial_shape, batch_size: int,
             n_dev: Optional[torch.Tensor] = None, out_dtype=None, n: Optional[int] = None) -> torch.Tensor:
    """SparseConvTensor.dense(): (B, C, D, H, W)."""
    _require_cuda(features, indices)
    assert features.is_contiguous() and indices.is_contiguous() and indices.dtype == torch.int32
    n = features.shape[0] if n is None else n
    c = features.shape[1]
    out_dtype = out_dtype or features.dtype
    shape = [int(s) for s in spatial_shape]
    dense = torch.empty((batch_size, c, *shape), dtype=out_dtype, device=features.device)
    check(lib().pcdb_to_dense(ptr(features), ptr(indices), n, ptr(n_dev), c, _dt(features), batch_size,
                              i32x3(shape), ptr(dense), BF16 if out_dtype == torch.bfloat16 else F32, _stream()),
          "pcdb_to_dense")
    return dense


def from_dense(grad_dense: torch.Tensor, indices: torch.Tensor, n: Optional[int] = None, n_dev: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, out_dtype=None) -> torch.Tensor:
    """Backward of to_dense: rows (n, C) gathered from a (B, C, D, H, W) tensor at the active sites."""
    _require_cuda(grad_dense, indices)
    assert grad_dense.is_contiguous() and grad_dense.dim() == 5 and indices.is_contiguous() and indices.dtype == torch.int32
    B, c, D, H, W = grad_dense.shape
    n = indices.shape[0] if n is None else n
    if out is None:
        out = torch.zeros((n, c), dtype=out_dtype or grad_dense.dtype, device=grad_dense.device)
    check(lib().pcdb_from_dense(ptr(grad_dense), _dt(grad_dense), ptr(indices), n, ptr(n_dev), c, B, i32x3([D, H, W]), ptr(out),
                                _dt(out), _stream()), "pcdb_from_dense")
    return out


def rulebook_invert(nbr: torch.Tensor, n_out: int, n_in: int, n_out_dev: Optional[torch.Tensor] = None,
                    n_in_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nbr_inv (K, n_in): the input-stationary reading of the output-stationary map nbr (K, >= n_out)."""
    _require_cuda(nbr)
    assert nbr.dtype == torch.int32 and nbr.is_contiguous() and nbr.shape[1] >= n_out
    K = nbr.shape[0]
    if out is None:
        out = torch.empty((K, n_in), dtype=torch.int32, device=nbr.device)
    check(lib().pcdb_rulebook_invert(ptr(nbr), nbr.shape[1], K, n_out, ptr(n_out_dev), ptr(out), out.shape[1], n_in, ptr(n_in_dev),
                                     _stream()), "pcdb_rulebook_invert")
    return out


# ----------------------------------------------------------------------------------------------
# rotated IoU / NMS
# ----------------------------------------------------------------------------------------------
def boxes_overlap_bev(boxes_a: torch.Tensor, boxes_b: torch.Tensor, out: Optional[torch.Tensor] = None):
    _require_cuda(boxes_a, boxes_b)
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    if out is None:
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    check(lib().pcdb_boxes_overlap_bev(ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out), _stream()),
          "pcdb_boxes_overlap_bev")
    return out


def boxes_iou_bev(boxes_a: torch.Tensor, boxes_b: torch.Tensor, out: Optional[torch.Tensor] = None):
    _require_cuda(boxes_a, boxes_b)
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    if out is None:
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    check(lib().pcdb_boxes_iou_bev(ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out), _stream()),
          "pcdb_boxes_iou_bev")
    return out


def boxes_iou3d(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    """(N,7),(M,7) LiDAR boxes [x,y,z,w,l,h,ry] -> (N,M) 3-D IoU in one launch (pcdb_boxes_iou3d)."""
    _require_cuda(boxes_a, boxes_b)
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    assert a.shape[1] == 7 and b.shape[1] == 7
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    check(lib().pcdb_boxes_iou3d(ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out), _stream()), "pcdb_boxes_iou3d")
    return out


def nms_sorted_batched(boxes: torch.Tensor, set_offsets: Sequence[int], thresh: float, normal: bool = False,
                       keep_stride: Optional[int] = None, set_counts: Optional[torch.Tensor] = None):
    """Greedy NMS of several score-sorted box sets in one go, entirely on the device.

    boxes (sum n_s, 5) f32 cuda; set_offsets host ints (S+1); set_counts optional (S,) int32 cuda: boxes actually
    present in every set (the offsets then give the capacities).  Returns (keep (S, keep_stride) int64 with
    -1 padding, num_keep (S,) int32), both on the device."""
    _require_cuda(boxes)
    boxes = boxes.contiguous().float()
    offs = np.ascontiguousarray(np.asarray(set_offsets, dtype=np.int32))
    s = offs.shape[0] - 1
    max_n = int(np.max(np.diff(offs))) if s > 0 else 0
    stride = int(keep_stride or max(max_n, 1))
    keep = torch.empty((s, stride), dtype=torch.int64, device=boxes.device)
    num = torch.empty((s,), dtype=torch.int32, device=boxes.device)
    L = lib()
    ws = workspace(L.pcdb_nms_workspace_bytes(s, max_n), boxes.device, "nms")
    if set_counts is not None:
        assert set_counts.dtype == torch.int32 and set_counts.is_cuda and set_counts.numel() == s
        check(L.pcdb_nms_counts(ptr(boxes), ptr(offs), ptr(set_counts), s, float(thresh), int(normal), ptr(keep), stride,
                                ptr(num), ptr(ws), ws.numel(), _stream()), "pcdb_nms_counts")
    else:
        check(L.pcdb_nms(ptr(boxes), ptr(offs), s, float(thresh), int(normal), ptr(keep), stride, ptr(num), ptr(ws),
                         ws.numel(), _stream()), "pcdb_nms")
    return keep, num


def boxes3d_to_bev(boxes3d: torch.Tensor) -> torch.Tensor:
    """boxes3d_to_bevboxes_lidar_torch (pcdet/utils/box_utils.py:237-250)."""
    _require_cuda(boxes3d)
    b = boxes3d.contiguous().float()
    assert b.shape[1] == 7
    out = torch.empty((b.shape[0], 5), dtype=torch.float32, device=b.device)
    check(lib().pcdb_boxes3d_to_bev(ptr(b), b.shape[0], ptr(out), _stream()), "pcdb_boxes3d_to_bev")
    return out


def decode_select(cls_preds: torch.Tensor, box_preds: torch.Tensor, anchors: torch.Tensor,
                  dir_cls_preds: Optional[torch.Tensor] = None, score_thresh: float = 0.1, pre_max: int = 4096,
                  num_dir_bins: int = 2, dir_offset: float = 0.0, dir_limit_offset: float = 0.0,
                  use_binary_dir_classifier: bool = False, out=None):
    """Front of Detector3D.post_processing for the class-agnostic path (detector3d.py:112-128, 166-215, 278-290;
    box_coder_utils.py:89-144): class max, sigmoid threshold, top-`pre_max`, decode, BEV boxes -- whole batch, no sync.

    cls_preds (B, A, C), box_preds (B, A, 7), anchors (A, 7), dir_cls_preds (B, A, bins) or None, all f32 cuda.
    Returns dict(boxes3d (B,K,7), boxes_bev (B,K,5), scores (B,K), labels (B,K) i32, anchor_index (B,K) i32,
    count (B,) i32) with K = pre_max; rows >= count are padding (see include/pcdet_b200.h)."""
    _require_cuda(cls_preds, box_preds, anchors)
    cls_preds = cls_preds.contiguous().float()
    box_preds = box_preds.contiguous().float()
    anchors = anchors.contiguous().float()
    bsz, a, c = cls_preds.shape
    assert box_preds.shape == (bsz, a, 7) and anchors.shape == (a, 7)
    bins = 0
    if dir_cls_preds is not None:
        dir_cls_preds = dir_cls_preds.contiguous().float().view(bsz, a, -1)
        bins = dir_cls_preds.shape[2]
        if not use_binary_dir_classifier:
            assert bins == num_dir_bins
    dev = cls_preds.device
    if out is None:
        out = dict(boxes3d=torch.empty((bsz, pre_max, 7), dtype=torch.float32, device=dev),
                   boxes_bev=torch.empty((bsz, pre_max, 5), dtype=torch.float32, device=dev),
                   scores=torch.empty((bsz, pre_max), dtype=torch.float32, device=dev),
                   labels=torch.empty((bsz, pre_max), dtype=torch.int32, device=dev),
                   anchor_index=torch.empty((bsz, pre_max), dtype=torch.int32, device=dev),
                   count=torch.empty((bsz,), dtype=torch.int32, device=dev))
    L = lib()
    ws = workspace(L.pcdb_decode_select_workspace_bytes(bsz, a, pre_max), dev, "decode_select")
    check(L.pcdb_decode_select(ptr(cls_preds), c, ptr(box_preds), ptr(dir_cls_preds) if dir_cls_preds is not None else None,
                               ptr(anchors), bsz, a, c, bins, float(dir_offset), float(dir_limit_offset), float(score_thresh),
                               int(pre_max), 1 if use_binary_dir_classifier else 0,
                               ptr(out["boxes3d"]), ptr(out["boxes_bev"]), ptr(out["scores"]), ptr(out["labels"]),
                               ptr(out["anchor_index"]), ptr(out["count"]), ptr(ws), ws.numel(), _stream()),
          "pcdb_decode_select")
    return out


def gather_kept(keep: torch.Tensor, front: dict, post_max: int, sigmoid_scores: bool = False, pad_score: float = 0.0,
                pad_label: int = 0):
    """Kept positions of the NMS (B, stride) + the output of decode_select -> dict(boxes (B,P,7), scores (B,P),
    labels (B,P) i64, selected (B,P) i64, num (B,) i32), P = post_max (detector3d.py:290-299, 211-219)."""
    bsz, pre_max = front["scores"].shape
    dev = keep.device
    out = dict(boxes=torch.empty((bsz, post_max, 7), dtype=torch.float32, device=dev),
               scores=torch.empty((bsz, post_max), dtype=torch.float32, device=dev),
               labels=torch.empty((bsz, post_max), dtype=torch.int64, device=dev),
               selected=torch.empty((bsz, post_max), dtype=torch.int64, device=dev),
               num=torch.empty((bsz,), dtype=torch.int32, device=dev))
    check(lib().pcdb_gather_kept(ptr(keep), keep.shape[1], ptr(front["count"]), bsz, pre_max, ptr(front["boxes3d"]),
                                 ptr(front["scores"]), ptr(front["labels"]), ptr(front["anchor_index"]), int(post_max),
                                 int(sigmoid_scores), float(pad_score), int(pad_label), ptr(out["boxes"]), ptr(out["scores"]),
                                 ptr(out["labels"]),
                                 ptr(out["selected"]), ptr(out["num"]), _stream()), "pcdb_gather_kept")
    return out


def filter_points(points: torch.Tensor, frame_offsets: torch.Tensor, batch_size: int, calib: Optional[torch.Tensor] = None,
                  range_xy: Optional[torch.Tensor] = None, want_index: bool = False):
    """FOV + range filter of raw clouds on the device (kitti_dataset.py:714-717, 236-253; common_utils.py:47-51).

    points (N,C) f32 cuda, frame_offsets (B+1) i32 cuda, calib (B,26) f32 cuda or None (see include/pcdet_b200.h),
    range_xy (4,) f32 cuda [x_min,y_min,x_max,y_max] or None.  Returns (out_points (N,C) capacity-sized, order kept,
    out_offsets (B+1) i32, out_index (N,) i32 or None); rows beyond out_offsets[B] are undefined.  No sync."""
    _require_cuda(points, frame_offsets)
    assert points.dtype == torch.float32 and points.dim() == 2 and points.is_contiguous()
    assert frame_offsets.dtype == torch.int32 and frame_offsets.numel() == batch_size + 1
    n, c = points.shape
    dev = points.device
    if calib is not None:
        assert calib.dtype == torch.float32 and calib.is_cuda and calib.is_contiguous() and calib.shape == (batch_size, 26)
    if range_xy is not None:
        assert range_xy.dtype == torch.float32 and range_xy.is_cuda and range_xy.numel() == 4
    out = torch.empty((max(n, 1), c), dtype=torch.float32, device=dev)
    offs = torch.empty((batch_size + 1,), dtype=torch.int32, device=dev)
    index = torch.empty((max(n, 1),), dtype=torch.int32, device=dev) if want_index else None
    L = lib()
    ws = workspace(L.pcdb_filter_points_workspace_bytes(n), dev, "filter_points")
    check(L.pcdb_filter_points(ptr(points), n, c, ptr(frame_offsets), batch_size, ptr(calib), ptr(range_xy), ptr(out), ptr(offs),
                               ptr(index), ptr(ws), ws.numel(), _stream()), "pcdb_filter_points")
    return out, offs, index
