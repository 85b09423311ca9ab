"""CPU oracle for the PCDet voxel hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package ``pcdet_b200`` never does.

It wraps ``oracle/pcdet_oracle.c`` (see that file's header for what each function restates and where
in the reference / spconv v1.0 @ 8da6f96 it comes from) and composes the functions into the same
stages the reference runs:

  voxelize        spconv.utils.VoxelGenerator.generate        (dataset.py:163-181)
  vfe_mean        MeanVoxelFeatureExtractor.forward           (pcdet/models/vfe/vfe_utils.py:26-34)
  rulebook        spconv.ops.get_indice_pairs, CPU path       
  indice_conv     spconv.ops.indice_conv                      
  backbone8x      BackBone8x.forward                          (pcdet/models/rpn/rpn_backbone.py:7-103)
  nms / iou       iou3d_nms_cuda.*                            (pcdet/ops/iou3d_nms/src/*.cu|cpp)
  post_process    Detector3D.predict_boxes / post_processing  (detectors/detector3d.py:112-299, box_coder_utils.py:89-144)

PARITY UNPINNED for voxelize / rulebook / indice_conv (spconv is absent from /root/reference and has
no golden vectors there); the rotated IoU / NMS functions are pinned against the reference's own
kernel through ``tests/golden/nms_ref.npz`` (generated on a B200 by ``tests/golden/make_golden_gpu.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "liborc.so")
    src = os.path.join(_HERE, "pcdet_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["sh", os.path.join(_HERE, "build.sh")], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_box_overlap.restype = C.c_float
        _LIB.orc_iou_bev.restype = C.c_float
        _LIB.orc_iou_normal.restype = C.c_float
        _LIB.orc_box_overlap64.restype = C.c_double
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i3(v):
    if np.isscalar(v):
        v = [v] * 3
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


# --------------------------------------------------------------------------------------------
# voxelisation
# --------------------------------------------------------------------------------------------
class VoxelGenerator:
    """Restates spconv.utils.VoxelGenerator (v1.0)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000,
                 overflow_break=True):
        self.point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        self.voxel_size = np.array(voxel_size, dtype=np.float32)
        grid = (self.point_cloud_range[3:] - self.point_cloud_range[:3]) / self.voxel_size
        self.grid_size = np.round(grid).astype(np.int64)
        self.max_num_points = int(max_num_points)
        self.max_voxels = int(max_voxels)
        self.overflow_break = bool(overflow_break)
        self._lut = None

    def generate(self, points, max_voxels=None, return_point_idx=False):
        points = np.ascontiguousarray(points, dtype=np.float32)
        mv = int(max_voxels or self.max_voxels)
        if self._lut is None:
            self._lut = np.full(tuple(int(g) for g in self.grid_size[::-1]), -1, dtype=np.int32)
        n, c = points.shape
        P = self.max_num_points
        voxels = np.zeros((mv, P, c), dtype=np.float32)
        coors = np.zeros((mv, 3), dtype=np.int32)
        num = np.zeros((mv,), dtype=np.int32)
        pidx = np.full((mv, P), -1, dtype=np.int32)
        grid = np.ascontiguousarray(self.grid_size.astype(np.int32))
        nv = lib().orc_points_to_voxel(
            _p(points, C.c_float), n, c, _p(self.voxel_size, C.c_float),
            _p(self.point_cloud_range, C.c_float), _p(grid, C.c_int), P, mv,
            int(self.overflow_break), _p(self._lut, C.c_int32), _p(voxels, C.c_float),
            _p(coors, C.c_int32), _p(num, C.c_int32), _p(pidx, C.c_int32))
        out = (voxels[:nv], coors[:nv], num[:nv])
        if return_point_idx:
            out = out + (pidx[:nv],)
        return out


def vfe_mean(voxels, num_points):
    voxels = np.ascontiguousarray(voxels, dtype=np.float32)
    num_points = np.ascontiguousarray(num_points, dtype=np.int32)
    v, p, c = voxels.shape
    out = np.empty((v, c), dtype=np.float32)
    lib().orc_vfe_mean(_p(voxels, C.c_float), _p(num_points, C.c_int32), v, p, c, _p(out, C.c_float))
    return out


def pillar_vfe(voxels, num_points, coords, weight, scale, shift, voxel_size, pc_range, with_distance=False):
    """PillarFeatureNetOld2.forward in eval mode (pcdet/models/vfe/vfe_utils.py:168-215 with one PFNLayer,
    :61-116), fp32 like the reference: decorations, padding mask, Linear (no bias), BatchNorm as scale/shift,
    ReLU, max over the P slots (padded slots included, they are relu(shift))."""
    f32 = np.float32
    voxels = np.asarray(voxels, f32)
    n, p, c = voxels.shape
    numf = np.asarray(num_points).astype(f32).reshape(-1, 1, 1)
    vx, vy, vz = (float(v) for v in voxel_size)
    x_off, y_off, z_off = vx / 2 + pc_range[0], vy / 2 + pc_range[1], vz / 2 + pc_range[2]        # :162-164
    mean = voxels[:, :, :3].sum(axis=1, keepdims=True, dtype=f32) / numf                             # :179
    f_cluster = voxels[:, :, :3] - mean
    f_center = np.zeros_like(voxels[:, :, :3])
    cf = np.asarray(coords).astype(f32)
    f_center[:, :, 0] = voxels[:, :, 0] - (cf[:, 3:4] * f32(vx) + f32(x_off))                        # :185-187
    f_center[:, :, 1] = voxels[:, :, 1] - (cf[:, 2:3] * f32(vy) + f32(y_off))
    f_center[:, :, 2] = voxels[:, :, 2] - (cf[:, 1:2] * f32(vz) + f32(z_off))
    parts = [voxels, f_cluster, f_center]
    if with_distance:
        parts.append(np.sqrt((voxels[:, :, :3] ** 2).sum(axis=2, keepdims=True, dtype=f32)))
    feats = np.concatenate(parts, axis=-1).astype(f32)
    mask = (np.asarray(num_points).reshape(-1, 1) > np.arange(p).reshape(1, -1)).astype(f32)[:, :, None]   # :198-203
    feats = feats * mask
    x = feats @ np.asarray(weight, f32).T                                                          # PFNLayer :102
    if scale is not None:
        x = x * np.asarray(scale, f32) + np.asarray(shift, f32)
    elif shift is not None:
        x = x + np.asarray(shift, f32)
    return np.maximum(x, 0).max(axis=1).astype(f32)                                                # :106-111


def pillar_scatter(features, coords, batch_size, output_shape):
    """PointPillarsScatter.forward (pcdet/models/rpn/pillar_scatter.py:23-55)."""
    nz, ny, nx = (int(v) for v in output_shape)
    f = features.shape[1]
    canvas = np.zeros((batch_size, f, nz * ny * nx), np.float32)
    coords = np.asarray(coords)
    for b in range(batch_size):
        m = coords[:, 0] == b
        idx = coords[m, 1] * nz + coords[m, 2] * nx + coords[m, 3]
        canvas[b][:, idx] = features[m].T
    return canvas.reshape(batch_size, f * nz, ny, nx)


def roiaware_pool3d(rois, pts, pts_feature, out_size, max_pts_each_voxel=128, pool_method="max"):
    """roiaware_pool3d_utils.RoIAwarePool3dFunction.forward (pcdet/ops/roiaware_pool3d): returns
    (pooled_features (N,ox,oy,oz,C), argmax, pts_idx_of_voxels)."""
    rois = np.ascontiguousarray(rois, np.float32); pts = np.ascontiguousarray(pts, np.float32)
    feat = np.ascontiguousarray(pts_feature, np.float32)
    ox, oy, oz = (out_size,) * 3 if isinstance(out_size, int) else tuple(out_size)
    n, c = rois.shape[0], feat.shape[1]
    pooled = np.zeros((n, ox, oy, oz, c), np.float32)
    argmax = np.zeros((n, ox, oy, oz, c), np.int32)
    idx = np.zeros((n, ox, oy, oz, max_pts_each_voxel), np.int32)
    lib().orc_roiaware_pool3d(_p(rois, C.c_float), n, _p(pts, C.c_float), pts.shape[0], _p(feat, C.c_float), c, ox, oy, oz,
                              max_pts_each_voxel, {"max": 0, "avg": 1}[pool_method], _p(argmax, C.c_int32),
                              _p(idx, C.c_int32), _p(pooled, C.c_float))
    return pooled, argmax, idx


def points_in_boxes(points, boxes):
    """roiaware_pool3d_utils.points_in_boxes_gpu: points (B,M,3), boxes (B,T,7) -> (B,M) int32, -1 = background."""
    points = np.ascontiguousarray(points, np.float32); boxes = np.ascontiguousarray(boxes, np.float32)
    out = np.empty(points.shape[:2], np.int32)
    lib().orc_points_in_boxes(_p(boxes, C.c_float), boxes.shape[0], boxes.shape[1], _p(points, C.c_float), points.shape[1],
                              _p(out, C.c_int32))
    return out


def collate(frames):
    """dataset.py:266-299: concat voxels/num_points, prepend the batch index to coordinates."""
    voxels = np.concatenate([f[0] for f in frames], axis=0)
    coords = np.concatenate(
        [np.pad(f[1], ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, f in enumerate(frames)],
        axis=0).astype(np.int32)
    num = np.concatenate([f[2] for f in frames], axis=0)
    return voxels, coords, num


# --------------------------------------------------------------------------------------------
# rulebook
# --------------------------------------------------------------------------------------------
def conv_output_size(in_shape, ksize, stride, pad, dil):
    """spconv.ops.get_conv_output_size."""
    return [int((i + 2 * p - d * (k - 1) - 1) // s + 1) for i, k, s, p, d in zip(in_shape, ksize, stride, pad, dil)]


_grids = {}


def _grid(cells):
    g = _grids.get(cells)
    if g is None:
        if len(_grids) > 4:
            _grids.clear()
        g = np.zeros((cells,), dtype=np.int32)  # lazily-mapped zero pages
        _grids[cells] = g
    return g


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                     subm=False):
    """Returns (out_ids (n_out,4), pairs (K,2,n_in) -1 padded, pair_num (K), out_shape)."""
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    ks, st, pd, dl = _i3(ksize), _i3(stride), _i3(padding), _i3(dilation)
    shape = _i3(list(spatial_shape))
    n = indices.shape[0]
    K = int(ks.prod())
    pairs = np.full((K, 2, max(n, 1)), -1, dtype=np.int32)
    num = np.zeros((K,), dtype=np.int32)
    if subm:
        out_shape = shape.copy()
        grid = _grid(int(batch_size) * int(shape.prod()))
        lib().orc_rulebook_subm(_p(indices, C.c_int32), n, int(batch_size), _p(shape, C.c_int),
                                _p(ks, C.c_int), _p(dl, C.c_int), _p(grid, C.c_int32),
                                _p(pairs, C.c_int32), _p(num, C.c_int32))
        return indices, pairs[:, :, :n] if n else pairs[:, :, :0], num, [int(v) for v in out_shape]
    out_shape = _i3(conv_output_size(shape, ks, st, pd, dl))
    grid = _grid(int(batch_size) * int(out_shape.prod()))
    out_ids = np.zeros((max(n * K, 1), 4), dtype=np.int32)
    n_out = lib().orc_rulebook_conv(_p(indices, C.c_int32), n, int(batch_size), _p(shape, C.c_int),
                                    _p(out_shape, C.c_int), _p(ks, C.c_int), _p(st, C.c_int),
                                    _p(pd, C.c_int), _p(dl, C.c_int), _p(grid, C.c_int32),
                                    _p(pairs, C.c_int32), _p(num, C.c_int32), _p(out_ids, C.c_int32))
    return out_ids[:n_out].copy(), pairs[:, :, :n] if n else pairs[:, :, :0], num, [int(v) for v in out_shape]


def pairs_to_sets(out_ids, in_ids, pairs, pair_num):
    """Canonical form for order-insensitive comparison: per offset, the set of
    ((b,z,y,x)_in, (b,z,y,x)_out) coordinate pairs."""
    res = []
    for k in range(pairs.shape[0]):
        n = int(pair_num[k])
        i = pairs[k, 0, :n]
        o = pairs[k, 1, :n]
        rows = np.concatenate([in_ids[i], out_ids[o]], axis=1) if n else np.zeros((0, 8), np.int32)
        rows = rows[np.lexsort(rows.T[::-1])] if n else rows
        res.append(rows)
    return res


# --------------------------------------------------------------------------------------------
# sparse convolution
# --------------------------------------------------------------------------------------------
def indice_conv(features, filters, pairs, pair_num, n_out, subm=False, inverse=False, acc64=False):
    """features (n_in,c_in) f32; filters (kz,ky,kx,c_in,c_out) f32 -> (n_out,c_out) f32."""
    features = np.ascontiguousarray(features, dtype=np.float32)
    filters = np.ascontiguousarray(filters, dtype=np.float32)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    pair_num = np.ascontiguousarray(pair_num, dtype=np.int32)
    c_in, c_out = filters.shape[-2:]
    K = pairs.shape[0]
    out = np.zeros((n_out, c_out), dtype=np.float32)
    lib().orc_indice_conv(_p(features, C.c_float), _p(filters, C.c_float), _p(pairs, C.c_int32),
                          _p(pair_num, C.c_int32), pairs.shape[2], n_out, c_in, c_out, K, int(subm),
                          int(inverse), int(acc64), _p(out, C.c_float))
    return out


def indice_conv_mm(features, filters, pairs, pair_num, n_out, subm=False, inverse=False):
    """Same arithmetic through torch CPU matmuls (gather -> mm -> index_add), i.e. the shape of the
    reference's own loop (App. A.4) with MKL in place of cuBLAS.  Used as the timed CPU baseline."""
    import torch
    f = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
    c_in, c_out = filters.shape[-2:]
    w = torch.from_numpy(np.ascontiguousarray(filters, dtype=np.float32)).reshape(-1, c_in, c_out)
    out = torch.zeros((n_out, c_out), dtype=torch.float32)
    centre = -1
    if subm:
        centre = int(np.argmax(pair_num))
        out = f @ w[centre]
    pt = torch.from_numpy(np.ascontiguousarray(pairs)).long()
    for k in range(w.shape[0]):
        n = int(pair_num[k])
        if k == centre or n <= 0:
            continue
        gi = pt[k, 1 if inverse else 0, :n]
        go = pt[k, 0 if inverse else 1, :n]
        out.index_add_(0, go, f.index_select(0, gi) @ w[k])
    return out.numpy()


def indice_conv_backward(features, filters, out_bp, pairs, pair_num, subm=False, inverse=False):
    """spconv v1.0 indiceConvBackward restated: (input_bp (n_in, c_in), filters_bp like filters), fp64 sums.

    filtersGrad = 0, inputGrad = 0; SubM centre offset: filtersGrad[c] = features^T @ outGrad, inputGrad = outGrad @ filters[c]^T;
    every other offset k with n_k > 0: gather both sides by the pairs of k, filtersGrad[k] = inBuf^T @ outBuf,
    inBuf = outBuf @ filters[k]^T scatter-added into inputGrad (within one offset every input row appears at most once)."""
    f = np.asarray(features, dtype=np.float64)
    g = np.asarray(out_bp, dtype=np.float64)
    w = np.asarray(filters, dtype=np.float64)
    c_in, c_out = w.shape[-2:]
    w3 = w.reshape(-1, c_in, c_out)
    fb = np.zeros_like(w3)
    ib = np.zeros_like(f)
    centre = -1
    if subm:
        centre = int(np.argmax(pair_num))
        fb[centre] = f.T @ g
        ib += g @ w3[centre].T
    for k in range(w3.shape[0]):
        n = int(pair_num[k])
        if k == centre or n <= 0:
            continue
        gi = pairs[k, 1 if inverse else 0, :n]
        go = pairs[k, 0 if inverse else 1, :n]
        fb[k] = f[gi].T @ g[go]
        np.add.at(ib, gi, g[go] @ w3[k].T)
    return ib.astype(np.float32), fb.reshape(w.shape).astype(np.float32)


def indice_maxpool(features, pairs, pair_num, n_out):
    """spconv indice_maxpool: output starts from zeros, running maximum over the pairs."""
    features = np.asarray(features, dtype=np.float32)
    out = np.zeros((n_out, features.shape[1]), dtype=np.float32)
    for k in range(pairs.shape[0]):
        n = int(pair_num[k])
        if n:
            np.maximum.at(out, pairs[k, 1, :n], features[pairs[k, 0, :n]])
    return out


def to_dense(features, indices, spatial_shape, batch_size):
    """SparseConvTensor.dense(): (B,C,D,H,W)."""
    c = features.shape[1]
    out = np.zeros((batch_size, *[int(s) for s in spatial_shape], c), dtype=features.dtype)
    idx = indices.astype(np.int64)
    out[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = features
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3))


# BackBone8x topology, rpn_backbone.py:12-51: (name, kind, c_in, c_out, ksize, stride, padding, key)
BACKBONE8X = [
    ("conv_input.0", "subm", None, 16, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm1"),
    ("conv1.0.0", "subm", 16, 16, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm1"),
    ("conv2.0.0", "spconv", 16, 32, (3, 3, 3), (2, 2, 2), (1, 1, 1), "spconv2"),
    ("conv2.1.0", "subm", 32, 32, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm2"),
    ("conv2.2.0", "subm", 32, 32, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm2"),
    ("conv3.0.0", "spconv", 32, 64, (3, 3, 3), (2, 2, 2), (1, 1, 1), "spconv3"),
    ("conv3.1.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm3"),
    ("conv3.2.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm3"),
    ("conv4.0.0", "spconv", 64, 64, (3, 3, 3), (2, 2, 2), (0, 1, 1), "spconv4"),
    ("conv4.1.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm4"),
    ("conv4.2.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm4"),
    ("conv_out.0", "spconv", 64, 128, (3, 1, 1), (2, 1, 1), (0, 0, 0), "spconv_down2"),
]


def backbone8x(features, indices, spatial_shape, batch_size, weights, bn=None, conv=indice_conv,
               collect=None, bf16=False):
    """BackBone8x.forward restated (rpn_backbone.py:54-77): 12 x (sparse conv -> BN(eval) -> ReLU), dense.

    weights: {layer name: (kz,ky,kx,cin,cout) f32}; bn: {layer name: (scale, shift)} folded eval-mode
    BatchNorm1d(eps=1e-3) (identity-ish at init: scale=1/sqrt(1+1e-3), shift=0).
    bf16=True rounds features and weights to bfloat16 between layers (storage model of the bf16 mode).
    Returns dense (B, C*D, H, W) and fills ``collect`` with per-layer (features, indices, pairs...)."""
    def rb(x):
        if not bf16:
            return x
        import torch
        return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()

    cache = {}
    x, idx, shape = rb(np.asarray(features, np.float32)), np.asarray(indices, np.int32), list(spatial_shape)
    for name, kind, _cin, _cout, ks, st, pd, key in BACKBONE8X:
        if key in cache:
            out_ids, pairs, num, out_shape = cache[key]
        else:
            out_ids, pairs, num, out_shape = get_indice_pairs(idx, batch_size, shape, ks, st, pd, 1,
                                                              subm=(kind == "subm"))
            cache[key] = (out_ids, pairs, num, out_shape)
        y = conv(x, rb(weights[name]), pairs, num, out_ids.shape[0], subm=(kind == "subm"))
        if bn is not None and name in bn:
            scale, shift = bn[name]
            y = y * scale[None, :] + shift[None, :]
        else:
            y = y * np.float32(1.0 / np.sqrt(1.0 + 1e-3))
        y = rb(np.maximum(y, 0).astype(np.float32))
        if collect is not None:
            collect[name] = dict(features=y, indices=out_ids, pairs=pairs, pair_num=num, shape=out_shape,
                                 in_indices=idx)
        x, idx, shape = y, out_ids, out_shape
    dense = to_dense(x, idx, shape, batch_size)
    n, c, d, h, w = dense.shape
    return dense.reshape(n, c * d, h, w)


def unet_v2(features, indices, spatial_shape, batch_size, sd, conv=indice_conv_mm):
    """UNetV2.forward restated (pcdet/models/rpn/rpn_unet.py:464-529, inference part) from a state dict `sd`
    ({name: numpy}) with the reference's keys; BatchNorm1d(eps=1e-3) in eval mode.  Returns the dict of
    rpn_unet.py:483,499-505 (spatial_features, u_seg_preds, u_reg_preds, seg_features) as numpy arrays."""
    f32 = np.float32
    rb = {}

    def pairs_of(key, idx, shape, kind, ks=3, st=1, pd=0):
        if key not in rb:
            rb[key] = get_indice_pairs(idx, batch_size, shape, ks, st, pd, 1, subm=(kind == "subm")) + (idx,)
        return rb[key]

    def bn(x, stem):
        scale = sd[stem + ".weight"] / np.sqrt(sd[stem + ".running_var"] + f32(1e-3))
        return (x * scale[None, :] + (sd[stem + ".bias"] - sd[stem + ".running_mean"] * scale)[None, :]).astype(f32)

    def block(x, idx, shape, stem, key, kind, ks=3, st=1, pd=0):
        """post_act_block (rpn_unet.py:435-462): conv -> BN -> ReLU; stem = SparseSequential holding (conv, bn, relu)"""
        w = sd[stem + ".0.weight"]
        if kind == "inverse":
            out_ids, pairs, num, _, in_idx = rb[key]                      # the strided conv's rulebook, roles swapped
            y = conv(x, w, pairs, num, in_idx.shape[0], inverse=True)
            out_idx, out_shape = in_idx, None
        else:
            out_ids, pairs, num, out_shape, _ = pairs_of(key, idx, shape, kind, ks, st, pd)
            y = conv(x, w, pairs, num, out_ids.shape[0], subm=(kind == "subm"))
            out_idx = out_ids
        return np.maximum(bn(y, stem + ".1"), 0), out_idx, out_shape

    def basic_block(x, idx, shape, stem, key):
        """SparseBasicBlock (resnet_utils.py:29-48)"""
        _, pairs, num, _, _ = pairs_of(key, idx, shape, "subm")
        y = np.maximum(bn(conv(x, sd[stem + ".conv1.weight"], pairs, num, idx.shape[0], subm=True), stem + ".bn1"), 0)
        y = bn(conv(y, sd[stem + ".conv2.weight"], pairs, num, idx.shape[0], subm=True), stem + ".bn2")
        return np.maximum(y + x, 0)

    x0, i1, s1 = np.asarray(features, f32), np.asarray(indices, np.int32), list(spatial_shape)
    x, _, _ = block(x0, i1, s1, "conv_input", "subm1", "subm")
    c1, _, _ = block(x, i1, s1, "conv1.0", "subm1", "subm")
    x, i2, s2 = block(c1, i1, s1, "conv2.0", "spconv2", "spconv", 3, 2, 1)
    x, _, _ = block(x, i2, s2, "conv2.1", "subm2", "subm")
    c2, _, _ = block(x, i2, s2, "conv2.2", "subm2", "subm")
    x, i3, s3 = block(c2, i2, s2, "conv3.0", "spconv3", "spconv", 3, 2, 1)
    x, _, _ = block(x, i3, s3, "conv3.1", "subm3", "subm")
    c3, _, _ = block(x, i3, s3, "conv3.2", "subm3", "subm")
    x, i4, s4 = block(c3, i3, s3, "conv4.0", "spconv4", "spconv", 3, 2, (0, 1, 1))
    x, _, _ = block(x, i4, s4, "conv4.1", "subm4", "subm")
    c4, _, _ = block(x, i4, s4, "conv4.2", "subm4", "subm")
    o, i5, s5 = block(c4, i4, s4, "conv_out", "spconv_down2", "spconv", (3, 1, 1), (2, 1, 1), 0)
    dense = to_dense(o, i5, s5, batch_size)
    n, c, d, h, w = dense.shape

    def ur_block(lat, bottom, idx, shape, t, m, inv, key, inv_key):
        """UR_block_forward (rpn_unet.py:420-428)"""
        xt = basic_block(lat, idx, shape, t, key)
        cat = np.concatenate([bottom, xt], axis=1)
        xm, _, _ = block(cat, idx, shape, m, key, "subm")
        red = cat.reshape(cat.shape[0], xm.shape[1], -1).sum(axis=2)              # channel_reduction :430-433
        y = (xm + red).astype(f32)
        if inv_key is None:
            return block(y, idx, shape, inv, key, "subm")[0]
        return block(y, idx, shape, inv, inv_key, "inverse")[0]

    u4 = ur_block(c4, c4, i4, s4, "conv_up_t4", "conv_up_m4", "inv_conv4", "subm4", "spconv4")
    u3 = ur_block(c3, u4, i3, s3, "conv_up_t3", "conv_up_m3", "inv_conv3", "subm3", "spconv3")
    u2 = ur_block(c2, u3, i2, s2, "conv_up_t2", "conv_up_m2", "inv_conv2", "subm2", "spconv2")
    u1 = ur_block(c1, u2, i1, s1, "conv_up_t1", "conv_up_m1", "conv5.0", "subm1", None)
    return {"spatial_features": dense.reshape(n, c * d, h, w), "seg_features": u1,
            "u_seg_preds": u1 @ sd["seg_cls_layer.weight"].T + sd["seg_cls_layer.bias"],
            "u_reg_preds": u1 @ sd["seg_reg_layer.weight"].T + sd["seg_reg_layer.bias"]}


# --------------------------------------------------------------------------------------------
# rotated IoU / NMS
# --------------------------------------------------------------------------------------------
def boxes3d_to_bev(boxes3d):
    """pcdet/utils/box_utils.py:237-250."""
    b = np.asarray(boxes3d, dtype=np.float32)
    out = np.empty((b.shape[0], 5), dtype=np.float32)
    half_l, half_w = b[:, 4] / np.float32(2), b[:, 3] / np.float32(2)
    out[:, 0], out[:, 1] = b[:, 0] - half_w, b[:, 1] - half_l
    out[:, 2], out[:, 3] = b[:, 0] + half_w, b[:, 1] + half_l
    out[:, 4] = b[:, 6]
    return out


def _mat(fn, a, b, dtype):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty((a.shape[0], b.shape[0]), dtype=dtype)
    ct = C.c_float if dtype == np.float32 else C.c_double
    fn(_p(a, C.c_float), a.shape[0], _p(b, C.c_float), b.shape[0], _p(out, ct))
    return out


def boxes_overlap_bev(a, b):
    return _mat(lib().orc_boxes_overlap_bev, a, b, np.float32)


def boxes_iou_bev(a, b):
    return _mat(lib().orc_boxes_iou_bev, a, b, np.float32)


def boxes_iou_bev64(a, b):
    return _mat(lib().orc_boxes_iou_bev64, a, b, np.float64)


def nms_sorted(boxes, thresh, normal=False):
    """iou3d_nms_cuda.nms_gpu on score-sorted boxes: kept positions (int64)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    keep = np.empty((boxes.shape[0],), dtype=np.int64)
    n = lib().orc_nms(_p(boxes, C.c_float), boxes.shape[0], C.c_float(thresh), int(normal), _p(keep, C.c_int64))
    return keep[:n].copy()


def nms(boxes, scores, thresh, pre_maxsize=None, normal=False):
    """iou3d_nms_utils.nms_gpu (iou3d_nms_utils.py:62-78): kept ORIGINAL indices in score order."""
    order = np.argsort(-np.asarray(scores), kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    keep = nms_sorted(np.asarray(boxes)[order], thresh, normal)
    return order[keep]


# --------------------------------------------------------------------------------------------
# post-processing front: decode + threshold + top-k (class-agnostic path)
# --------------------------------------------------------------------------------------------
def decode_boxes(box_preds, anchors, dir_cls_preds=None, num_dir_bins=2, dir_offset=0.0, dir_limit_offset=0.0,
                 use_binary_dir_classifier=False):
    """ResidualCoder.decode_torch + decode_with_head_direction_torch (pcdet/utils/box_coder_utils.py:89-144),
    fp32 operation by operation.  box_preds (..., 7) residuals, anchors broadcastable to it."""
    f = np.float32
    t = np.asarray(box_preds, dtype=f)
    a = np.broadcast_to(np.asarray(anchors, dtype=f), t.shape)
    xa, ya, za, wa, la, ha, ra = [a[..., i] for i in range(7)]
    xt, yt, zt, wt, lt, ht, rt = [t[..., i] for i in range(7)]
    za = za + ha / f(2)                                            # :99
    diagonal = np.sqrt(la * la + wa * wa)                          # :101
    xg = xt * diagonal + xa                                        # :102-104
    yg = yt * diagonal + ya
    zg = zt * ha + za
    lg = np.exp(lt) * la                                           # :106-108
    wg = np.exp(wt) * wa
    hg = np.exp(ht) * ha
    rg = rt + ra                                                   # :109
    zg = zg - hg / f(2)                                            # :111
    if dir_cls_preds is not None:
        d = np.asarray(dir_cls_preds, dtype=f)
        d = d.reshape(t.shape[:-1] + (d.shape[-1],))
        dir_labels = np.argmax(d, axis=-1)                         # first maximum, like torch.max
        if use_binary_dir_classifier:                              # :126-133
            opp = (rg > 0) ^ (dir_labels != 0)
            rg = rg + np.where(opp, f(np.pi), f(0))
        else:                                                      # :135-141, common_utils.py:95-96
            period = f(2 * np.pi / num_dir_bins)
            val = rg - f(dir_offset)
            dir_rot = val - np.floor(val / period + f(dir_limit_offset)) * period
            rg = dir_rot + f(dir_offset) + period * dir_labels.astype(f)
    return np.stack([xg, yg, zg, wg, lg, hg, rg], axis=-1).astype(f)


def sigmoid32(x):
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


def class_agnostic_select(cls_preds, score_thresh, pre_max):
    """Detector3D.post_processing / class_agnostic_nms up to the NMS call (detector3d.py:193-197, 278-288) for ONE
    frame: (selected anchor indices in score order, their rank scores, labels in [1, C]).  Ties: lower anchor first."""
    cls = np.asarray(cls_preds, dtype=np.float32)
    rank = cls.max(axis=-1)
    labels = cls.argmax(axis=-1) + 1
    cand = np.nonzero(sigmoid32(rank) >= np.float32(score_thresh))[0]
    order = cand[np.argsort(-rank[cand], kind="stable")][:pre_max]
    return order, rank[order], labels[order].astype(np.int32)


def post_process(cls_preds, box_preds, anchors, dir_cls_preds=None, score_thresh=0.1, nms_thresh=0.01, pre_max=4096,
                 post_max=500, num_dir_bins=2, dir_offset=0.0, dir_limit_offset=0.0, use_binary_dir_classifier=False):
    """Detector3D.predict_boxes + post_processing, class-agnostic NMS, USE_RAW_SCORE (detector3d.py:112-128, 156-223,
    278-299) for a batch: list of dict(boxes (n,7), scores (n,), labels (n,), selected (n,) anchor indices)."""
    out = []
    for b in range(np.asarray(cls_preds).shape[0]):
        sel, scores, labels = class_agnostic_select(cls_preds[b], score_thresh, pre_max)
        boxes = decode_boxes(np.asarray(box_preds[b])[sel], np.asarray(anchors)[sel],
                             None if dir_cls_preds is None else np.asarray(dir_cls_preds[b]).reshape(len(box_preds[b]), -1)[sel],
                             num_dir_bins, dir_offset, dir_limit_offset, use_binary_dir_classifier)
        keep = nms_sorted(boxes3d_to_bev(boxes), nms_thresh)[:post_max] if len(sel) else np.zeros((0,), np.int64)
        out.append(dict(boxes=boxes[keep], scores=scores[keep], labels=labels[keep], selected=sel[keep],
                        pre_nms=dict(boxes=boxes, scores=scores, labels=labels, selected=sel)))
    return out


# --------------------------------------------------------------------------------------------
# ingest: camera-FOV and range filters in front of the voxelizer
# --------------------------------------------------------------------------------------------
def lidar_to_rect_matrix(V2C, R0):
    """(4,3) matrix M with [x y z 1] . M = rectified camera coordinates (pcdet/utils/calibration.py:72)."""
    return np.dot(np.asarray(V2C, np.float32).T, np.asarray(R0, np.float32).T)


def fov_flag(points, V2C, R0, P2, img_shape):
    """Calibration.lidar_to_rect + rect_to_img (calibration.py:66-85) + get_fov_flag (kitti_dataset.py:236-253):
    True for the points that project into the (h, w) image with non-negative depth.  Also returns the pixel
    coordinates and depths in float64 so that a test can tell which decisions sit on a border."""
    f = np.float32
    xyz1 = np.concatenate([np.asarray(points, f)[:, :3], np.ones((len(points), 1), f)], axis=1)
    rect = np.dot(xyz1, lidar_to_rect_matrix(V2C, R0))
    rect1 = np.concatenate([rect, np.ones((len(points), 1), f)], axis=1)
    P2 = np.asarray(P2, f)
    hom = np.dot(rect1, P2.T)
    u, v = hom[:, 0] / rect[:, 2], hom[:, 1] / rect[:, 2]
    depth = hom[:, 2] - P2[2, 3]
    flag = (u >= 0) & (u < img_shape[1]) & (v >= 0) & (v < img_shape[0]) & (depth >= 0)
    # fp64 shadow computation for border detection
    rect64 = xyz1.astype(np.float64) @ (np.asarray(V2C, np.float64).T @ np.asarray(R0, np.float64).T)
    hom64 = np.concatenate([rect64, np.ones((len(points), 1))], axis=1) @ P2.astype(np.float64).T
    with np.errstate(divide="ignore", invalid="ignore"):
        shadow = np.stack([hom64[:, 0] / rect64[:, 2], hom64[:, 1] / rect64[:, 2], hom64[:, 2] - float(P2[2, 3])], axis=1)
    return flag, shadow


def mask_points_by_range(points, limit_range):
    """pcdet/utils/common_utils.py:47-51 (x and y only, both ends inclusive)."""
    p = np.asarray(points)
    m = (p[:, 0] >= limit_range[0]) & (p[:, 0] <= limit_range[3]) & (p[:, 1] >= limit_range[1]) & (p[:, 1] <= limit_range[4])
    return m


def filter_points(frames, calibs=None, img_shapes=None, pc_range=None):
    """KittiDataset.__getitem__ (kitti_dataset.py:714-717) + dataset.py:184 for a list of frames: filtered frames,
    order preserved.  calibs: list of dict(V2C, R0, P2) or None."""
    out = []
    for b, pts in enumerate(frames):
        m = np.ones((len(pts),), bool)
        if pc_range is not None:
            m &= mask_points_by_range(pts, pc_range)
        if calibs is not None:
            m &= fov_flag(pts, calibs[b]["V2C"], calibs[b]["R0"], calibs[b]["P2"], img_shapes[b])[0]
        out.append(np.asarray(pts)[m])
    return out


def proposal_layer(cls_preds, boxes, pre_max, post_max, nms_thresh, normal=False):
    """pcdet/models/model_utils/proposal_layer.py:7-68 (batch_idx None): per frame class max, top-`pre_max` by score,
    NMS, first `post_max`; rois zero padded, raw scores padded with -100000, labels padded with 1."""
    cls_preds, boxes = np.asarray(cls_preds, np.float32), np.asarray(boxes, np.float32)
    bsz = cls_preds.shape[0]
    rois = np.zeros((bsz, post_max, boxes.shape[-1]), np.float32)
    raw = np.full((bsz, post_max), -100000, np.float32)
    labels = np.ones((bsz, post_max), np.int64)
    for b in range(bsz):
        order, sc, lab = class_agnostic_select(cls_preds[b], 0.0, pre_max)       # sigmoid >= 0: every row is a candidate
        keep = nms_sorted(boxes3d_to_bev(boxes[b][order]), nms_thresh, normal)[:post_max]
        n = len(keep)
        rois[b, :n], raw[b, :n], labels[b, :n] = boxes[b][order][keep], sc[keep], lab[keep]
    return dict(rois=rois, roi_raw_scores=raw, roi_labels=labels)
