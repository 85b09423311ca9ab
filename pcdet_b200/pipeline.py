"""SECOND hot path, device resident and free of host synchronisation.

    raw points (B frames)  ->  voxel hash + mean VFE  ->  BackBone8x (8 rulebook builds, 12 fused
    conv+BN+ReLU kernels, dense)  ->  [RPN head: out of scope, stock cuDNN]  ->  rotated NMS

This is the same sequence SECONDNet.forward_rpn / predict_boxes drive through the module API
(pcdet/models/detectors/second_net.py:13-44, detector3d.py:278-299), but with every data-dependent
size (voxel count, active sites per level) kept in device memory: buffers are allocated once at
capacity, kernels read the counts from the device, and a whole step is a fixed launch sequence that
can be captured into a CUDA graph.  The module API (pcdet_b200.spconv) calls the very same kernels
with exact shapes, paying one device->host sync per strided rulebook build.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import functional as F
from ._lib import BF16, CONV_PDL, CONV_SHALLOW_RING, EPI_RELU, F32, RB_CLEARED, RB_UNDONE, WEIGHT_PACKED, check, f32xN, i32x3, lib, ptr
from .backbone import BACKBONE8X_LAYERS, BackBone8x


@dataclass
class HotPathConfig:
    voxel_size: Sequence[float] = (0.05, 0.05, 0.1)
    point_cloud_range: Sequence[float] = (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)
    max_num_points: int = 5
    max_voxels: int = 40000            # per frame, reference semantics (second.yaml:26)
    batch_size: int = 4
    num_point_features: int = 4
    dtype: torch.dtype = torch.bfloat16
    # capacities (rows) of the per-level buffers for the whole batch; None = derive from max_points_total
    max_points_total: int = 4 * 24000
    # [L1, L2, L3, L4, Lout] row capacities of the five levels.  Default: n1 = min(max_points_total, B * max_voxels) voxels and
    # [n1, 1.6 n1, n1, n1/2, n1/2] -- a stride-2 3x3x3 conv can only dilate the active set by the factor lidar surfaces
    # show (1.36 on the synthetic KITTI frames, 0.85 on nuScenes; the worst case is 8x).  Exceeding a capacity never writes
    # out of bounds: the level's overflow flag is raised (step()["level_counts"], checked by HostRunner.result) and the
    # sites beyond the capacity are dropped.
    level_capacity: Optional[Sequence[int]] = None
    nms_boxes_per_frame: int = 4096     # NMS_PRE_MAXSIZE_LAST (second.yaml:158)
    nms_keep_per_frame: int = 500       # NMS_POST_MAXSIZE_LAST
    nms_thresh: float = 0.01
    overflow_break: bool = True
    conv_algo: int = 0                  # 0 auto, 1 SIMT, 2 tcgen05
    # two-stage shared-memory ring in the tcgen05 convolutions (PCDB_CONV_SHALLOW_RING): for instances that run next to
    # other instances on the same GPU (several steps in flight); costs ~4 % of a step that has the GPU to itself
    conv_shallow_ring: bool = False
    # every rulebook of the backbone in four launches (pcdb_rulebook_chain): same site and pair SETS, rows of levels >= 1
    # in ascending (b,z,y,x) order (the reference's CUDA rulebook); False = one build per map in the row order of the
    # reference's CPU loop (24 launches, a chain)
    rulebook_chain: bool = True


class SecondHotPath:
    def __init__(self, cfg: HotPathConfig, backbone: BackBone8x, device="cuda"):
        self.cfg = cfg
        self.dev = torch.device(device)
        self.lib = lib()
        B = cfg.batch_size
        self.grid_xyz = F.grid_size(cfg.voxel_size, cfg.point_cloud_range)
        # second_net.py:10: sparse_shape = grid_size[::-1] + [1, 0, 0]
        self.sparse_shape = [int(self.grid_xyz[2]) + 1, int(self.grid_xyz[1]), int(self.grid_xyz[0])]
        n1 = min(cfg.max_points_total, B * cfg.max_voxels)
        caps = list(cfg.level_capacity) if cfg.level_capacity else [n1, int(n1 * 1.6), n1, n1 // 2, n1 // 2]
        self.caps = [max(int(c), 64) for c in caps]
        self.tc = cfg.dtype == torch.bfloat16
        self.cin0 = 16 if self.tc else cfg.num_point_features   # bf16: pad 4 -> 16 channels for the MMA K step
        self._prepare_weights(backbone)
        self._allocate()

    # ------------------------------------------------------------------------------------------
    def _prepare_weights(self, backbone: BackBone8x):
        dt = self.cfg.dtype
        self.layers = []
        for (stem, kind, c_in, c_out, ks, st, pd, key), (_s, conv, bn) in zip(BACKBONE8X_LAYERS, backbone.conv_modules()):
            w = conv.weight.detach().to(self.dev, torch.float32)
            K = ks[0] * ks[1] * ks[2]
            w = w.reshape(K, w.shape[-2], w.shape[-1])
            if stem == "conv_input.0" and self.cin0 != w.shape[1]:
                wp = torch.zeros((K, self.cin0, w.shape[2]), dtype=torch.float32, device=self.dev)
                wp[:, :w.shape[1]] = w
                w = wp
            scale = (bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)).to(self.dev)
            shift = (bn.bias.detach().float().to(self.dev) - bn.running_mean.detach().float().to(self.dev) * scale)
            use_tc = self.cfg.conv_algo != 1 and F.tc_eligible(dt, w.shape[1], w.shape[2], K)
            wd = w.to(dt)
            self.layers.append(dict(stem=stem, kind=kind, K=K, c_in=w.shape[1], c_out=w.shape[2], ks=list(ks),
                                    st=list(st), pd=list(pd), key=key,
                                    w=(F.pack_conv_weights(wd.contiguous()) if use_tc else wd.contiguous()),
                                    wflags=(WEIGHT_PACKED if use_tc else 0),
                                    scale=scale.contiguous(), shift=shift.contiguous()))

    def _allocate(self):
        cfg, dev, B = self.cfg, self.dev, self.cfg.batch_size
        dt = cfg.dtype
        i32 = dict(dtype=torch.int32, device=dev)
        c1, c2, c3, c4, c5 = self.caps
        self.level_of_key = {"subm1": 0, "spconv2": 1, "subm2": 1, "spconv3": 2, "subm3": 2, "spconv4": 3,
                             "subm4": 3, "spconv_down2": 4}
        # level shapes
        shapes = [self.sparse_shape]
        for stem, kind, *_rest in BACKBONE8X_LAYERS:
            if kind == "spconv":
                _, _, _, _, ks, st, pd, _ = next(l for l in BACKBONE8X_LAYERS if l[0] == stem)
                shapes.append(F.conv_output_size(shapes[-1], ks, st, pd, (1, 1, 1)))
        self.shapes = shapes                     # 5 levels
        self.coords = [torch.empty((c, 4), **i32) for c in self.caps]
        # [count, overflow flag] of levels 1-4 (level 0's count is voxel_offsets[B]); one block, so that a host caller
        # fetches all of them with one copy
        self.counts_all = torch.zeros((5, 2), **i32)
        self.counts = [None] + [self.counts_all[lv] for lv in range(1, 5)]
        self.voxel_offsets = torch.zeros((B + 1,), **i32)
        self.num_points = torch.empty((c1,), **i32)
        self.nbr = {}
        for stem, kind, _ci, _co, ks, _st, _pd, key in BACKBONE8X_LAYERS:
            if key not in self.nbr:
                K = ks[0] * ks[1] * ks[2]
                self.nbr[key] = torch.full((K, self.caps[self.level_of_key[key]]), -1, **i32)
        # two ping-pong feature buffers per level, sized for the widest channel count used there
        widths = [16, 32, 64, 64, 128]
        self.feat = [[torch.empty((c, max(w, self.cin0)), dtype=dt, device=dev) for _ in range(2)]
                     for c, w in zip(self.caps, widths)]
        self.vfe = torch.empty((c1, self.cin0), dtype=dt, device=dev)
        L = self.lib
        nbytes = max(L.pcdb_voxelize_workspace_bytes(cfg.max_points_total, B, cfg.max_num_points, cfg.max_voxels),
                     max(L.pcdb_rulebook_workspace_bytes(a, 27, b) for a, b in
                         [(c1, c1), (c1, c2), (c2, c2), (c2, c3), (c3, c3), (c3, c4), (c4, c4), (c4, c5)]))
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        # one workspace per strided build: its site table is read again by the SubM build of the level it produces
        self.ws_conv = {}
        lvl = 0
        for stem, kind, _ci, _co, ks, _st, _pd, key in BACKBONE8X_LAYERS:
            if kind != "subm" and key not in self.ws_conv:
                out = self.level_of_key[key]
                self.ws_conv[key] = torch.empty((L.pcdb_rulebook_workspace_bytes(self.caps[lvl], ks[0] * ks[1] * ks[2],
                                                                                 self.caps[out]),), dtype=torch.uint8, device=dev)
            lvl = self.level_of_key[key]
        d, h, w = self.shapes[4]
        # zero except for the rows of the last scatter, which dense_rows / dense_count remember: every step undoes
        # the previous scatter (pcdb_dense_clear_rows) instead of clearing all of it
        self.dense = torch.zeros((B, 128, d, h, w), dtype=dt, device=dev)
        self.dense_rows = torch.zeros((self.caps[4], 4), **i32)
        self.dense_count = torch.zeros((2,), **i32)
        # NMS
        nb = cfg.nms_boxes_per_frame
        self.nms_offsets = np.arange(B + 1, dtype=np.int32) * nb
        self.nms_ws = torch.empty((L.pcdb_nms_workspace_bytes(B, nb),), dtype=torch.uint8, device=dev)
        self.keep = torch.empty((B, cfg.nms_keep_per_frame), dtype=torch.int64, device=dev)
        self.num_keep = torch.empty((B,), dtype=torch.int32, device=dev)
        self.side_stream = torch.cuda.Stream(device=dev)
        self.side_stream_b = torch.cuda.Stream(device=dev)
        self.ws_b = torch.empty((max(L.pcdb_rulebook_workspace_bytes(c, 27, c) for c in (c1, c2, c3, c4)),),
                                dtype=torch.uint8, device=dev)
        self.rows_hint = [0] * 5                  # expected rows per level (0 = unknown: the capacity is assumed)
        self._events = {key: torch.cuda.Event() for key in self.nbr}
        self._dense_cleared = torch.cuda.Event()
        self._rb_cleared = torch.cuda.Event()
        self._sites_ready = torch.cuda.Event()
        self._dense_clear_issued = False
        self.side_stream_c = torch.cuda.Stream(device=dev)
        self.conv_stream = torch.cuda.Stream(device=dev)
        self._site_events = {key: torch.cuda.Event() for key in self.nbr}
        self._maps_ready = torch.cuda.Event()
        self._maps_l1_ready = torch.cuda.Event()
        self._fork = torch.cuda.Event()
        self._chain_undone = torch.cuda.Event()
        self._insert_done = torch.cuda.Event()
        if cfg.rulebook_chain:
            self._prepare_chain()

    # ------------------------------------------------------------------------------------------
    def _count_ptr(self, level):
        if level == 0:
            return C.c_void_p(self.voxel_offsets.data_ptr() + 4 * self.cfg.batch_size)
        return ptr(self.counts[level])

    def voxelize(self, points: torch.Tensor, frame_offsets: torch.Tensor, stream, phase: str = "all"):
        """phase "sites": coordinates + counts (what the rulebook builds wait for); "points": voxel contents and
        the mean VFE (what the first convolution waits for); "all": both."""
        cfg, L = self.cfg, self.lib
        n = points.shape[0]
        assert n <= cfg.max_points_total, f"{n} points exceed max_points_total={cfg.max_points_total}"
        fn = {"all": L.pcdb_voxelize, "sites": L.pcdb_voxelize_sites, "points": L.pcdb_voxelize_points}[phase]
        check(fn(ptr(points), n, points.shape[1], ptr(frame_offsets), cfg.batch_size,
                 f32xN(np.asarray(cfg.voxel_size, np.float32)),
                 f32xN(np.asarray(cfg.point_cloud_range, np.float32)), i32x3(self.grid_xyz),
                 cfg.max_num_points, cfg.max_voxels, int(cfg.overflow_break), None,
                 ptr(self.coords[0]), ptr(self.num_points), ptr(self.vfe),
                 BF16 if self.tc else F32, self.cin0, None, ptr(self.voxel_offsets), ptr(self.ws),
                 self.ws.numel(), stream), "pcdb_voxelize")

    def _build_rulebook(self, lyr, level, out_level, stream, ws, site_table=None):
        """site_table = (workspace, n_in_cap, K, n_out_cap) of the strided build that produced `level`: its hash
        table already maps this level's sites to rows, so the SubM build only looks neighbours up."""
        L, B, key = self.lib, self.cfg.batch_size, lyr["key"]
        if lyr["kind"] == "subm" and site_table is not None:
            tws, t_in_cap, t_k, t_out_cap = site_table
            check(L.pcdb_rulebook_subm_reuse(ptr(self.coords[level]), self.caps[level], self._count_ptr(level), B,
                                             i32x3(self.shapes[level]), i32x3(lyr["ks"]), i32x3([1, 1, 1]),
                                             ptr(self.nbr[key]), self.caps[level], ptr(tws), t_in_cap, t_k, t_out_cap,
                                             RB_CLEARED, stream), "pcdb_rulebook_subm_reuse")
        elif lyr["kind"] == "subm":
            check(L.pcdb_rulebook_subm(ptr(self.coords[level]), self.caps[level], self._count_ptr(level), B,
                                       i32x3(self.shapes[level]), i32x3(lyr["ks"]), i32x3([1, 1, 1]),
                                       ptr(self.nbr[key]), self.caps[level], ptr(ws), ws.numel(),
                                       stream), "pcdb_rulebook_subm")
        else:
            raise AssertionError("strided builds go through _build_sites / _build_pairs")

    def _build_sites(self, lyr, level, out_level, stream, ws):
        """First half of a strided build: the output sites (coordinates, count, site table)."""
        L, B = self.lib, self.cfg.batch_size
        check(L.pcdb_rulebook_conv_sites(ptr(self.coords[level]), self.caps[level], self._count_ptr(level), B,
                                         i32x3(self.shapes[level]), i32x3(self.shapes[out_level]),
                                         i32x3(lyr["ks"]), i32x3(lyr["st"]), i32x3(lyr["pd"]), i32x3([1, 1, 1]),
                                         ptr(self.coords[out_level]), self.caps[out_level],
                                         ptr(self.counts[out_level]), ptr(ws), ws.numel(), RB_CLEARED, stream),
              "pcdb_rulebook_conv_sites")

    def _build_pairs(self, lyr, level, out_level, stream, ws):
        """Second half: the neighbour map the strided convolution itself consumes."""
        check(self.lib.pcdb_rulebook_conv_pairs(self.caps[level], self._count_ptr(level), i32x3(lyr["ks"]), i32x3(lyr["st"]),
                                                i32x3([1, 1, 1]), self.caps[out_level], ptr(self.nbr[lyr["key"]]),
                                                self.caps[out_level], None, 0, ptr(ws), RB_CLEARED, stream),
              "pcdb_rulebook_conv_pairs")

    def _prepare_chain(self):
        """Static arguments of pcdb_rulebook_chain: geometry of the strided convolutions, capacities, device pointers."""
        L = self.lib
        strided = [l for l in self.layers if l["kind"] != "subm"]
        subm_of_level = {}
        level = 0
        for lyr in self.layers:
            if lyr["kind"] == "subm":
                subm_of_level.setdefault(level, lyr)
            level = self.level_of_key[lyr["key"]]
        n_levels = len(strided) + 1
        flat = lambda rows: (C.c_int32 * (3 * len(rows)))(*[int(v) for r in rows for v in r])
        ptrs = lambda ts: (C.c_void_p * n_levels)(*[None if t is None else t.data_ptr() for t in ts])
        self._chain = dict(
            n_levels=n_levels, shapes=flat(self.shapes), ksize=flat([l["ks"] for l in strided]), stride=flat([l["st"] for l in strided]),
            padding=flat([l["pd"] for l in strided]), caps=(C.c_int32 * n_levels)(*self.caps),
            coords=ptrs([None] + self.coords[1:]), counts=ptrs([None] + self.counts[1:]),
            nbr_conv=ptrs([None] + [self.nbr[l["key"]] for l in strided]),
            subm_ksize=flat([subm_of_level[lv]["ks"] if lv in subm_of_level else [0, 0, 0] for lv in range(n_levels)]),
            nbr_subm=ptrs([self.nbr[subm_of_level[lv]["key"]] if lv in subm_of_level else None for lv in range(n_levels)]),
            maps=[(self.nbr[l["key"]], self.level_of_key[l["key"]], l["K"]) for l in strided] +
                 [(self.nbr[l["key"]], lv, l["K"]) for lv, l in subm_of_level.items()])
        nbytes = L.pcdb_rulebook_chain_workspace_bytes(self.cfg.batch_size, n_levels, self._chain["shapes"], self._chain["caps"])
        assert nbytes > 0, "rulebook chain: a level exceeds the cell-index limits; use rulebook_chain=False"
        self.ws_chain = torch.empty((nbytes,), dtype=torch.uint8, device=self.dev)
        self._clear_chain_workspace(C.c_void_p(torch.cuda.current_stream().cuda_stream))     # afterwards every build undoes itself

    def _clear_chain_workspace(self, stream):
        ch = self._chain
        check(self.lib.pcdb_rulebook_chain_clear(ptr(self.ws_chain), self.ws_chain.numel(), self.cfg.batch_size, ch["n_levels"],
                                                 ch["shapes"], ch["caps"], None, None, None, None, None, None, stream),
              "pcdb_rulebook_chain_clear")

    def _build_chain(self, stream, phase=7):
        """phase mask: 1 occupancy of every level, 4 level 1's SubM map (what conv_input / conv1 wait for), 2 everything else."""
        ch = self._chain
        hint = (C.c_int32 * ch["n_levels"])(*self.rows_hint) if any(self.rows_hint) else None
        check(self.lib.pcdb_rulebook_chain(ptr(self.coords[0]), self._count_ptr(0), self.cfg.batch_size, ch["n_levels"], ch["shapes"],
                                           ch["ksize"], ch["stride"], ch["padding"], ch["caps"], ch["coords"], ch["counts"],
                                           ch["nbr_conv"], ch["subm_ksize"], ch["nbr_subm"], hint, ptr(self.ws_chain),
                                           self.ws_chain.numel(), RB_CLEARED | RB_UNDONE, phase, stream), "pcdb_rulebook_chain")

    def _clear_rulebook_buffers(self, stream):
        """Everything the strided builds and the table-reusing SubM builds would clear first (they pass
        PCDB_RB_CLEARED): hash tables, owner masks, and the rows of the neighbour maps that the previous step filled."""
        L = self.lib
        if self.cfg.rulebook_chain:
            ch = self._chain
            # only the rows of every map that its previous build wrote (the maps start all -1), all maps in one launch
            check(L.pcdb_rulebook_chain_clear(None, 0, self.cfg.batch_size, ch["n_levels"], ch["shapes"],
                                              ch["caps"], ch["ksize"], ch["subm_ksize"], self._count_ptr(0), ch["counts"],
                                              ch["nbr_conv"], ch["nbr_subm"], stream), "pcdb_rulebook_chain_clear")
            return
        lvl = 0
        done = set()
        for lyr in self.layers:
            key, out = lyr["key"], self.level_of_key[lyr["key"]]
            if key not in done:
                done.add(key)
                if lyr["kind"] != "subm":
                    ws = self.ws_conv[key]
                    check(L.pcdb_rulebook_conv_clear(ptr(ws), ws.numel(), self.caps[lvl], lyr["K"], self.caps[out],
                                                     None, self.caps[out], stream), "pcdb_rulebook_conv_clear")
                if lyr["kind"] != "subm" or lvl > 0:
                    # the neighbour map: only the rows the previous build of this map wrote (its row count is still in
                    # the level's device counter; the maps start all -1), not K x capacity
                    rows = self.level_of_key[key]
                    check(L.pcdb_fill_rows_i32(ptr(self.nbr[key]), self.caps[rows], lyr["K"], self._count_ptr(rows), self.caps[rows],
                                               -1, stream), "pcdb_fill_rows_i32")
            lvl = out

    def clear_dense_async(self, fork=None):
        """Everything of a step that depends on nothing, on its own stream: the zeroing of the dense BEV tensor -- not the
        72 MB of a KITTI batch of 4, only the cells the previous step scattered into (pcdb_dense_clear_rows) -- and, for the
        one-build-per-map rulebooks, their memsets (hash tables, neighbour maps).  The rulebook branches wait for the first
        event, the scatter into the dense tensor for the second.  (The four-launch chain clears its buffers at the END of the
        step instead, next to the NMS sweep: see `step`.)"""
        main = torch.cuda.current_stream()
        if fork is not None:
            self.side_stream_c.wait_event(fork)       # branch off where the step began, not behind what main has queued since
        else:
            self.side_stream_c.wait_stream(main)
        with torch.cuda.stream(self.side_stream_c):
            sc = C.c_void_p(self.side_stream_c.cuda_stream)
            self._clear_rulebook_buffers(sc)
            self._rb_cleared.record(self.side_stream_c)
            if not self.cfg.rulebook_chain:
                self._clear_dense(sc)
        self._dense_clear_issued = True

    def _clear_dense(self, sc):
        check(self.lib.pcdb_dense_clear_rows(ptr(self.dense_rows), self.caps[4], ptr(self.dense_count), 128, self.cfg.batch_size,
                                             i32x3(self.shapes[4]), ptr(self.dense), BF16 if self.tc else F32, sc),
              "pcdb_dense_clear_rows")
        self._dense_cleared.record(self.side_stream_c)

    def backbone(self, stream=None, sites_ready=None):
        """8 rulebook builds + 12 fused conv kernels + dense as three branches of the captured graph.

        Rulebooks depend on voxel COORDINATES only, convolutions on features.  The site numbering of the four
        strided convolutions is a chain (each produces the next level's coordinates) on side stream A; the
        neighbour maps -- pairs of each strided conv, SubM map of each level (looked up in the strided build's
        site table) -- only need their level's sites and run on side stream B in the order the convolutions
        consume them; the convolution chain on the main stream waits for the event of the map it consumes."""
        L, B = self.lib, self.cfg.batch_size
        main = torch.cuda.current_stream()
        side_a, side_b = self.side_stream, self.side_stream_b
        if self.cfg.rulebook_chain:
            return self._backbone_chain(sites_ready)
        if sites_ready is not None:              # the rulebook branches only need the voxel coordinates
            side_a.wait_event(sites_ready)
            side_b.wait_event(sites_ready)
        else:
            side_a.wait_stream(main)
            side_b.wait_stream(main)
        if not self._dense_clear_issued:
            self.clear_dense_async()
        self._dense_clear_issued = False
        side_a.wait_event(self._rb_cleared)
        side_b.wait_event(self._rb_cleared)
        events = {}
        site_tables = {}         # level -> workspace of the strided build whose outputs are that level's sites
        site_tables_of_key = {}
        level_of_layer = []
        level = 0
        for lyr in self.layers:
            level_of_layer.append(level)
            level = self.level_of_key[lyr["key"]]
        sites_done = {}          # level -> event: its coordinates, count and site table exist
        with torch.cuda.stream(side_a):                      # sites of spconv2 -> spconv3 -> spconv4 -> spconv_down2
            sa = C.c_void_p(side_a.cuda_stream)
            for lyr, lvl in zip(self.layers, level_of_layer):
                key = lyr["key"]
                if lyr["kind"] != "subm" and key not in site_tables_of_key:
                    out = self.level_of_key[key]
                    self._build_sites(lyr, lvl, out, sa, self.ws_conv[key])
                    self._site_events[key].record(side_a)
                    sites_done[out] = self._site_events[key]
                    site_tables[out] = (self.ws_conv[key], self.caps[lvl], lyr["K"], self.caps[out])
                    site_tables_of_key[key] = True
        with torch.cuda.stream(side_b):                      # subm1, then per level: pairs of its strided conv, its SubM
            sb = C.c_void_p(side_b.cuda_stream)
            for lyr, lvl in zip(self.layers, level_of_layer):
                key = lyr["key"]
                if key in events:
                    continue
                if lyr["kind"] == "subm":
                    if lvl in sites_done:
                        side_b.wait_event(sites_done[lvl])               # this level's coordinates exist
                    self._build_rulebook(lyr, lvl, lvl, sb, self.ws_b, site_tables.get(lvl))
                else:
                    out = self.level_of_key[key]
                    side_b.wait_event(sites_done[out])
                    self._build_pairs(lyr, lvl, out, sb, self.ws_conv[key])
                self._events[key].record(side_b)
                events[key] = self._events[key]
        self._convs_and_dense(main, events)
        main.wait_stream(side_a)
        main.wait_stream(side_b)

    def _backbone_chain(self, sites_ready):
        """Two branches: (A) all rulebooks in four launches as soon as the voxel coordinates exist, (conv) the 12
        convolutions + dense once the maps and the VFE features are there."""
        main = torch.cuda.current_stream()
        side_a = self.side_stream
        if sites_ready is not None:
            side_a.wait_event(sites_ready)
        else:
            side_a.wait_stream(main)
        if not self._dense_clear_issued:
            self.clear_dense_async()
        self._dense_clear_issued = False
        first_key = self.layers[0]["key"]
        side_b = self.side_stream_b
        with torch.cuda.stream(side_a):
            self._build_chain(C.c_void_p(side_a.cuda_stream), 1)
            self._insert_done.record(side_a)
        # level 1's SubM map needs the level-0 table only: on its own stream, beside the numbering of the other levels
        side_b.wait_event(self._insert_done)
        side_b.wait_event(self._rb_cleared)                  # the maps' extents are -1 again
        with torch.cuda.stream(side_b):
            self._build_chain(C.c_void_p(side_b.cuda_stream), 4)
            self._events[first_key].record(side_b)           # the first level's convolutions can start
        # The undo of the previous step's BEV scatter is only needed by to_dense at the very end; issued at the start of
        # the step (round 1) its 1.2 M scattered stores ran beside the voxel gather and the first rulebook kernels -- the
        # serial stretch every convolution waits for (kernel timeline: rbc_count and the level-1 map started 11 us after
        # rbc_insert had ended).  Behind the level-1 map it overlaps the first convolutions instead.
        sc = self.side_stream_c
        sc.wait_event(self._events[first_key])
        with torch.cuda.stream(sc):
            self._clear_dense(C.c_void_p(sc.cuda_stream))
        side_a.wait_event(self._rb_cleared)
        level1_keys = {l["key"] for l in self.layers if self.level_of_key[l["key"]] == 1}
        with torch.cuda.stream(side_a):
            # numbering of every level + the maps of level 1 first: conv2.x can start while the deeper maps are built
            # (kernel timeline: they used to wait 30 us for ONE launch that built all seven remaining maps)
            self._build_chain(C.c_void_p(side_a.cuda_stream), 2 | 32)
            self._maps_l1_ready.record(side_a)
            self._build_chain(C.c_void_p(side_a.cuda_stream), 16)
            self._maps_ready.record(side_a)
            self._build_chain(C.c_void_p(side_a.cuda_stream), 8)      # leaves the workspace clean for the next step
            self._chain_undone.record(side_a)
        self._convs_and_dense(main, {key: (self._events[first_key] if key == first_key else
                                           (self._maps_l1_ready if key in level1_keys else self._maps_ready)) for key in self.nbr})
        main.wait_event(self._chain_undone)

    def _convs_and_dense(self, main, events):
        """The convolution chain on its own stream (stream priorities were measured: no effect in either direction); it
        waits for the event of every neighbour map before the first layer that consumes it."""
        L, B = self.lib, self.cfg.batch_size
        conv = self.conv_stream
        conv.wait_stream(main)
        stream = C.c_void_p(conv.cuda_stream)
        waited = set()
        waited_events = set()
        level = 0
        x = self.vfe
        flip = 0
        for lyr in self.layers:
            key = lyr["key"]
            out_level = self.level_of_key[key]
            if key not in waited:
                if id(events[key]) not in waited_events:
                    conv.wait_event(events[key])
                    waited_events.add(id(events[key]))
                waited.add(key)
            flip ^= 1
            out = self.feat[out_level][flip]
            # the buffer is wider than some layers need: address it as a dense (cap, c_out) matrix
            out_view = out.view(-1)[: self.caps[out_level] * lyr["c_out"]].view(self.caps[out_level], lyr["c_out"])
            # PDL (set-up overlaps the previous layer's tail): the previous kernel of this stream only produces
            # this layer's input features; the first layer is excluded, its row count comes from the voxelizer
            check(L.pcdb_sparse_conv_fwd(ptr(x), x.shape[0], ptr(lyr["w"]), ptr(self.nbr[key]), self.caps[out_level], lyr["K"],
                                         self.caps[out_level], self._count_ptr(out_level), lyr["c_in"], lyr["c_out"],
                                         BF16 if self.tc else F32, ptr(lyr["scale"]), ptr(lyr["shift"]), None,
                                         EPI_RELU | lyr["wflags"] | (CONV_PDL if (self.tc and lyr is not self.layers[0]) else 0)
                                         | (CONV_SHALLOW_RING if self.cfg.conv_shallow_ring else 0),
                                         ptr(out_view), self.cfg.conv_algo | (self.rows_hint[out_level] << 8), stream),
                  "pcdb_sparse_conv_fwd")
            x = out_view
            level = out_level
        self.last_features = x
        conv.wait_event(self._dense_cleared)
        check(L.pcdb_to_dense(ptr(x), ptr(self.coords[4]), self.caps[4], self._count_ptr(4), 128,
                              BF16 if self.tc else F32, B, i32x3(self.shapes[4]), ptr(self.dense),
                              (BF16 if self.tc else F32) | 0x100, stream), "pcdb_to_dense")
        # remember what was scattered: the next step clears exactly these rows (the conv stream has the level-4 sites:
        # pcdb_to_dense just read them)
        with torch.cuda.stream(conv):
            self.dense_rows.copy_(self.coords[4], non_blocking=True)
            self.dense_count.copy_(self.counts[4], non_blocking=True)
        main.wait_stream(conv)

    def nms(self, boxes_bev_sorted: torch.Tensor, stream):
        """boxes (B * nms_boxes_per_frame, 5) f32, each frame's block sorted by descending score."""
        cfg = self.cfg
        check(self.lib.pcdb_nms(ptr(boxes_bev_sorted), ptr(self.nms_offsets), cfg.batch_size, cfg.nms_thresh, 0,
                                ptr(self.keep), cfg.nms_keep_per_frame, ptr(self.num_keep), ptr(self.nms_ws),
                                self.nms_ws.numel(), stream), "pcdb_nms")

    def step(self, points: torch.Tensor, frame_offsets: torch.Tensor, boxes_bev_sorted: torch.Tensor):
        """One pass of the hot path over one batch.  Returns device tensors; no host sync."""
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        # The voxelizer is issued BEFORE the side-stream clears: a replayed graph starts its root branches in issue order,
        # and with the clears first the voxel hash waited 35 us behind twelve small memsets / fills (kernel timeline).
        fork = self._fork
        fork.record(torch.cuda.current_stream())
        self.voxelize(points, frame_offsets, stream, "sites")
        self._sites_ready.record(torch.cuda.current_stream())
        self.clear_dense_async(fork)      # rulebook buffers + dense tensor, concurrent with the voxelizer
        self.voxelize(points, frame_offsets, stream, "points")     # overlaps the first rulebook builds
        self.backbone(sites_ready=self._sites_ready)
        self.nms(boxes_bev_sorted, stream)
        d = self.dense
        # level_counts: [count, overflow flag] of the active sites of levels 1-4 (row 0 unused).  A set flag means the level's
        # row capacity (HotPathConfig.level_capacity) was too small for this batch: sites beyond it were dropped and the BEV
        # map is incomplete.  HostRunner checks the flags with every result; device-side callers must do so themselves.
        return dict(spatial_features=d.view(d.shape[0], d.shape[1] * d.shape[2], d.shape[3], d.shape[4]),
                    keep=self.keep, num_keep=self.num_keep, voxel_offsets=self.voxel_offsets, level_counts=self.counts_all)

    # ------------------------------------------------------------------------------------------
    def capture(self, points_buf: torch.Tensor, offsets_buf: torch.Tensor, boxes_buf: torch.Tensor):
        """Captures one `step` over STATIC input buffers into a CUDA graph.  points_buf has
        max_points_total rows; the real point count is offsets_buf[B] on the device, so the same graph
        serves every batch written into the buffers.  Returns (graph, outputs)."""
        assert points_buf.shape[0] == self.cfg.max_points_total
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.step(points_buf, offsets_buf, boxes_buf)       # warm-up outside capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # The buffers are capacities; how many rows each level really holds is only known on the device.  The
        # warm-up step's counts tell the conv launcher what to expect (ring depth / CTAs per SM -- tuning only).
        self.rows_hint = [int(self.voxel_offsets[-1])] + [int(c[0]) for c in self.counts[1:]]
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.step(points_buf, offsets_buf, boxes_buf)
        return graph, out

    def make_host_runner(self):
        """End-to-end entry point: host frames in, kept-box indices out (see HostRunner)."""
        return HostRunner(self)

    def level_counts(self) -> List[int]:
        """Host copy of the active-site counts (synchronises; for tests / reporting only)."""
        res = [int(self.voxel_offsets[self.cfg.batch_size].item())]
        for lv in range(1, 5):
            cnt, overflow = self.counts[lv].tolist()
            assert overflow == 0, f"level {lv} capacity {self.caps[lv]} exceeded"
            res.append(cnt)
        return res

    def launches_per_step(self) -> int:
        """KERNELS of libpcdet_b200.so launched by one `step` (memset nodes and torch fills not counted)."""
        vox = 6                                  # sites: insert, count, rank, write_coords; points: assign, gather
        convs = 12
        dense = 2                                # undo of the previous scatter, scatter
        nms = 5                                  # prepare, mask (candidates), resolve, diag, sweep
        if self.cfg.rulebook_chain:
            rulebooks = 7                        # rbc_insert, rbc_maps (level 0), rbc_count, rbc_assign, rbc_maps x 2, rbc_undo
            clears = 1                           # rbc_fill_maps: the extents of the 4 strided maps and of level 1's SubM map
        else:
            rulebooks = (2 + 3) + 4 * 4          # SubM: insert + neighbours, 3 x neighbours; strided: insert, mark, number | fill
            clears = 4 + 3                       # pcdb_fill_rows_i32: the maps of the 4 strided convs and of the SubM levels 2-4
        return vox + rulebooks + convs + dense + clears + nms


class HostRunner:
    """The call a user makes with HOST data.

        keep, num = runner(frames, boxes_bev)            # synchronous
        t = runner.submit(frames, boxes_bev); ...; keep, num = runner.result(t)     # pipelined, `depth` in flight

    frames: list of B numpy (N_b, C) float32 point clouds; boxes_bev: (B*nms_boxes_per_frame, 5) float32
    score-sorted BEV boxes (the head's output in a full detector).  Per call the frames are packed into a
    pinned staging buffer, copied host->device on a copy stream, the captured graph of the hot path is
    replayed on the compute stream, and the keep lists are copied device->host.  With `submit`/`result` the
    host-side packing and the H2D copy of batch i+1 overlap the GPU work of batch i (two input slots; the
    hot path's internal buffers are reused because the replays are serialised on one stream).
    Results are numpy views of pinned memory, valid until the slot is reused `depth` submits later."""

    def __init__(self, hp, depth: int = 2):
        """hp: one SecondHotPath (the slots share its internal buffers, so their replays are serialised on one
        stream), or a list of `depth` independent SecondHotPath instances: every slot then owns its buffers and its
        compute stream, and the GPU overlaps the latency-bound phases of one batch (voxel hash, first rulebooks,
        NMS sweep) with the convolutions of the other."""
        hps = list(hp) if isinstance(hp, (list, tuple)) else [hp] * depth
        assert len(hps) == depth
        self.hp = hp = hps[0]
        cfg = hp.cfg
        dev = hp.dev
        C_ = cfg.num_point_features
        self.depth = depth
        own_stream = len({id(h) for h in hps}) == depth and depth > 1
        self.compute = torch.cuda.current_stream(dev)
        self.copy = torch.cuda.Stream(device=dev)
        self.slots = []
        nb = cfg.batch_size * cfg.nms_boxes_per_frame
        for i in range(depth):
            hp = hps[i]
            sl = dict(
                compute=torch.cuda.Stream(device=dev) if own_stream else self.compute,
                points_dev=torch.zeros((cfg.max_points_total, C_), dtype=torch.float32, device=dev),
                offsets_dev=torch.zeros((cfg.batch_size + 1,), dtype=torch.int32, device=dev),
                boxes_dev=torch.zeros((nb, 5), dtype=torch.float32, device=dev),
                points_pin=torch.zeros((cfg.max_points_total, C_), dtype=torch.float32).pin_memory(),
                offsets_pin=torch.zeros((cfg.batch_size + 1,), dtype=torch.int32).pin_memory(),
                boxes_pin=torch.zeros((nb, 5), dtype=torch.float32).pin_memory(),
                keep_pin=torch.zeros((cfg.batch_size, cfg.nms_keep_per_frame), dtype=torch.int64).pin_memory(),
                num_pin=torch.zeros((cfg.batch_size,), dtype=torch.int32).pin_memory(),
                counts_pin=torch.zeros((5, 2), dtype=torch.int32).pin_memory(),
                h2d_done=torch.cuda.Event(), done=torch.cuda.Event(), busy=False)
            sl["graph"], sl["out"] = hp.capture(sl["points_dev"], sl["offsets_dev"], sl["boxes_dev"])
            self.slots.append(sl)
        self.ticket = 0
        self.h2d_bytes = 0
        self.d2h_bytes = cfg.batch_size * cfg.nms_keep_per_frame * 8 + cfg.batch_size * 4 + 5 * 2 * 4

    def submit(self, frames, boxes_bev) -> int:
        """frames: B point clouds (N_b, C) float32 -- numpy arrays, or torch CPU tensors.  PINNED torch tensors (``t.pin_memory()``)
        go to the device slot directly, one cudaMemcpyAsync per frame and no staging copy on the host; anything else is
        first packed into the slot's pinned staging buffer (a 20 MB memcpy per nuScenes batch, which was what bounded the
        end-to-end rate of that workload).  boxes_bev likewise."""
        cfg = self.hp.cfg
        assert len(frames) == cfg.batch_size
        t = self.ticket
        sl = self.slots[t % self.depth]
        if sl["busy"]:
            sl["done"].synchronize()          # the slot's previous result must have left the GPU
        offs = sl["offsets_pin"].numpy()
        offs[0] = 0
        sizes = [int(f.shape[0]) for f in frames]
        for b, n in enumerate(sizes):
            offs[b + 1] = offs[b] + n
        total = int(offs[cfg.batch_size])
        assert total <= cfg.max_points_total, "batch exceeds max_points_total"
        direct = all(isinstance(f, torch.Tensor) and f.is_pinned() and f.dtype == torch.float32 and f.is_contiguous() for f in frames)
        if not direct:
            pin = sl["points_pin"].numpy()
            for b, f in enumerate(frames):
                pin[offs[b]:offs[b + 1]] = f.numpy() if isinstance(f, torch.Tensor) else f
        boxes_direct = isinstance(boxes_bev, torch.Tensor) and boxes_bev.is_pinned() and boxes_bev.dtype == torch.float32
        if not boxes_direct:
            sl["boxes_pin"].numpy()[...] = boxes_bev.numpy() if isinstance(boxes_bev, torch.Tensor) else boxes_bev
        with torch.cuda.stream(self.copy):
            if direct:
                for b, f in enumerate(frames):
                    sl["points_dev"][offs[b]:offs[b + 1]].copy_(f, non_blocking=True)
            else:
                sl["points_dev"][:total].copy_(sl["points_pin"][:total], non_blocking=True)
            sl["offsets_dev"].copy_(sl["offsets_pin"], non_blocking=True)
            sl["boxes_dev"].copy_(boxes_bev.view_as(sl["boxes_dev"]) if boxes_direct else sl["boxes_pin"], non_blocking=True)
            sl["h2d_done"].record(self.copy)
        self.h2d_bytes = total * sl["points_dev"].shape[1] * 4 + offs.nbytes + sl["boxes_pin"].numel() * 4
        with torch.cuda.stream(sl["compute"]):
            sl["compute"].wait_event(sl["h2d_done"])
            sl["graph"].replay()
            sl["keep_pin"].copy_(sl["out"]["keep"], non_blocking=True)
            sl["num_pin"].copy_(sl["out"]["num_keep"], non_blocking=True)
            sl["counts_pin"].copy_(sl["out"]["level_counts"], non_blocking=True)
            sl["done"].record(sl["compute"])
        sl["busy"] = True
        self.ticket += 1
        return t

    def result(self, ticket: int):
        sl = self.slots[ticket % self.depth]
        sl["done"].synchronize()
        counts = sl["counts_pin"].numpy()
        if counts[1:, 1].any():
            from ._lib import PcdbError
            lv = [int(i) for i in np.nonzero(counts[:, 1])[0]]
            raise PcdbError(f"active-site capacity exceeded at level(s) {lv} (capacities {self.hp.caps}): sites were dropped; "
                            "raise HotPathConfig.level_capacity / max_points_total")
        return sl["keep_pin"].numpy(), sl["num_pin"].numpy()

    def __call__(self, frames, boxes_bev):
        return self.result(self.submit(frames, boxes_bev))
