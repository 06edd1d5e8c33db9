"""Spatial tiling with per-frame NEIGHBOUR halo exchange for ONE long clip (BASELINE.json configs[3]).

The recurrence cannot be parallelised in time, so a single high-resolution clip is split in SPACE: the LR frame is cut
into a gy x gx grid; every tile is processed together with a halo of `halo` LR pixels (the "extended tile") exactly
like an independent smaller image, and after every frame the recurrent state inside the halo is refreshed from the
neighbours' interiors.  Per frame the state update has a receptive field of about 17 LR px + max|flow| (SURVEY.md
8(e)), so with halo >= 18 + max|flow| the interior of every tile equals the untiled result up to fp32 rounding of
the sampling coordinates (tile-local instead of frame-global integers); `halo="auto"` sizes it from the measured
max|flow| of the clip.  The LR-only work (FNet, encoder_lr) has a receptive field > 100 px and is cheap (~7 % of the
MACs): it is computed on the FULL frame by every rank and cropped (replicated, no exchange).

One process per GPU: rank r owns tiles r, r+world, ...  The ONLY communication is the exchange step between frames:
for every ordered pair (source tile k, destination tile j) whose regions overlap — interior(k) x extended(j), i.e. the
8 neighbours on a regular grid — the owner of k sends exactly that strip of its state (HR 4 ch + L1 24 ch, packed
into one buffer per pair) to the owner of j, all pairs of a frame in ONE `batch_isend_irecv` (a single
ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd over NVLink); pairs inside one process are plain device copies.
The state never leaves the extended-tile buffers, the clip inputs are cropped once per clip, and the result stays
sharded (`gather_output=False`) or is gathered once at the end of the clip.  With a single process all tiles run back
to back on one GPU — this is how the tests check the tiled result against the untiled forward.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.distributed as dist

from . import _lib as L

STATE_HALO_BASE = 18   # receptive field of one frame's state update in LR pixels, without the flow


def tile_plan(h: int, w: int, gy: int, gx: int, halo: int):
    """[(interior (y0,y1,x0,x1), extended (ey0,ey1,ex0,ex1))] in LR pixels; interiors partition the frame."""
    if gy < 1 or gx < 1 or halo < 0 or gy > h or gx > w:
        raise ValueError("bad tiling")
    ys = [round(i * h / gy) for i in range(gy + 1)]
    xs = [round(j * w / gx) for j in range(gx + 1)]
    plan = []
    for i in range(gy):
        for j in range(gx):
            y0, y1, x0, x1 = ys[i], ys[i + 1], xs[j], xs[j + 1]
            plan.append(((y0, y1, x0, x1), (max(0, y0 - halo), min(h, y1 + halo), max(0, x0 - halo), min(w, x1 + halo))))
    return plan


def halo_pairs(plan):
    """[(src tile k, dst tile j, (y0, y1, x0, x1))]: the strip of k's INTERIOR that lies inside j's EXTENDED region (k != j),
    in LR pixels and in a fixed order every rank reproduces."""
    pairs = []
    for k, ((y0, y1, x0, x1), _) in enumerate(plan):
        for j, (_, (ey0, ey1, ex0, ex1)) in enumerate(plan):
            if j == k:
                continue
            a0, a1, b0, b1 = max(y0, ey0), min(y1, ey1), max(x0, ex0), min(x1, ex1)
            if a0 < a1 and b0 < b1:
                pairs.append((k, j, (a0, a1, b0, b1)))
    return pairs


def _strip_views(state_hr, state_l1, ext, region):
    """Views of `region` (LR pixels, frame coordinates) inside a tile's extended-state buffers."""
    ey0, _, ex0, _ = ext
    a0, a1, b0, b1 = region
    return (state_hr[:, 8 * (a0 - ey0):8 * (a1 - ey0), 8 * (b0 - ex0):8 * (b1 - ex0)],
            state_l1[:, 2 * (a0 - ey0):2 * (a1 - ey0), 2 * (b0 - ex0):2 * (b1 - ex0)])


def exchange_halos(states, plan, pairs, world, rank, group=None):
    """The one exchange step of a frame.  `states[k] = (state_hr (n,8eh,8ew,4), state_l1 (n,2eh,2ew,24))` for the tiles
    this rank owns (k % world == rank).  Every strip of `pairs` moves from its source tile's buffers into its destination
    tile's buffers: a device copy when both tiles live here, otherwise one packed send / recv per pair, all of them in
    one batch.  Returns the number of bytes this rank received over the wire."""
    ops, unpack, keep = [], [], []
    got = 0
    for (k, j, region) in pairs:
        src_here, dst_here = k % world == rank, j % world == rank
        if not (src_here or dst_here):
            continue
        if src_here:
            shr, sl1 = _strip_views(*states[k], plan[k][1], region)
        if dst_here:
            dhr, dl1 = _strip_views(*states[j], plan[j][1], region)
        if src_here and dst_here:
            dhr.copy_(shr)
            dl1.copy_(sl1)
            continue
        if src_here:
            buf = torch.cat([shr.reshape(-1), sl1.reshape(-1)])
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, _global_rank(j % world, group), group))
        else:
            buf = torch.empty(dhr.numel() + dl1.numel(), device=dhr.device, dtype=dhr.dtype)
            ops.append(dist.P2POp(dist.irecv, buf, _global_rank(k % world, group), group))
            unpack.append((buf, dhr, dl1))
            got += buf.numel() * buf.element_size()
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, dhr, dl1 in unpack:
        dhr.copy_(buf[:dhr.numel()].view(dhr.shape))
        dl1.copy_(buf[dhr.numel():].view(dl1.shape))
    return got


def _global_rank(group_rank, group):
    return group_rank if group is None else dist.get_global_rank(group, group_rank)


class TiledClipRunner:
    """Runs `model` (a crfp_b200 CRFP_DSV-family module) on one clip as a grid of halo-extended tiles."""

    def __init__(self, model, grid=(2, 4), halo=32, group=None, gather_output=True, distributed=True):
        """`halo`: LR pixels, or "auto" = 18 + ceil(max|flow|) measured on the clip (one device->host scalar per clip).
        `gather_output=False` leaves every output frame sharded: rank r's result holds the interiors of its own
        tiles and zeros elsewhere (a display pipeline scans the tiles out from their GPUs)."""
        self.model, self.grid, self.halo, self.group, self.gather_output = model, grid, halo, group, gather_output
        self.distributed = distributed   # False: run every tile in this process even under torch.distributed (reference run)
        self.last = {}   # per-clip facts of the last forward: halo used, max|flow|, halo bytes received per frame

    def _world(self):
        if self.distributed and dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    @torch.no_grad()
    def forward(self, lrs, fvs, mks):
        m = self.model
        lrs, fvs, mks = m._check_inputs(lrs, fvs, mks)
        n, t, _, h, w = lrs.shape
        dev = lrs.device
        Cc = m.mid_channels
        lib = L.lib()
        world, rank = self._world()
        f32 = dict(device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            W = m._weights(dev)
            # ---- LR-only stage on the full frame (replicated on every rank: FNet's receptive field is > 100 px)
            shp = L.DsvShape(n=n, t=t, h=h, w=w, mid_channels=Cc)
            pws = torch.empty(lib.crfp_dsv_prepare_workspace(C.byref(shp)), device=dev, dtype=torch.uint8)
            lr4 = torch.empty(n, t, h, w, 4, **f32)
            x_lr = torch.empty(n, t, h, w, Cc, **f32)
            flows = torch.zeros(n, t, h, w, 2, **f32)
            L.check(lib.crfp_dsv_prepare(C.byref(shp), C.byref(W), lrs.data_ptr(), None, lr4.data_ptr(), x_lr.data_ptr(),
                                         flows.data_ptr(), pws.data_ptr(), pws.numel(), st), "dsv_prepare")
            del pws
            max_flow = None
            halo = self.halo
            if halo == "auto" or halo is None:
                max_flow = float(flows.abs().max()) if t > 1 else 0.0       # identical on every rank (replicated FNet)
                halo = STATE_HALO_BASE + int(math.ceil(max_flow))
            plan = tile_plan(h, w, self.grid[0], self.grid[1], halo)
            pairs = halo_pairs(plan)
            mine = [k for k in range(len(plan)) if k % world == rank]
            # ---- per-tile buffers; the clip inputs are cropped ONCE per clip (contiguous per tile, all frames)
            tiles = {}
            fws = 0
            for k in mine:
                (_, (ey0, ey1, ex0, ex1)) = plan[k]
                eh, ew = ey1 - ey0, ex1 - ex0
                ts = L.DsvShape(n=n, t=1, h=eh, w=ew, mid_channels=Cc)
                need = lib.crfp_dsv_frame_workspace(C.byref(ts))
                if need == 0:
                    raise L.CrfpError(f"tile {k} ({eh}x{ew}) is too small")
                fws = max(fws, need)
                tiles[k] = dict(shape=ts, eh=eh, ew=ew,
                                state_hr=torch.zeros(n, 8 * eh, 8 * ew, 4, **f32),
                                state_l1=torch.zeros(n, 2 * eh, 2 * ew, 24, **f32),
                                out=torch.empty(n, 3, 8 * eh, 8 * ew, **f32),
                                lr4=lr4[:, :, ey0:ey1, ex0:ex1].contiguous(),
                                x_lr=x_lr[:, :, ey0:ey1, ex0:ex1].contiguous(),
                                flow=flows[:, :, ey0:ey1, ex0:ex1].contiguous(),
                                fvs=fvs[:, :, :, 8 * ey0:8 * ey1, 8 * ex0:8 * ex1].contiguous(),
                                mks=mks[:, :, :, 8 * ey0:8 * ey1, 8 * ex0:8 * ex1].contiguous())
            del lr4, x_lr, flows
            ws = torch.empty(max(fws, 1), device=dev, dtype=torch.uint8)
            out = torch.zeros(n, t, 3, 8 * h, 8 * w, **f32)
            d = L.DsvFrameDesc()
            d.skip_outside_fovea = int(m.skip_outside_fovea)
            halo_bytes = 0
            for i in range(t):
                for k in mine:
                    ((y0, y1, x0, x1), (ey0, ey1, ex0, ex1)) = plan[k]
                    T = tiles[k]
                    eh, ew = T["eh"], T["ew"]
                    d.shape, d.first = T["shape"], int(i == 0)
                    d.lr4, d.lr4_clip_stride = T["lr4"][:, i].data_ptr(), t * eh * ew * 4
                    d.x_lr, d.x_lr_clip_stride = T["x_lr"][:, i].data_ptr(), t * eh * ew * Cc
                    d.flow, d.flow_clip_stride = T["flow"][:, i].data_ptr(), t * eh * ew * 2
                    d.fvs, d.fvs_clip_stride = T["fvs"][:, i].data_ptr(), t * 3 * 64 * eh * ew
                    d.mks, d.mks_clip_stride = T["mks"][:, i].data_ptr(), t * 64 * eh * ew
                    d.state_hr, d.state_l1 = T["state_hr"].data_ptr(), T["state_l1"].data_ptr()
                    d.out, d.out_clip_stride = T["out"].data_ptr(), 3 * 64 * eh * ew
                    L.check(lib.crfp_dsv_frame(C.byref(d), C.byref(W), ws.data_ptr(), ws.numel(), st), f"tile {k} frame {i}")
                    iy, ix = y0 - ey0, x0 - ex0      # interior of this tile -> this rank's slab of the output
                    out[:, i, :, 8 * y0:8 * y1, 8 * x0:8 * x1] = T["out"][:, :, 8 * iy:8 * (iy + y1 - y0), 8 * ix:8 * (ix + x1 - x0)]
                if i + 1 < t:   # the one exchange step of the frame: neighbour strips only
                    states = {k: (tiles[k]["state_hr"], tiles[k]["state_l1"]) for k in mine}
                    halo_bytes = exchange_halos(states, plan, pairs, world, rank, self.group)
            self.last = {"halo": halo, "max_flow": max_flow, "halo_bytes_received_per_frame": halo_bytes,
                         "tiles_per_rank": len(mine), "pairs": len(pairs)}
            if world > 1 and self.gather_output:
                out = self._gather_output(out, plan, world, rank)
        return out

    def _gather_output(self, out, plan, world, rank):
        """Once per clip: every rank's interiors to every rank (packed, equal-size slots, ONE all-gather)."""
        n, t = out.shape[:2]
        nel = lambda y0, y1, x0, x1: n * t * 3 * 64 * (y1 - y0) * (x1 - x0)
        slot = max(nel(*it) for it, _ in plan)
        per_rank = (len(plan) + world - 1) // world
        send = torch.zeros(per_rank * slot, device=out.device, dtype=out.dtype)
        for k, ((y0, y1, x0, x1), _) in enumerate(plan):
            if k % world == rank:
                o = (k // world) * slot
                send[o:o + nel(y0, y1, x0, x1)].view(n, t, 3, 8 * (y1 - y0), 8 * (x1 - x0)).copy_(out[:, :, :, 8 * y0:8 * y1, 8 * x0:8 * x1])
        recv = torch.empty(world * per_rank * slot, device=out.device, dtype=out.dtype)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        for k, ((y0, y1, x0, x1), _) in enumerate(plan):
            owner = k % world
            if owner != rank:
                o = (owner * per_rank + k // world) * slot
                out[:, :, :, 8 * y0:8 * y1, 8 * x0:8 * x1] = recv[o:o + nel(y0, y1, x0, x1)].view(n, t, 3, 8 * (y1 - y0), 8 * (x1 - x0))
        return out

    __call__ = forward
