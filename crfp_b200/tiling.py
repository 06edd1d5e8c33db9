"""Spatial tiling with per-frame halo refresh for ONE long clip (BASELINE.json configs[3]).

The recurrence cannot be parallelised in time, so a single high-resolution clip is split in SPACE: the LR frame is cut
into a gy x gx grid; every tile is processed together with a halo of `halo` LR pixels (the "extended tile") exactly
like an independent smaller image, and after every frame the recurrent state inside the halo is refreshed from the
neighbours' interiors.  Per frame the state update has a receptive field of about 17 LR px + max|flow| (SURVEY.md
8(e)), so with halo >= 18 + max|flow| the interior of every tile equals the untiled result up to fp32 rounding of
the sampling coordinates (tile-local instead of frame-global integers).  The LR-only work (FNet, encoder_lr) has a
receptive field > 100 px and is cheap (~7 % of the MACs): it is computed on the FULL frame by every rank and cropped.

One process per GPU: rank r owns tiles r, r+world, ...; the only communication is one all-gather of the tiles' interior
state per frame (`torch.distributed`, NCCL over NVLink on GPUs; no collective inside a frame).  With a single process
(`world_size == 1`) all tiles run back to back on one GPU — this is how the tests check the tiled result against the
untiled forward without needing several GPUs.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L


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


class TiledClipRunner:
    """Runs `model` (a crfp_b200 CRFP_DSV-family module) on one clip as a grid of halo-extended tiles."""

    def __init__(self, model, grid=(2, 4), halo=32, group=None, gather_output=True):
        """`gather_output=False` leaves every output frame sharded: rank r's result holds the interiors of its own
        tiles and zeros elsewhere (a display pipeline scans the tiles out from their GPUs)."""
        self.model, self.grid, self.halo, self.group, self.gather_output = model, grid, halo, group, gather_output

    def _world(self):
        if dist.is_available() and dist.is_initialized():
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
        plan = tile_plan(h, w, self.grid[0], self.grid[1], self.halo)
        mine = [k for k in range(len(plan)) if k % world == rank]
        f32 = dict(device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            W = m._weights(dev)
            # ---- LR-only stage on the full frame (replicated on every rank)
            shp = L.DsvShape(n=n, t=t, h=h, w=w, mid_channels=Cc)
            pws = torch.empty(lib.crfp_dsv_prepare_workspace(C.byref(shp)), device=dev, dtype=torch.uint8)
            lr4 = torch.empty(n, t, h, w, 4, **f32)
            x_lr = torch.empty(n, t, h, w, Cc, **f32)
            flows = torch.zeros(n, t, h, w, 2, **f32)
            L.check(lib.crfp_dsv_prepare(C.byref(shp), C.byref(W), lrs.data_ptr(), None, lr4.data_ptr(), x_lr.data_ptr(),
                                         flows.data_ptr(), pws.data_ptr(), pws.numel(), st), "dsv_prepare")
            del pws
            # ---- per-tile buffers
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
                tiles[k] = dict(shape=ts, state_hr=torch.zeros(n, 8 * eh, 8 * ew, 4, **f32),
                                state_l1=torch.zeros(n, 2 * eh, 2 * ew, 24, **f32),
                                out=torch.empty(n, 3, 8 * eh, 8 * ew, **f32))
            ws = torch.empty(max(fws, 1), device=dev, dtype=torch.uint8)
            out = torch.zeros(n, t, 3, 8 * h, 8 * w, **f32)
            full_hr = torch.zeros(n, 8 * h, 8 * w, 4, **f32)
            full_l1 = torch.zeros(n, 2 * h, 2 * w, 24, **f32)
            d = L.DsvFrameDesc()
            d.skip_outside_fovea = int(m.skip_outside_fovea)
            for i in range(t):
                for k in mine:
                    ((y0, y1, x0, x1), (ey0, ey1, ex0, ex1)) = plan[k]
                    T = tiles[k]
                    eh, ew = ey1 - ey0, ex1 - ex0
                    lr4_t = lr4[:, i, ey0:ey1, ex0:ex1].contiguous()
                    xlr_t = x_lr[:, i, ey0:ey1, ex0:ex1].contiguous()
                    fl_t = flows[:, i, ey0:ey1, ex0:ex1].contiguous()
                    fv_t = fvs[:, i, :, 8 * ey0:8 * ey1, 8 * ex0:8 * ex1].contiguous()
                    mk_t = mks[:, i, :, 8 * ey0:8 * ey1, 8 * ex0:8 * ex1].contiguous()
                    d.shape, d.first = T["shape"], int(i == 0)
                    d.lr4, d.lr4_clip_stride = lr4_t.data_ptr(), eh * ew * 4
                    d.x_lr, d.x_lr_clip_stride = xlr_t.data_ptr(), eh * ew * Cc
                    d.flow, d.flow_clip_stride = fl_t.data_ptr(), eh * ew * 2
                    d.fvs, d.fvs_clip_stride = fv_t.data_ptr(), 3 * 64 * eh * ew
                    d.mks, d.mks_clip_stride = mk_t.data_ptr(), 64 * eh * ew
                    d.state_hr, d.state_l1 = T["state_hr"].data_ptr(), T["state_l1"].data_ptr()
                    d.out, d.out_clip_stride = T["out"].data_ptr(), 3 * 64 * eh * ew
                    L.check(lib.crfp_dsv_frame(C.byref(d), C.byref(W), ws.data_ptr(), ws.numel(), st), f"tile {k} frame {i}")
                    # interior of this tile -> global output and global state
                    iy, ix = y0 - ey0, x0 - ex0
                    out[:, i, :, 8 * y0:8 * y1, 8 * x0:8 * x1] = T["out"][:, :, 8 * iy:8 * (iy + y1 - y0), 8 * ix:8 * (ix + x1 - x0)]
                    full_hr[:, 8 * y0:8 * y1, 8 * x0:8 * x1] = T["state_hr"][:, 8 * iy:8 * (iy + y1 - y0), 8 * ix:8 * (ix + x1 - x0)]
                    full_l1[:, 2 * y0:2 * y1, 2 * x0:2 * x1] = T["state_l1"][:, 2 * iy:2 * (iy + y1 - y0), 2 * ix:2 * (ix + x1 - x0)]
                if i + 1 < t:
                    # ---- the one exchange step of the frame: everybody gets every tile's interior state
                    if world > 1:
                        self._allgather_interiors(full_hr, full_l1, plan, world, rank)
                    for k in mine:   # refresh the halo (and keep the interior) of the extended tiles
                        (_, (ey0, ey1, ex0, ex1)) = plan[k]
                        tiles[k]["state_hr"].copy_(full_hr[:, 8 * ey0:8 * ey1, 8 * ex0:8 * ex1])
                        tiles[k]["state_l1"].copy_(full_l1[:, 2 * ey0:2 * ey1, 2 * ex0:2 * ex1])
            if world > 1 and self.gather_output:   # result gather: every rank contributed the interiors of its tiles, the rest is zero
                dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def _allgather_interiors(self, full_hr, full_l1, plan, world, rank):
        """The one exchange step per frame.  Every rank packs the interior state (HR + L1) of the tiles it owns into
        one flat buffer (fixed-size slots, so ragged tiles need no size negotiation), ONE all-gather moves all of
        them, and every rank unpacks the other ranks' interiors into its full-frame state."""
        n = full_hr.shape[0]
        nel = lambda y0, y1, x0, x1: n * (y1 - y0) * (x1 - x0) * (64 * 4 + 4 * 24)
        slot = max(nel(*it) for it, _ in plan)
        per_rank = (len(plan) + world - 1) // world
        send = torch.zeros(per_rank * slot, device=full_hr.device, dtype=full_hr.dtype)
        for k, ((y0, y1, x0, x1), _) in enumerate(plan):
            if k % world != rank:
                continue
            o = (k // world) * slot
            a = n * 64 * (y1 - y0) * (x1 - x0) * 4
            b = n * 4 * (y1 - y0) * (x1 - x0) * 24
            send[o:o + a].view(n, 8 * (y1 - y0), 8 * (x1 - x0), 4).copy_(full_hr[:, 8 * y0:8 * y1, 8 * x0:8 * x1])
            send[o + a:o + a + b].view(n, 2 * (y1 - y0), 2 * (x1 - x0), 24).copy_(full_l1[:, 2 * y0:2 * y1, 2 * x0:2 * x1])
        recv = torch.empty(world * per_rank * slot, device=full_hr.device, dtype=full_hr.dtype)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        for k, ((y0, y1, x0, x1), _) in enumerate(plan):
            owner = k % world
            if owner == rank:
                continue
            o = (owner * per_rank + k // world) * slot
            a = n * 64 * (y1 - y0) * (x1 - x0) * 4
            b = n * 4 * (y1 - y0) * (x1 - x0) * 24
            full_hr[:, 8 * y0:8 * y1, 8 * x0:8 * x1] = recv[o:o + a].view(n, 8 * (y1 - y0), 8 * (x1 - x0), 4)
            full_l1[:, 2 * y0:2 * y1, 2 * x0:2 * x1] = recv[o + a:o + a + b].view(n, 2 * (y1 - y0), 2 * (x1 - x0), 24)

    __call__ = forward
