"""One big stereo pair across several GPUs: the disparity-slab partition of SURVEY.md 8(e) (BASELINE config C5).

The reference has no such mode (its only parallelism is disjoint pair windows, match.py:26-28); the stage order
and every per-cell result are those of match.py:131-175, only *where* a cell is computed changes:

  features              row bands (each rank runs the net on its rows plus the 5-row halo of the receptive field),
                        then every rank receives the other bands: the cost volume needs whole feature maps
  cost volume, CBCA     rank g owns disparities [d_base_g, d_base_g + d_count_g): independent per disparity plane
                        (pf:94-95 runs along w inside one d; cross regions ignore d)
  SGM                   needs all disparities of a pixel (pf:549-566): the volumes are re-partitioned
                        d-slabs -> row slabs (the two horizontal passes) -> column slabs (the two vertical passes)
                        -> d-slabs, three exchanges per volume (grouped send/recv = all-to-all over NVLink),
                        blocks packed / unpacked by mccnn_copy3d
  WTA                   per-slab first minimum + its cost, all-gather, first strict minimum in slab order
  sub-pixel             the three cells around d* may sit in a neighbour slab (the slab seam): every slab contributes
                        the cells it owns, summed over the ranks (exact: the other terms are 0)
  LR check, median, bilateral   O(H*W), replicated.

``SlabPlan`` is pure index arithmetic (tested on CPU), ``SlabRank`` owns one rank's device buffers and issues its
C-ABI calls, ``run_slabs`` drives one or more ranks through the phases with a communicator: ``DistComm``
(torch.distributed: NCCL on GPUs, gloo in the CPU tests) or ``LocalComm`` (all ranks inside one process on one
GPU: used to check the partition against the single-GPU pipeline bit for bit).
"""
import ctypes

import numpy as np

try:
    from . import _ffi
    from . import process_functional as _pf
    from .pipeline import DEFAULTS
except ImportError:
    import _ffi
    import process_functional as _pf
    from pipeline import DEFAULTS


def _split(n, parts):
    """Balanced contiguous split of range(n): list of (lo, hi)."""
    base, extra = divmod(int(n), int(parts))
    out, lo = [], 0
    for r in range(parts):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


class SlabPlan(object):
    """Who owns what.  Disparities are split in whole 16-byte granules (4 disparities) so that a slab is itself
    an HWD volume; rows and columns are split evenly."""

    def __init__(self, H, W, D, world):
        self.H, self.W, self.D, self.world = int(H), int(W), int(D), int(world)
        self.G = _ffi.dpitch(self.D) // 4
        assert self.world >= 1 and self.G >= self.world and self.H >= self.world and self.W >= self.world, \
            "need at least one disparity granule, one row and one column per rank"
        self.granules = _split(self.G, self.world)
        self.rows = _split(self.H, self.world)
        self.cols = _split(self.W, self.world)

    def d_base(self, r):
        return 4 * self.granules[r][0]

    def d_count(self, r):
        """Disparities of slab r (the last slab may end inside its last granule)."""
        return min(self.D, 4 * self.granules[r][1]) - self.d_base(r)

    def g_count(self, r):
        return self.granules[r][1] - self.granules[r][0]

    def h_count(self, r):
        return self.rows[r][1] - self.rows[r][0]

    def w_count(self, r):
        return self.cols[r][1] - self.cols[r][0]

    def region_floats(self):
        """Floats of the largest d-slab / row-slab / column-slab volume any rank holds (the size of one region of
        the symmetric arena: identical on every rank, as peer memory requires)."""
        Dp = 4 * self.G
        return max(max(self.H * self.W * 4 * self.g_count(r), self.h_count(r) * self.W * Dp, self.H * self.w_count(r) * Dp)
                   for r in range(self.world))

    def owner_of_disparity(self, d):
        for r in range(self.world):
            if self.d_base(r) <= d < self.d_base(r) + self.d_count(r):
                return r
        raise ValueError(d)


# ---------------------------------------------------------------------------------------------- communicators
class LocalComm(object):
    """All ranks live in this process (one GPU): an exchange is a set of device copies."""

    def __init__(self, world):
        self.world = int(world)

    def exchange_begin(self, sends, recvs):
        """sends[i][j] = block rank i sends to rank j; recvs[j][i] = where rank j receives it."""
        for i in range(self.world):
            for j in range(self.world):
                if recvs[j][i].data_ptr() != sends[i][j].data_ptr():
                    recvs[j][i].copy_(sends[i][j])
        return None

    def exchange_end(self, handle):
        pass

    def exchange(self, sends, recvs):
        self.exchange_end(self.exchange_begin(sends, recvs))

    def make_arenas(self, nfloats):
        """One arena per rank for the volumes peers write into (here: ordinary device memory of the one GPU)."""
        torch = _pf._torch()
        self.arenas = [torch.empty(int(nfloats), dtype=torch.float32, device=_pf._dev()) for _ in range(self.world)]
        return self.arenas

    def peer_bases(self, local_index):
        return [a.data_ptr() for a in self.arenas]

    def barrier(self):
        pass

    def all_gather(self, parts):
        """parts[i] = rank i's tensor; returns per rank the stacked [world, ...] tensor."""
        import torch
        g = torch.stack(list(parts), 0)
        return [g for _ in parts]

    def all_reduce_sum(self, parts):
        acc = parts[0].clone()
        for p in parts[1:]:
            acc += p
        for p in parts:
            p.copy_(acc)


class DistComm(object):
    """One rank per process over torch.distributed (NCCL on GPUs; gloo for the CPU tests of the plumbing)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()

    def exchange_begin(self, sends, recvs):
        """Start the grouped send/recv (one ncclGroupStart/End: an all-to-all over NVLink); it runs on the
        communicator's stream, after everything already queued on the current stream, and alongside what is
        queued next -- exchange_end() makes the current stream wait for it."""
        dist = self.dist
        send, recv = sends[0], recvs[0]
        if recv[self.rank].data_ptr() != send[self.rank].data_ptr():
            recv[self.rank].copy_(send[self.rank])
        ops = []
        for j in range(self.world):
            if j != self.rank:
                ops.append(dist.P2POp(dist.irecv, recv[j], j))
        for j in range(self.world):
            if j != self.rank:
                ops.append(dist.P2POp(dist.isend, send[j], j))
        return dist.batch_isend_irecv(ops) if ops else []

    def exchange_end(self, handle):
        for w in handle:
            w.wait()

    def exchange(self, sends, recvs):
        self.exchange_end(self.exchange_begin(sends, recvs))

    def make_arenas(self, nfloats):
        """The arena peers write into: symmetric memory (same size on every rank), mapped into every rank's address
        space over NVLink by the rendezvous.  The allocation is local and may fail on some ranks only; the rendezvous
        is collective.  So the ranks vote after the allocation and enter the rendezvous only if every one of them
        succeeded -- otherwise all raise together (a rank that failed alone would leave the others blocked in the
        rendezvous)."""
        import torch
        ok, why, arena, symm_mem = 1, "", None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            arena = symm_mem.empty(int(nfloats), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        except (ImportError, AttributeError, RuntimeError) as e:
            ok, why = 0, repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            raise RuntimeError("symmetric memory unavailable: %s" % (why or "on another rank"))
        self.arena = arena
        self.symm = symm_mem.rendezvous(self.arena, self.dist.group.WORLD)
        return [self.arena]

    def peer_bases(self, local_index):
        return [int(p) for p in self.symm.buffer_ptrs]

    def barrier(self):
        self.symm.barrier()                          # device-side, on the current stream

    def all_gather(self, parts):
        import torch
        t = parts[0].contiguous()
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t) if t.is_cuda else self.dist.all_gather(list(out.unbind(0)), t)
        return [out]

    def all_reduce_sum(self, parts):
        self.dist.all_reduce(parts[0], op=self.dist.ReduceOp.SUM)


# ---------------------------------------------------------------------------------------------- one rank
def _copy3d(src, src_off, dst, dst_off, n0, n1, n2, ss0, ss1, ds0, ds1):
    """mccnn_copy3d on two float32 device tensors; offsets and strides in 16-byte granules."""
    _ffi.call("mccnn_copy3d", ctypes.c_void_p(src.data_ptr() + 16 * int(src_off)),
              ctypes.c_void_p(dst.data_ptr() + 16 * int(dst_off)), int(n0), int(n1), int(n2), int(ss0), int(ss1),
              int(ds0), int(ds1), _ffi.stream_ptr())


class SlabRank(object):
    """Buffers and C-ABI calls of one rank of the partition."""

    def __init__(self, plan, rank, checkpoint=None, cbca_mode=None, arena=None, **hp):
        torch = _pf._torch()
        self.torch = torch
        self.plan, self.rank = plan, int(rank)
        self.hp = dict(DEFAULTS)
        self.hp.update(hp)
        H, W, D, N = plan.H, plan.W, plan.D, plan.world
        assert D >= 2 and W >= D + 2, "need ndisp >= 2 and W >= ndisp + 2 (pf:94-95, :547-566)"
        dev = _pf._dev()
        f32 = torch.float32
        e = lambda *shape, dtype=f32: torch.empty(shape, dtype=dtype, device=dev)
        self.pad = (int(self.hp["patch_size"]) - 1) // 2
        self.weights = _pf.resolve_weights(checkpoint, num_layers=self.pad)
        self.Dl, self.Dlp = plan.d_count(rank), 4 * plan.g_count(rank)
        self.dbase = plan.d_base(rank)
        self.Dp = 4 * plan.G
        self.h0, self.h1 = plan.rows[rank]
        self.w0, self.w1 = plan.cols[rank]
        Hr, Wc = self.h1 - self.h0, self.w1 - self.w0
        self.img = [e(H, W), e(H, W)]
        self.feat = [e(H, W, 64), e(H, W, 64)]
        nb = int(_ffi.lib().mccnn_features_scratch_bytes(H, W, self.pad, self.pad))
        self.feat_scratch = e((nb + 3) // 4)
        # d-slab volumes (A: cost volume / CBCA2 output, B: CBCA1 output and SGM result, S: CBCA scratch)
        self.volA = [e(H, W, self.Dlp), e(H, W, self.Dlp)]
        self.volS = e(H, W, self.Dlp)
        # B and the same cells as row slabs (all disparities of rows [h0, h1)) and column slabs (columns [w0, w1)):
        # either private buffers filled through staged exchanges, or six regions of an arena that peers write into
        # directly (region k of every rank's arena starts at k * plan.region_floats())
        self.region = plan.region_floats()
        if arena is None:
            self.volB = [e(H, W, self.Dlp), e(H, W, self.Dlp)]
            self.rowv = [e(Hr, W, self.Dp), e(Hr, W, self.Dp)]
            self.colv = [e(H, Wc, self.Dp), e(H, Wc, self.Dp)]
            # exchange staging: every re-partition moves exactly one slab's worth of cells out and in
            nstage = max(H * W * self.Dlp, Hr * W * self.Dp, H * Wc * self.Dp)
            self.stage_out = [e(nstage), e(nstage)]        # per volume: the two volumes' exchanges are in flight together
            self.stage_in = [e(nstage), e(nstage)]
        else:
            reg = lambda k, *shape: arena[k * self.region:k * self.region + int(np.prod(shape))].view(*shape)
            self.volB = [reg(0, H, W, self.Dlp), reg(1, H, W, self.Dlp)]
            self.rowv = [reg(2, Hr, W, self.Dp), reg(3, Hr, W, self.Dp)]
            self.colv = [reg(4, H, Wc, self.Dp), reg(5, H, Wc, self.Dp)]
        self.arms = [e(H, W, 4, dtype=torch.uint8), e(H, W, 4, dtype=torch.uint8)]
        self.count = [e(H, W, dtype=torch.int32), e(H, W, dtype=torch.int32)]
        ns = int(_ffi.lib().mccnn_sgm_scratch_bytes(H, W, D))
        self.sgm_flags = e((ns + 3) // 4, dtype=torch.int32)
        self.wta_local = e(4, H, W)                        # (disp L, min L, disp R, min R) of this slab
        self.disp = [e(H, W), e(H, W)]
        self.tmp = [e(H, W), e(H, W)]
        self.triple = e(3, H, W)
        self.labels = e(H, W, dtype=torch.int32)
        self.table = _pf._to_dev(_pf.bilateral_table(5, 5, 0, self.hp["blur_sigma"]))
        self.cbca_mode = _pf.CBCA_MODE if cbca_mode is None else int(cbca_mode)
        self.cbca_modes = [_pf.CBCA_SEPARABLE if self.cbca_mode == _pf.CBCA_AUTO else self.cbca_mode] * 2
        self.result = None

    def set_images(self, left_image, right_image):
        for i, im in enumerate((left_image, right_image)):
            self.img[i].copy_(_pf._image2d(im))

    # ---- phases --------------------------------------------------------------------------------
    def features_band(self):
        """Features of this rank's rows (match.py:132): the net sees 5 rows beyond the band on either side (its
        receptive field is 11), so rows [h0 - 5, h1 + 5) are run and only [h0, h1) are kept -- the rest arrives from
        the ranks that own it."""
        p, call, sp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.plan
        a, b = max(0, self.h0 - self.pad), min(pl.H, self.h1 + self.pad)
        for i in range(2):
            call("mccnn_features_prepared", ctypes.c_void_p(self.img[i].data_ptr() + 4 * a * pl.W), b - a, pl.W, self.pad, self.pad,
                 self.weights.w_table, self.weights.b_table, p(self.weights.prepared),
                 ctypes.c_void_p(self.feat[i].data_ptr() + 4 * a * pl.W * 64), p(self.feat_scratch), sp())

    def send_features(self, i):
        return [self.feat[i][self.h0:self.h1] for _ in range(self.plan.world)]

    def recv_features(self, i):
        return [self.feat[i][lo:hi] for lo, hi in self.plan.rows]

    def front(self, bases=None):
        """This slab of the cost volume, cross arms, CBCA x iters1 (match.py:137-143).  With `bases` (the ranks' arena
        addresses) the last column pass of the aggregation stores straight into the row slabs of the owners."""
        p, call, sp, hp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.hp, self.plan
        H, W, D = pl.H, pl.W, pl.D
        call("mccnn_cost_volume_slab", p(self.feat[0]), p(self.feat[1]), p(self.volA[0]), p(self.volA[1]), H, W, 64, D,
             self.dbase, self.Dl, sp())
        for i in range(2):
            call("mccnn_cross_arms", p(self.img[i]), p(self.arms[i]), p(self.count[i]), H, W,
                 ctypes.c_float(np.float32(hp["cbca_intensity"])), int(hp["cbca_distance"]), sp())
            # (every rank sees the same images, so every rank makes the same choice)
            self.cbca_modes[i] = _pf.cbca_auto_mode(self.arms[i]) if self.cbca_mode == _pf.CBCA_AUTO else self.cbca_mode
        it1 = int(hp["cbca_num_iterations1"])
        # the fused hand-over exists for the chained separable rounds only: another mode aggregates in place and pushes
        if bases is None or it1 < 1 or any(m != _pf.CBCA_SEPARABLE for m in self.cbca_modes):
            for v in range(2):
                self._cbca(v, self.volA, self.volB, it1)
            if bases is not None:
                self.push_rows(bases)
            return
        n = pl.world
        bounds = (ctypes.c_int * (n + 1))(*([lo for lo, _ in pl.rows] + [H]))
        for v in range(2):
            dst = (ctypes.c_void_p * n)(*[bases[j] + 4 * (2 + v) * self.region for j in range(n)])
            call("mccnn_cbca_to", p(self.volA[v]), p(self.volB[v]), p(self.volS), p(self.arms[v]), p(self.count[v]), self.Dl,
                 H, W, it1, int(hp["cbca_distance"]), n, bounds, dst, pl.granules[self.rank][0], pl.G, sp())

    def _cbca(self, i, src, dst, iters):
        p, call, sp, hp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.hp, self.plan
        call("mccnn_cbca", p(src[i]), p(dst[i]), p(self.volS), p(self.arms[i]), p(self.count[i]), self.Dl, pl.H, pl.W,
             iters, int(hp["cbca_distance"]), int(self.cbca_modes[i]), sp())

    def _views(self, flat, shapes):
        out, off = [], 0
        for shp in shapes:
            n = int(np.prod(shp))
            out.append(flat[off:off + n].view(*shp))
            off += n
        return out

    # d-slabs -> row slabs: block for rank j = its rows of my slab (already contiguous)
    def send_rows(self, v):
        return [self.volB[v][lo:hi] for lo, hi in self.plan.rows]

    def recv_rows(self, v):
        pl = self.plan
        Hr = self.h1 - self.h0
        return self._views(self.stage_in[v], [(Hr, pl.W, 4 * pl.g_count(j)) for j in range(pl.world)])

    def unpack_rows(self, v, blocks):
        pl = self.plan
        n = (self.h1 - self.h0) * pl.W
        for j, blk in enumerate(blocks):
            gj = pl.g_count(j)
            _copy3d(blk, 0, self.rowv[v], pl.granules[j][0], 1, n, gj, 0, gj, 0, pl.G)

    def sgm_rows(self, v):
        """(0,1) then (0,-1), pf:195-198, on rows [h0, h1) of volume v (0 left, 1 right)."""
        p, call, sp, hp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.hp, self.plan
        off = 4 * self.h0 * pl.W
        il = ctypes.c_void_p(self.img[0].data_ptr() + off)
        ir = ctypes.c_void_p(self.img[1].data_ptr() + off)
        call("mccnn_sgm_passes_slab", p(self.rowv[0]) if v == 0 else None, p(self.rowv[1]) if v == 1 else None, il, ir,
             p(self.sgm_flags), pl.D, self.h1 - self.h0, pl.W, 0, pl.W, 0, float(hp["sgm_P1"]), float(hp["sgm_P2"]),
             float(hp["sgm_Q1"]), float(hp["sgm_Q2"]), float(hp["sgm_D"]), float(hp["sgm_V"]), sp())

    # row slabs -> column slabs: block for rank j = its columns of my rows (packed), lands as rows of its slab
    def send_cols(self, v):
        pl = self.plan
        Hr = self.h1 - self.h0
        views = self._views(self.stage_out[v], [(Hr, pl.w_count(j), self.Dp) for j in range(pl.world)])
        for j, blk in enumerate(views):
            wj = pl.w_count(j)
            _copy3d(self.rowv[v], pl.cols[j][0] * pl.G, blk, 0, Hr, wj, pl.G, pl.W * pl.G, pl.G, wj * pl.G, pl.G)
        return views

    def recv_cols(self, v):
        return [self.colv[v][lo:hi] for lo, hi in self.plan.rows]

    def sgm_cols(self, v):
        """(-1,0) then (1,0), pf:203-208, on columns [w0, w1) of volume v."""
        p, call, sp, hp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.hp, self.plan
        call("mccnn_sgm_passes_slab", p(self.colv[0]) if v == 0 else None, p(self.colv[1]) if v == 1 else None,
             p(self.img[0]), p(self.img[1]), p(self.sgm_flags), pl.D, pl.H, pl.W, self.w0, self.w1 - self.w0, 1,
             float(hp["sgm_P1"]), float(hp["sgm_P2"]), float(hp["sgm_Q1"]), float(hp["sgm_Q2"]), float(hp["sgm_D"]),
             float(hp["sgm_V"]), sp())

    # column slabs -> d-slabs: block for rank j = its disparities of my columns (packed); unpacked into columns
    def send_slabs(self, v):
        pl = self.plan
        Wc = self.w1 - self.w0
        views = self._views(self.stage_out[v], [(pl.H, Wc, 4 * pl.g_count(j)) for j in range(pl.world)])
        for j, blk in enumerate(views):
            gj = pl.g_count(j)
            _copy3d(self.colv[v], pl.granules[j][0], blk, 0, 1, pl.H * Wc, gj, 0, pl.G, 0, gj)
        return views

    def recv_slabs(self, v):
        pl = self.plan
        return self._views(self.stage_in[v], [(pl.H, pl.w_count(j), self.Dlp) for j in range(pl.world)])

    def unpack_slabs(self, v, blocks):
        pl = self.plan
        gl = pl.g_count(self.rank)
        for j, blk in enumerate(blocks):
            wj = pl.w_count(j)
            _copy3d(blk, 0, self.volB[v], pl.cols[j][0] * gl, pl.H, wj, gl, wj * gl, gl, pl.W * gl, gl)

    # ---- the same re-partitions over peer memory: no staging, no unpacking --------------------------------
    def push_rows(self, bases):
        """d-slabs -> row slabs: my disparities of rank j's rows go straight into rank j's row slab."""
        pl = self.plan
        gme = pl.g_count(self.rank)
        for v in range(2):
            for j, (lo, hi) in enumerate(pl.rows):
                dst = bases[j] + 4 * (2 + v) * self.region + 16 * pl.granules[self.rank][0]
                _ffi.call("mccnn_copy3d", ctypes.c_void_p(self.volB[v].data_ptr() + 4 * lo * pl.W * self.Dlp),
                          ctypes.c_void_p(dst), 1, (hi - lo) * pl.W, gme, 0, gme, 0, pl.G, _ffi.stream_ptr())

    def _tables(self, bases, first_region, bounds):
        n = self.plan.world
        b = (ctypes.c_int * (n + 1))(*bounds)
        left = (ctypes.c_void_p * n)(*[bases[j] + 4 * first_region * self.region for j in range(n)])
        right = (ctypes.c_void_p * n)(*[bases[j] + 4 * (first_region + 1) * self.region for j in range(n)])
        return n, b, left, right

    def sgm_rows_to(self, bases):
        """(0,1) in place, then (0,-1) storing every pixel into the column slab of the rank that owns its column."""
        p, call, sp, hp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.hp, self.plan
        off = 4 * self.h0 * pl.W
        n, b, left, right = self._tables(bases, 4, [lo for lo, _ in pl.cols] + [pl.W])
        call("mccnn_sgm_passes_slab_to", p(self.rowv[0]), p(self.rowv[1]), ctypes.c_void_p(self.img[0].data_ptr() + off),
             ctypes.c_void_p(self.img[1].data_ptr() + off), p(self.sgm_flags), pl.D, self.h1 - self.h0, pl.W, 0, pl.W, 0,
             float(hp["sgm_P1"]), float(hp["sgm_P2"]), float(hp["sgm_Q1"]), float(hp["sgm_Q2"]), float(hp["sgm_D"]),
             float(hp["sgm_V"]), n, b, left, right, self.h0, sp())

    def sgm_cols_to(self, bases):
        """(-1,0) in place, then (1,0) storing every granule into the disparity slab of the rank that owns it."""
        p, call, sp, hp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.hp, self.plan
        n, b, left, right = self._tables(bases, 0, [lo for lo, _ in pl.granules] + [pl.G])
        call("mccnn_sgm_passes_slab_to", p(self.colv[0]), p(self.colv[1]), p(self.img[0]), p(self.img[1]), p(self.sgm_flags),
             pl.D, pl.H, pl.W, self.w0, self.w1 - self.w0, 1, float(hp["sgm_P1"]), float(hp["sgm_P2"]),
             float(hp["sgm_Q1"]), float(hp["sgm_Q2"]), float(hp["sgm_D"]), float(hp["sgm_V"]), n, b, left, right, 0, sp())

    def cbca2(self, v):
        """CBCA x iters2 of volume v (match.py:154-155)."""
        self._cbca(v, self.volB, self.volA, int(self.hp["cbca_num_iterations2"]))

    def wta(self):
        """This slab's winners and their costs (match.py:159)."""
        p, call, sp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.plan
        for i in range(2):
            call("mccnn_wta_slab", p(self.volA[i]), p(self.wta_local[2 * i]), p(self.wta_local[2 * i + 1]), self.Dl, pl.H,
                 pl.W, self.dbase, sp())
        return self.wta_local

    def combine(self, gathered):
        """gathered [world][4][H][W]: first minimum in slab order, then LR check / interpolation (match.py:163)."""
        p, call, sp, pl = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.plan
        P = pl.H * pl.W
        base = gathered.data_ptr()
        for i in range(2):
            disps = ctypes.c_void_p(base + 4 * (2 * i) * P)
            mins = ctypes.c_void_p(base + 4 * (2 * i + 1) * P)
            call("mccnn_wta_combine", mins, disps, p(self.disp[i]), pl.world, 4 * P, pl.H, pl.W, sp())
        call("mccnn_lr_interp", p(self.disp[0]), p(self.disp[1]), p(self.tmp[0]), p(self.labels), pl.H, pl.W, pl.D, sp())
        call("mccnn_subpixel_gather", p(self.tmp[0]), p(self.volA[0]), p(self.triple), self.Dl, pl.H, pl.W, self.dbase,
             pl.D, sp())
        return self.triple

    def finish(self):
        """sub-pixel from the summed triple, median, bilateral (match.py:167-175)."""
        p, call, sp, pl, hp = _ffi.ptr, _ffi.call, _ffi.stream_ptr, self.plan, self.hp
        call("mccnn_subpixel_triple", p(self.tmp[0]), p(self.triple), p(self.tmp[1]), pl.D, pl.H, pl.W, sp())
        call("mccnn_median", p(self.tmp[1]), p(self.tmp[0]), pl.H, pl.W, 5, 5, sp())
        call("mccnn_bilateral", p(self.img[0]), p(self.tmp[0]), p(self.tmp[1]), p(self.table), pl.H, pl.W, 5, 5,
             ctypes.c_float(np.float32(hp["blur_threshold"])), sp())
        self.result = self.tmp[1]
        return self.result


def run_slabs(ranks, comm, marks=None):
    """Drive the ranks held by this process (one under torch.distributed, all of them with LocalComm) through the
    pipeline; returns each rank's final disparity map (identical on every rank).  `marks`, if a list, receives
    (phase name, CUDA event recorded after the phase) pairs."""
    def mark(name):
        if marks is not None:
            ev = ranks[0].torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    # The two volumes are independent until WTA: the exchange of one runs under the passes of the other.
    begin, end = comm.exchange_begin, comm.exchange_end
    mark("start")
    for r in ranks:
        r.features_band()
    if ranks[0].plan.world > 1:
        hf = [begin([r.send_features(i) for r in ranks], [r.recv_features(i) for r in ranks]) for i in range(2)]
        for h in hf:
            end(h)
    mark("features")
    for r in ranks:
        r.front()
    mark("front")
    recv_rows = [[r.recv_rows(v) for r in ranks] for v in range(2)]
    h_rows = [begin([r.send_rows(v) for r in ranks], recv_rows[v]) for v in range(2)]
    h_cols, h_slabs = [None, None], [None, None]
    for v in range(2):
        end(h_rows[v])
        for r, blocks in zip(ranks, recv_rows[v]):
            r.unpack_rows(v, blocks)
        for r in ranks:
            r.sgm_rows(v)
        h_cols[v] = begin([r.send_cols(v) for r in ranks], [r.recv_cols(v) for r in ranks])
    recv_slabs = [None, None]
    for v in range(2):
        end(h_cols[v])
        for r in ranks:
            r.sgm_cols(v)
        recv_slabs[v] = [r.recv_slabs(v) for r in ranks]
        h_slabs[v] = begin([r.send_slabs(v) for r in ranks], recv_slabs[v])
    mark("sgm_and_exchanges")
    for v in range(2):
        end(h_slabs[v])
        for r, blocks in zip(ranks, recv_slabs[v]):
            r.unpack_slabs(v, blocks)
        for r in ranks:
            r.cbca2(v)
    mark("cbca2")
    gathered = comm.all_gather([r.wta() for r in ranks])
    comm.all_reduce_sum([r.combine(g) for r, g in zip(ranks, gathered)])
    out = [r.finish() for r in ranks]
    mark("wta_refine")
    return out


def run_slabs_p2p(ranks, comm, marks=None):
    """run_slabs with the three re-partitions done over peer memory (the ranks' volumes live in arenas obtained
    from comm.make_arenas): all three hand-overs are fused into
    the kernels that produce the cells -- the last column pass of the first aggregation stores each row into the row
    slab of the rank that owns it, the last horizontal pass stores each pixel into
    the column slab of the rank that owns its column, the last vertical pass each granule into the disparity slab of
    the rank that owns it -- so the transfers ride under the recurrence and nothing is packed, staged or unpacked.
    A device-side barrier separates the phases."""
    def mark(name):
        if marks is not None:
            ev = ranks[0].torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    for r in ranks:
        r.features_band()
    if ranks[0].plan.world > 1:
        hf = [comm.exchange_begin([r.send_features(i) for r in ranks], [r.recv_features(i) for r in ranks]) for i in range(2)]
        for h in hf:
            comm.exchange_end(h)
    mark("features")
    bases = [comm.peer_bases(i) for i in range(len(ranks))]
    for r, b in zip(ranks, bases):
        r.front(b)
    mark("front")
    comm.barrier()
    for r, b in zip(ranks, bases):
        r.sgm_rows_to(b)
    comm.barrier()
    for r, b in zip(ranks, bases):
        r.sgm_cols_to(b)
    comm.barrier()
    mark("sgm_and_exchanges")
    for r in ranks:
        for v in range(2):
            r.cbca2(v)
    mark("cbca2")
    gathered = comm.all_gather([r.wta() for r in ranks])
    comm.all_reduce_sum([r.combine(g) for r, g in zip(ranks, gathered)])
    out = [r.finish() for r in ranks]
    comm.barrier()                                   # (nobody starts the next pair's pushes before every rank is done)
    mark("wta_refine")
    return out


class SlabMatcher(object):
    """This process's rank of a disparity-slab partitioned pair under torch.distributed."""

    def __init__(self, H, W, ndisp, checkpoint=None, transport="auto", **hp):
        """transport "p2p": volumes in symmetric memory, re-partitions fused into the kernels over NVLink peer
        stores; "nccl": staged exchanges with grouped ncclSend/ncclRecv; "auto": p2p while a slab's run of
        disparities per pixel is at least 512 bytes (peer stores of shorter runs are partial-line writes and lose to
        bulk transfers: measured at 2000x3000x400, p2p / nccl ms per pair = 222 / 236 on 2 GPUs, 136 / 136 on 4,
        85 / 79 on 8)."""
        assert transport in ("auto", "p2p", "nccl")
        self.comm = DistComm()
        self.plan = SlabPlan(H, W, ndisp, self.comm.world)
        if transport == "auto":
            runs = min(self.plan.g_count(r) for r in range(self.plan.world)) * 16
            transport = "p2p" if runs >= 512 else "nccl"
        self.transport = transport
        arena = None
        if transport == "p2p":
            # symmetric memory is a young torch API: every rank tries, and all fall back to the staged NCCL
            # transport together if any of them cannot map its peers (still GPU to GPU, never through the host)
            # (make_arenas votes before its collective step, so it raises on every rank or on none)
            try:
                arena = self.comm.make_arenas(6 * self.plan.region_floats())[0]
            except RuntimeError as e:
                if self.comm.rank == 0:
                    print("slab: peer-memory transport unavailable (%s); using staged NCCL exchanges" % e)
                arena, transport = None, "nccl"
                self.transport = transport
        self.rank = SlabRank(self.plan, self.comm.rank, checkpoint=checkpoint, arena=arena, **hp)
        self.H, self.W, self.D = self.plan.H, self.plan.W, self.plan.D
        self.hp = self.rank.hp

    def set_images(self, left_image, right_image):
        self.rank.set_images(left_image, right_image)

    def _run(self, marks=None):
        fn = run_slabs_p2p if self.transport == "p2p" else run_slabs
        return fn([self.rank], self.comm, marks)[0]

    def run(self):
        return self._run()

    def run_timed(self):
        """run() with a CUDA event after every phase: {phase: milliseconds} on this rank."""
        marks = []
        self._run(marks)
        self.rank.torch.cuda.synchronize()
        return {name: marks[i - 1][1].elapsed_time(ev) for i, (name, ev) in enumerate(marks) if i > 0}

    def run_host(self, left_image, right_image):
        """NumPy images in, NumPy disparity out (every rank uploads the pair, every rank ends with the map)."""
        torch = self.rank.torch
        if not hasattr(self, "_host_in"):
            H, W = self.H, self.W
            self._host_in = [torch.empty((H, W), dtype=torch.float32, pin_memory=True) for _ in range(2)]
            self._host_out = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        for i, im in enumerate((left_image, right_image)):
            a = np.asarray(im, dtype=np.float32)
            if a.ndim == 3:
                a = a[:, :, 0]
            self._host_in[i].numpy()[...] = a
            self.rank.img[i].copy_(self._host_in[i], non_blocking=True)
        d = self.run()
        self._host_out.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_out.numpy().copy()
