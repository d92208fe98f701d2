"""Streaming front end of the hot path: a three-stage software pipeline over batches of image pairs.

The reference processes one pair at a time (``FeatureMatchingModel.forward``, lib/models/matching/model.py:23-40:
matcher -> solver, synchronously). On the GPU the two halves have opposite characters — the matcher is
tensor-core bound and fills every SM, the RANSAC back end is a handful of latency-bound fp64 blocks — so for
a *stream* of pairs they are overlapped:

    copy stream   : host (pinned) -> device of batch i+1
    engine stream : LoFTR forward of batch i          (then a device copy of its matches into slot i % 2)
    solver stream : essential-matrix RANSAC + scale-from-depth of batch i-1 (reads slot (i-1) % 2)

``submit()`` enqueues a batch and returns the finished result of the batch submitted one call earlier
(None on the first call); ``drain()`` returns the last one. Results are ``(R [B,3,3], t [B,3], inliers [B])``
float32 / int32 tensors in pinned host memory, written by an asynchronous device->host copy that is
complete when they are handed out.
"""
import torch

from . import pose_solver as _ps


class RelocPipeline:
    def __init__(self, engine, backend, pix_thr, scale_thr, num_hyp=_ps.DEFAULT_NUM_HYP, depth_hw=None):
        # `engine` may be a list of identical engines: consecutive batches then alternate between them, each on
        # its own stream, so that the HBM-bound kernels of one batch (attention, LayerNorm, dual softmax, token
        # GEMMs) share the SMs with the tensor-core-bound convolutions of the next.
        self.engines = list(engine) if isinstance(engine, (list, tuple)) else [engine]
        engine = self.engines[0]
        self.eng, self.be = engine, backend
        self.pix_thr, self.scale_thr, self.num_hyp = float(pix_thr), float(scale_thr), int(num_hyp)
        dev = engine.device
        self.dev = dev
        self.s_copy = torch.cuda.Stream(device=dev)
        self.s_solve = torch.cuda.Stream(device=dev)
        B, cap = engine.B, engine.capacity
        self.slots = []
        for _ in range(2):
            self.slots.append({
                "k0": torch.empty(cap, 2, device=dev), "k1": torch.empty(cap, 2, device=dev),
                "off": torch.empty(B + 1, dtype=torch.int32, device=dev),
                "rec": torch.empty(B, 13, device=dev),
                "rec_host": torch.zeros(B, 13).pin_memory(),
                "ev_fwd": torch.cuda.Event(), "ev_solved": torch.cuda.Event(), "ev_in": torch.cuda.Event(),
                "inputs": None, "busy": False,
            })
        self.stage = None
        if depth_hw is not None:                      # device staging for host inputs (two slots)
            h, w = depth_hw
            self.stage = [{"g0": torch.empty(B, h, w, device=dev), "g1": torch.empty(B, h, w, device=dev),
                           "d0": torch.empty(B, h, w, device=dev), "d1": torch.empty(B, h, w, device=dev),
                           "K0": torch.empty(B, 3, 3, device=dev), "K1": torch.empty(B, 3, 3, device=dev)}
                          for _ in range(2)]
        self.i = 0
        self._outs_all = [e.device_outputs() for e in self.engines]

    # ---------------------------------------------------------------- stages
    def _enqueue(self, g0, g1, d0, d1, K0, K1):
        """g0/g1/d0/d1/K0/K1: device tensors of this batch (already ordered after whatever produced them on the
        current stream or on self.s_copy via slot['ev_in']). K0 / K1 are the intrinsics of the reference and the
        query frame: Map-free frames have per-frame intrinsics (lib/datasets/mapfree.py:39-54) and the solver
        normalises each image's points with its own K (pose_solver.py:39-43, 150-151)."""
        eng, slot = self.engines[self.i % len(self.engines)], self.slots[self.i % 2]
        outs = self._outs_all[self.i % len(self.engines)]
        if slot["busy"]:                              # the solver of batch i-2 must be done with this slot
            eng.stream.wait_event(slot["ev_solved"])
        eng.forward_async(g0, g1)
        with torch.cuda.stream(eng.stream):
            slot["k0"].copy_(outs["mkpts0_f"], non_blocking=True)
            slot["k1"].copy_(outs["mkpts1_f"], non_blocking=True)
            slot["off"].copy_(outs["m_offsets"], non_blocking=True)
            slot["ev_fwd"].record(eng.stream)
        with torch.cuda.stream(self.s_solve):
            self.s_solve.wait_event(slot["ev_fwd"])
            R, t, n = self.be.emat_metric(slot["k0"], slot["k1"], slot["off"], K0, K1, d0, d1, self.pix_thr,
                                          self.scale_thr, self.num_hyp)
            B = eng.B
            slot["rec"][:, :9] = R.reshape(B, 9)
            slot["rec"][:, 9:12] = t.reshape(B, 3)
            slot["rec"][:, 12] = n.to(torch.float32)
            slot["rec_host"].copy_(slot["rec"], non_blocking=True)
            slot["ev_solved"].record(self.s_solve)
        for x in (d0, d1, K0, K1):
            x.record_stream(self.s_solve)
        slot["inputs"] = (g0, g1, d0, d1, K0, K1)     # keep alive until the slot is reused
        slot["busy"] = True

    def _collect(self, j):
        slot = self.slots[j % 2]
        slot["ev_solved"].synchronize()
        rec = slot["rec_host"]
        B = self.eng.B
        return rec[:, :9].reshape(B, 3, 3).clone(), rec[:, 9:12].clone(), rec[:, 12].to(torch.int32)

    # ---------------------------------------------------------------- public
    def submit_device(self, g0, g1, d0, d1, K0, K1=None):
        """Batch already resident on the device ([B,3,3] float32 intrinsics per image; K1=None: both frames share
        K0). Returns the result of the previous batch (or None)."""
        self._enqueue(g0, g1, d0, d1, K0, K0 if K1 is None else K1)
        self.i += 1
        return self._collect(self.i - 2) if self.i >= 2 else None

    def submit_host(self, g0, g1, d0, d1, K0, K1=None):
        """Batch in pinned host memory ([B,h,w] float32 images / depths, [B,3,3] intrinsics of each frame; K1=None:
        both frames share K0): the copy to the device runs on its own stream, overlapping the forward of the
        previous batch."""
        assert self.stage is not None, "construct the pipeline with depth_hw=(h, w) to feed host batches"
        st, slot = self.stage[self.i % 2], self.slots[self.i % 2]
        with torch.cuda.stream(self.s_copy):
            if slot["busy"]:                          # staging tensors of batch i-2 are read by its solver
                self.s_copy.wait_event(slot["ev_solved"])
            for k, src in (("g0", g0), ("g1", g1), ("d0", d0), ("d1", d1), ("K0", K0), ("K1", K0 if K1 is None else K1)):
                if isinstance(src, (list, tuple)):    # one pinned tensor per pair
                    for j, row in enumerate(src):
                        st[k][j].copy_(row, non_blocking=True)
                else:
                    st[k].copy_(src, non_blocking=True)
            slot["ev_in"].record(self.s_copy)
        self.engines[self.i % len(self.engines)].stream.wait_event(slot["ev_in"])
        self.s_solve.wait_event(slot["ev_in"])
        return self.submit_device(st["g0"], st["g1"], st["d0"], st["d1"], st["K0"], st["K1"])

    def drain(self):
        """Result of the last submitted batch (blocks until it is done)."""
        return self._collect(self.i - 1) if self.i >= 1 else None

    def reset(self):
        """Waits for everything in flight and forgets it: the next submit starts a new stream (returns None)."""
        for e in self.engines:
            e.stream.synchronize()
        self.s_copy.synchronize(); self.s_solve.synchronize()
        for slot in self.slots:
            slot["busy"] = False
            slot["inputs"] = None
        self.i = 0
