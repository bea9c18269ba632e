"""CUDA-graph replay of the whole forward hot path.

The fused path has no host round trip (every data-dependent size lives in device counters, buffers
are sized by capacity), so one forward over a batch of fixed point counts is a static launch
sequence: ~600 kernels of libu3d_b200 / cuDNN / cuBLAS on three streams. `GraphedForward` captures
it once and replays it per batch: the per-step CPU cost drops from ~10^3 launches to one
cudaGraphLaunch, which is what the end-to-end (host buffers in, boxes out) rate is bound by.
"""
import torch


class GraphedForward:
    """forward + NMSFreeCoder top-k (+ device post-processing: bottom-centre shift, per-class NMS,
    thresholds) for batches of a fixed signature (points per scene, C, dtype).

    run(host_points=None) copies the (pinned) host batch into the static device buffer, replays the
    graph and returns the static outputs (boxes (B,max_num,7|9), scores, labels, mask); the caller
    reads them back / synchronises as needed."""

    def __init__(self, model, lens, channels, random_point=None, warmup=3, postprocess=True):
        dev = next(model.parameters()).device
        self.model, self.lens = model, [int(n) for n in lens]
        B, nq = len(self.lens), model.num_query
        self.points = torch.zeros(sum(self.lens), channels, dtype=torch.float32, device=dev)
        off = torch.tensor([0] + list(torch.tensor(self.lens).cumsum(0).tolist()), dtype=torch.int32)
        self.pt_off = off.to(dev)
        self.random_point = random_point.to(dev) if random_point is not None else \
            torch.rand(B, nq, 3, device=dev)
        self.coder = model.pts_bbox_head.bbox_coder
        pp = model.pts_bbox_head.post_processing
        # the device-side get_bboxes (per-class NMS ...) rides in the same graph when the config's
        # post-processing is one the device path implements; otherwise the graph ends at the top-k
        self.postprocess = bool(postprocess) and (pp is None or pp.get("type") == "nms")
        # fixed shapes, replayed forever: let cuDNN time its engines for every dense conv during the warm-up instead of
        # trusting the heuristic pick (torch.backends.cudnn.benchmark; the capture below re-uses the cached choices).
        # U3D_CUDNN_BENCHMARK=0 keeps the heuristics.
        import os
        if os.environ.get("U3D_CUDNN_BENCHMARK", "1") != "0":
            torch.backends.cudnn.benchmark = True
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up: cuDNN plans, weight packing, allocator pools
            for _ in range(max(warmup, 1)):
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import ops
        n0 = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = self._forward()
        self.launches_per_replay = ops.launch_count() - n0   # libu3d kernels inside the graph

    def _forward(self):
        outs, _ = self.model.forward_raw(None, random_point=self.random_point,
                                         concat=(self.points, self.pt_off, self.lens))
        if self.postprocess:
            return self.model.pts_bbox_head.postprocess_fixed(outs)
        return self.coder.decode_fixed(outs)

    def load(self, host_points):
        """host_points: (Ntot,C) f32 (pinned for an async copy) -> static device buffer."""
        self.points.copy_(host_points, non_blocking=True)

    # ---- pipelined serving: H2D of batch i+1 and D2H of batch i-1 overlap the replay of batch i ----
    def start_pipeline(self):
        """Allocate the staging buffers of `submit` / `collect`: a second device point buffer fed by a copy
        stream, pinned host result buffers, and the events that order the three streams."""
        dev = self.points.device
        self._stage = torch.empty_like(self.points)
        self._copy_in, self._copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._staged, self._consumed, self._done = (torch.cuda.Event() for _ in range(3))
        self._host_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in self.outputs]
        self._dev_out = [torch.empty_like(t) for t in self.outputs]
        self._consumed.record(torch.cuda.current_stream())
        self._pending = False

    def submit(self, host_points):
        """Enqueue one batch: pinned host points -> staging buffer (copy stream), staging -> graph input + replay
        (compute stream), results -> pinned host buffers (copy-out stream). Returns immediately; `collect()`
        waits for the results of the batch submitted BEFORE this one (call it after the next submit)."""
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_in):
            self._copy_in.wait_event(self._consumed)          # the previous replay has copied the staging buffer out
            self._stage.copy_(host_points, non_blocking=True)
            self._staged.record(self._copy_in)
        cur.wait_event(self._staged)
        if self._pending:
            cur.wait_event(self._done)                        # results of the previous batch have left _dev_out
        self.points.copy_(self._stage, non_blocking=True)     # device-to-device, 10 MB: microseconds
        self._consumed.record(cur)
        self.graph.replay()
        for d, o in zip(self._dev_out, self.outputs):
            d.copy_(o, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._copy_out):
            self._copy_out.wait_event(ready)
            for h, d in zip(self._host_out, self._dev_out):
                h.copy_(d, non_blocking=True)
            self._done.record(self._copy_out)
        self._pending = True

    def collect(self):
        """Block until the most recently submitted batch's results are in the pinned host buffers; returns them."""
        self._done.synchronize()
        return self._host_out

    def run(self, host_points=None, redraw_random_group=False):
        """redraw_random_group=True refills the reference points of the random 4th query group in place
        before the replay (the reference draws `torch.rand` on every forward, uni3detr_head.py:445-447);
        the default keeps them fixed so repeated runs are reproducible."""
        if host_points is not None:
            self.load(host_points)
        if redraw_random_group:
            self.random_point.uniform_()
        self.graph.replay()
        return self.outputs
