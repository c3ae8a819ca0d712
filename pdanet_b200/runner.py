"""Scene-batch inference runner: the call a user makes to push host point clouds through PDA-SSD.

`SceneRunner.infer(points_host)` takes the reference's batch layout — a (B*N, 5) fp32 array
[batch_idx, x, y, z, intensity] in host memory (pcdet/datasets/dataset.py:173-178) — stages it through a
pinned buffer, runs the detector on the runner's device / stream and returns the per-scene predictions
as host tensors.  Scenes are independent: with several GPUs each rank owns a
`SceneRunner` and a disjoint slice of the scenes (`shard_scenes`), with no collective on the data path.
"""
from __future__ import annotations

from typing import List

import torch

from .config import load_config
from .iassd import build_model


def shard_scenes(num_scenes: int, world_size: int, rank: int) -> range:
    """Contiguous block of scene ids owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(num_scenes, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class SceneRunner:
    def __init__(self, cfg="kitti", device="cuda:0", batch_size=16, num_points=None, seed=0, model=None,
                 tc_passes=None):
        self.cfg = load_config(cfg) if isinstance(cfg, str) else cfg
        self.device = torch.device(device)
        self.batch_size = batch_size
        self.num_points = num_points or self.cfg.NUM_POINTS
        if model is None:
            torch.manual_seed(seed)
            model = build_model(self.cfg)
        self.model = model.to(self.device).eval()
        if tc_passes is not None:  # tensor-core product mode of every module that has one (see tc_linear.PackedLinear)
            for m in self.model.modules():
                if hasattr(m, "tc_passes"):
                    m.tc_passes = tc_passes
        rows = self.batch_size * self.num_points
        self._pinned = torch.empty(rows, 5, dtype=torch.float32).pin_memory()
        self._dev = torch.empty(rows, 5, dtype=torch.float32, device=self.device)
        self.h2d_bytes = self._pinned.numel() * 4
        self.d2h_bytes = 0

    @torch.no_grad()
    def infer_device(self, points_dev: torch.Tensor, batch_size=None):
        """points already in HBM -> pred_dicts on the device."""
        return self.model({"batch_size": batch_size or self.batch_size, "points": points_dev})[0]

    @torch.no_grad()
    def infer(self, points_host: torch.Tensor) -> List[dict]:
        """Host (B*N,5) -> host predictions; H2D and D2H copies included."""
        assert points_host.shape == self._pinned.shape and points_host.dtype == torch.float32
        if points_host.is_pinned():
            src = points_host
        else:
            self._pinned.copy_(points_host)
            src = self._pinned
        with torch.cuda.device(self.device):
            self._dev.copy_(src, non_blocking=True)
            preds = self.infer_device(self._dev)
            out, nbytes = [], 0
            for p in preds:
                host = {k: v.to("cpu", non_blocking=True) for k, v in p.items()}
                nbytes += sum(v.numel() * v.element_size() for v in host.values())
                out.append(host)
            torch.cuda.current_stream().synchronize()
        self.d2h_bytes = nbytes
        return out


class ScenePipeline:
    """`depth` batches in flight on one GPU, one CUDA stream per slot, the whole forward of a slot captured ONCE in a
    CUDA graph (fixed shapes: batch_size x num_points), replayed per batch.

    Why: distance-FPS is a serial latency chain that occupies one SM per scene for milliseconds (hard
    part 1); run back to back it idles ~130 SMs.  With several batches in flight the FPS of batch k+1 runs on its 16 SMs
    while the tensor-core / gather kernels of batch k use the others, and the tails of small kernels overlap (measured at the
    end of round 1: depth 2 -> 1380, 4 -> 1450, 6 -> 1650, 8 -> 1690, 12 -> 1710 scenes/s); the graph removes the host launch
    cost of the ~270 kernels of a step.  The persistent tensor-core kernels are told to leave one SM per in-flight scene free
    (`pdab_set_persistent_ctas`) so their grid never queues behind an FPS CTA.

    Results are identical to `SceneRunner.infer` (same kernels, same order per batch); tests check it.
    """

    def __init__(self, runner: SceneRunner, depth: int = 4, graphs: bool = True, warm_points: torch.Tensor = None,
                 reserve_sms: int = None):
        from types import SimpleNamespace
        from . import _lib
        from .synthetic import make_batch
        self.runner, self.depth, self.graphs = runner, depth, graphs
        dev, B, N = runner.device, runner.batch_size, runner.num_points
        self.model = runner.model
        self.launches_per_step = 0
        # N > 16384: a scene's FPS runs on a cluster of CTAs; pipelined, the smallest cluster that holds the scene (16384
        # points per CTA) keeps the chain on few SMs beside the other batches' kernels instead of taking the whole GPU
        fps_ctas = -(-N // 16384)
        self._fps_cluster = 16
        if depth > 1 and fps_ctas > 1:
            cl = 1
            while cl < fps_ctas:
                cl *= 2
            fps_ctas = cl
            self._fps_cluster = min(16, cl)
        if reserve_sms is None:  # one SM per FPS CTA of ONE in-flight FPS kernel (deeper pipelines rarely overlap two)
            reserve_sms = B * fps_ctas if depth > 1 and B * fps_ctas < 100 else 0
        self.reserve_sms = reserve_sms
        self._ctas = torch.cuda.get_device_properties(dev).multi_processor_count - reserve_sms
        if warm_points is None:
            warm_points = make_batch(B, N, runner.cfg.POINT_CLOUD_RANGE)["points"]
        self.slots = []
        with torch.cuda.device(dev), torch.no_grad(), self._launch_policy():
            for _ in range(depth):
                s = SimpleNamespace(stream=torch.cuda.Stream(dev), inp=torch.empty(B * N, 5, device=dev),
                                    pinned_in=torch.empty(B * N, 5).pin_memory(), graph=None, out=None,
                                    pinned_out=None, done=torch.cuda.Event(), busy=False, tag=None)
                s.inp.copy_(warm_points.to(dev))
                torch.cuda.synchronize(dev)
                with torch.cuda.stream(s.stream):
                    for _ in range(2):  # builds the packed weights / folded parameters, primes cuBLAS and cuDNN
                        s.out = self._forward(s.inp)
                s.stream.synchronize()
                if graphs:
                    before = sum(_lib.launch_counts.values())
                    s.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(s.graph, stream=s.stream):
                        s.out = self._forward(s.inp)
                    self.launches_per_step = sum(_lib.launch_counts.values()) - before
                s.pinned_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in s.out.items()}
                self.slots.append(s)
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.slots[0].pinned_out.values())
        self.h2d_bytes = B * N * 5 * 4

    def _launch_policy(self):
        """The library's thread-local launch policy for THIS pipeline's launches (grid of the persistent tensor-core kernels,
        FPS cluster cap), restored to the defaults on exit: other runners / direct op calls of the process are not affected,
        and a captured CUDA graph keeps the shapes it was captured with."""
        import contextlib
        from . import _lib

        @contextlib.contextmanager
        def scope():
            lib = _lib.lib()
            _lib.check("pdab_set_persistent_ctas", lib.pdab_set_persistent_ctas(self._ctas))
            _lib.check("pdab_set_fps_max_cluster", lib.pdab_set_fps_max_cluster(self._fps_cluster))
            try:
                yield
            finally:
                lib.pdab_set_persistent_ctas(0)
                lib.pdab_set_fps_max_cluster(16)
        return scope()

    def _forward(self, points_dev):
        prev = self.model.output_padded
        self.model.output_padded = True
        try:
            return self.model({"batch_size": self.runner.batch_size, "points": points_dev})
        finally:
            self.model.output_padded = prev

    def _launch(self, s, src, to_host):
        from . import _lib
        with torch.cuda.stream(s.stream):
            s.inp.copy_(src, non_blocking=True)
            if self.graphs:
                s.graph.replay()
            else:
                with self._launch_policy():
                    s.out = self._forward(s.inp)
            if to_host:
                for k, v in s.out.items():
                    s.pinned_out[k].copy_(v, non_blocking=True)
            s.done.record(s.stream)
        s.busy = True

    def _collect(self, s, to_host):
        s.done.synchronize()
        s.busy = False
        if not to_host:
            return None
        return self.model.unpack_padded({k: v.clone() for k, v in s.pinned_out.items()})

    @torch.no_grad()
    def run(self, host_batches, start_event: torch.cuda.Event = None) -> List[List[dict]]:
        """Host (B*N,5) batches -> per-batch lists of host pred_dicts; H2D, forward and D2H of different batches overlap.
        `start_event` (recorded by the caller) is waited on by every slot stream before its first work."""
        return self._run(host_batches, True, start_event)

    @torch.no_grad()
    def run_device(self, dev_batches, start_event: torch.cuda.Event = None):
        """Batches already resident in HBM; predictions stay on the device (overwritten per slot) — throughput leg."""
        return self._run(dev_batches, False, start_event)

    def _run(self, batches, to_host, start_event):
        results = [None] * len(batches)
        with torch.cuda.device(self.runner.device):
            if start_event is not None:
                for s in self.slots:
                    s.stream.wait_event(start_event)
            for k, b in enumerate(batches):
                s = self.slots[k % self.depth]
                if s.busy:
                    results[s.tag] = self._collect(s, to_host)
                src = b
                if to_host and not b.is_pinned():
                    s.pinned_in.copy_(b)
                    src = s.pinned_in
                s.tag = k
                self._launch(s, src, to_host)
            for s in self.slots:
                if s.busy:
                    results[s.tag] = self._collect(s, to_host)
        return results

    def join_event(self) -> torch.cuda.Event:
        """An event on the current stream that fires after everything enqueued on the slot streams."""
        cur = torch.cuda.current_stream(self.runner.device)
        for s in self.slots:
            cur.wait_event(s.done)
        e = torch.cuda.Event(enable_timing=True)
        e.record(cur)
        return e
