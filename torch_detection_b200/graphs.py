"""CUDA-graph replay of the feature-extraction step.

One ResNet + FPN inference step is 50-60 kernel launches through the C ABI; at batch 1 the GPU finishes them faster
than the host can enqueue them (ctypes + cudaLaunchKernelEx per launch), so small-batch latency is launch-bound.
``GraphedFeatureExtractor`` captures ``neck(backbone(x))`` -- the plans' launches, their metadata reset and the PDL edges
between consecutive kernels -- into ONE ``cudaGraph`` on static input / output buffers and replays it with a single
launch.  Nothing but launch plumbing changes: the replayed kernels are the plans' own, bit for bit.
"""
import torch


class GraphedFeatureExtractor(object):
    """``g = GraphedFeatureExtractor(backbone, neck, example); outs = g(x)``.

    `example`: a CUDA batch with the shape / dtype / strides every later call will have.  The returned tensors are
    the graph's static output buffers: they are overwritten by the next call (clone what must survive)."""

    def __init__(self, backbone, neck=None, example=None, warmup=2):
        if example is None or not example.is_cuda:
            raise NotImplementedError("GraphedFeatureExtractor needs a CUDA example batch: there is no CPU path")
        if backbone.training or (neck is not None and neck.training):
            raise NotImplementedError("CUDA-graph replay covers inference (eval mode) only")
        self.backbone, self.neck = backbone, neck
        self.static_x = example.clone()
        dev = example.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):      # builds the plans, sets the kernels' attributes (not capturable)
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.outs = self._step()

    def _step(self):
        feats = self.backbone(self.static_x)
        return self.neck(feats) if self.neck is not None else feats

    def __call__(self, x):
        if x.shape != self.static_x.shape or x.dtype != self.static_x.dtype:
            raise ValueError("graph captured for %s %s, got %s %s" % (tuple(self.static_x.shape), self.static_x.dtype,
                                                                      tuple(x.shape), x.dtype))
        self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.outs
