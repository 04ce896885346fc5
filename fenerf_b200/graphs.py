"""CUDA-graph capture of the no_grad render step.

A render step is ~25 small launches (torch's RNG draws, the mapping network's cuBLAS gemv +
leaky_relu kernels, camera / ray set-up / resample / composite kernels) around two long point-network
launches; at 64x64 the host cannot queue them as fast as the GPU retires them.
``GraphedRender`` captures ``generator(*latents, **metadata)`` once per (batch, metadata) key and replays
it: one ``cudaGraphLaunch`` per step.  The random draws stay torch's (its CUDA generator is graph-safe:
the Philox offset advances per replay), so a graphed step consumes the same stream of draws as an eager
one.  Weight changes are NOT seen by a captured graph (the packed-weight check is host code): re-capture
(``invalidate()``) after an optimizer step / EMA swap -- this is an inference / sampling tool.
"""
import torch


class GraphedRender:
    def __init__(self, generator, example_latents, metadata, method="forward", warmup=3):
        self.generator = generator
        self.metadata = dict(metadata)
        self.method = method
        self.static_in = [torch.empty_like(z) for z in example_latents]
        for s, z in zip(self.static_in, example_latents):
            s.copy_(z)
        self.graph = None
        self.out = None
        self._capture(warmup)

    def _run(self):
        fn = getattr(self.generator, self.method) if self.method != "forward" else self.generator
        return fn(*self.static_in, **self.metadata)

    def _capture(self, warmup):
        dev = self.static_in[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):          # packs the weights, sizes the workspace, fills the table caches
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self._run()

    def invalidate(self, warmup=1):
        self._capture(warmup)

    def __call__(self, *latents):
        """Copies the latents into the captured input buffers (device or pinned-host tensors) and replays.
        Returns the captured output tensors -- overwritten by the next call."""
        for s, z in zip(self.static_in, latents):
            s.copy_(z, non_blocking=True)
        self.graph.replay()
        return self.out
