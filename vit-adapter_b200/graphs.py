"""Whole-step CUDA graph capture for the adapter path.

At the reference's batch sizes (2 images per GPU for ViT-Adapter-B, 1 for -L) one adapter interaction is ~170 kernels of
a few microseconds each: the GPU work is 0.9 ms and the Python / launch path 2.8 ms (profiles/r1_block_profile_b2_*).
Everything this package launches is capturable - no host synchronisation (MSDeformAttn's shape check is memoised,
deform_inputs is memoised), no allocation outside torch's pool, every kernel on the stream it is handed - so a static-shape
step can be recorded once and replayed with one launch.

    step = GraphedStep(lambda: train_step(static_images, static_labels))   # warms up, then captures
    for batch in loader:
        static_images.copy_(batch.images); static_labels.copy_(batch.labels)
        loss = step()                                                       # replays; returns the captured outputs

`fn` must be a fixed-shape step on tensors that stay alive (inputs are refreshed with copy_), must not synchronise, and if it
steps an optimizer that optimizer must be capturable (e.g. AdamW(..., fused=True, capturable=True)).
"""
import torch


class GraphedStep:
    """Capture `fn` (a fixed-shape step) once and replay it.

    warm_fn         what the eager warm-up iterations run (default: `fn`). A training step usually warms up with the full
                    eager step (zero_grad + forward + backward + optimizer) and captures the same without the zero_grad.
    before_capture  called once between warm-up and capture (e.g. `optimizer.zero_grad(set_to_none=True)`, so that the
                    captured backward allocates the gradients inside the graph's pool and every replay overwrites them).
    After construction `launches` holds the number of kernels of THIS library inside one replay."""

    def __init__(self, fn, warmup=3, stream=None, warm_fn=None, before_capture=None):
        if not torch.cuda.is_available():
            raise RuntimeError('GraphedStep needs a CUDA device (Not implemented on the CPU)')
        from . import _cabi
        self._fn = fn
        side = stream if stream is not None else torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up off the default stream, as torch.cuda.graph requires
            for _ in range(max(1, int(warmup))):
                (warm_fn or fn)()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if before_capture is not None:
            before_capture()
        self.graph = torch.cuda.CUDAGraph()
        c0 = _cabi.launch_count()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()
        self.launches = _cabi.launch_count() - c0

    def __call__(self):
        self.graph.replay()
        return self.outputs

    replay = __call__
