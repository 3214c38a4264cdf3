"""CUDA-graph replay of the per-image hot sequence.

At one image's worth of RoIs (tens of objects) the native path is launch-bound: 10 head launches + 2 solver launches +
the score stage + NMS, each preceded by Python / cffi glue.  ``GraphedSequence`` captures a callable that only issues
work on the current CUDA stream (every entry point of libmonorun_head / libmonorun_pnp does, and none synchronises)
into one ``torch.cuda.CUDAGraph`` and replays it on static input buffers.
"""
import torch


class GraphedSequence:
    """``fn(*static_inputs) -> tensor | tuple of tensors``; ``__call__(*inputs)`` copies the inputs into the static
    buffers (same shapes / dtypes as the examples), replays the graph and returns the static outputs."""

    def __init__(self, fn, example_inputs, warmup=3):
        self.static_in = [x.clone() if torch.is_tensor(x) else x for x in example_inputs]
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(warmup):          # allocations, lazy library state (redo lists, workspaces) happen here
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
