"""Python front-end with the reference's ``Interpreter`` interface
(LCE/tflite/python/interpreter.py:6-58, interpreter_base.py:36-95): same constructor
flags and properties, ``predict`` on numpy arrays -- but TRUE batching: the reference
feeds one sample at a time (interpreter_base.py:10-27 yields batch-1 inputs because
the converter pins the batch to 1, LCE/mlir/transforms/set_batch_size.cc:9-22); here
``predict`` resizes the input tensors to the whole (mini-)batch, re-runs every op's
``prepare`` -- what TFLite's ResizeInputTensor + AllocateTensors does -- and runs the
graph once per mini-batch on the device, replaying a captured CUDA graph.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, List, Optional, Union

import numpy as np

from . import host as _host

Data = Union[np.ndarray, List[np.ndarray]]

__all__ = ["Interpreter"]


class Interpreter:
    def __init__(self, flatbuffer_model: bytes, num_threads: int = 1,
                 use_reference_bconv: bool = False, use_indirect_bgemm: bool = False,
                 use_xnnpack: bool = False, batch_size: Optional[int] = None,
                 use_cuda_graph: bool = True, fuse: bool = True):
        """flatbuffer_model: a serialized LCE model (`.tflite` bytes).
        num_threads / use_xnnpack: accepted for interface parity; the device path has
        no CPU threads to configure. use_reference_bconv / use_indirect_bgemm select
        which of the reference's registrations' validation rules apply (all three run
        the same CUDA kernel). batch_size: mini-batch used by predict (default: all).
        fuse: apply the graph-level fusions (residual-block tail into LceBconv2d's epilogue,
        max-pool + blur-pool); the outputs are bit-identical either way."""
        if use_reference_bconv and use_indirect_bgemm:
            import warnings
            warnings.warn("'use_reference_bconv' and `use_indirect_bgemm` are both set to true. "
                          "use_indirect_bgemm==true will have no effect.")
        try:
            self._g = _host.HostGraph.from_tflite(bytes(flatbuffer_model), device_arena=True)
        except _host.HostError as e:
            raise ValueError(f"Could not build the interpreter: {e}") from None
        self.num_threads = num_threads
        self.batch_size = batch_size
        self.fused_nodes_removed = self._g.fuse_all() if fuse else 0
        self._g.allocate_tensors()
        if use_cuda_graph:
            self._g.enable_cuda_graph(True)
        self._cur_batch = None

    # ---- properties, as in interpreter_base.py:40-78 ----
    def _types(self, idx):
        return [np.dtype(self._g.dtype(t)).type for t in idx]

    @property
    def input_types(self) -> list:
        return self._types(self._g.inputs())

    @property
    def input_shapes(self) -> list:
        return [self._g.shape(t) for t in self._g.inputs()]

    @property
    def input_scales(self) -> list:
        return [self._scale(t) for t in self._g.inputs()]

    @property
    def input_zero_points(self) -> list:
        return [self._zp(t) for t in self._g.inputs()]

    @property
    def output_types(self) -> list:
        return self._types(self._g.outputs())

    @property
    def output_shapes(self) -> list:
        return [self._g.shape(t) for t in self._g.outputs()]

    @property
    def output_scales(self) -> list:
        return [self._scale(t) for t in self._g.outputs()]

    @property
    def output_zero_points(self) -> list:
        return [self._zp(t) for t in self._g.outputs()]

    def _scale(self, t):
        s = _host.lib().lce_host_tensor_scale(self._g._g, t)
        return float(s) if s != 0.0 else None

    def _zp(self, t):
        s = _host.lib().lce_host_tensor_scale(self._g._g, t)
        return int(_host.lib().lce_host_tensor_zero_point(self._g._g, t)) if s != 0.0 else None

    # ---- execution ----
    def _set_batch(self, n):
        if n == self._cur_batch:
            return
        for t in self._g.inputs():
            shape = list(self._g.shape(t))
            shape[0] = n
            self._g.resize_input(t, shape)
        self._g.allocate_tensors()
        self._cur_batch = n

    def _run(self, inputs: List[np.ndarray]) -> List[np.ndarray]:
        n = inputs[0].shape[0]
        self._set_batch(n)
        for t, a in zip(self._g.inputs(), inputs):
            want = self._g.shape(t)
            if tuple(a.shape) != tuple(want):
                raise ValueError(f"input shape {a.shape} does not match {want}")
            self._g.write(t, a)
        self._g.invoke()
        return [self._g.read(t) for t in self._g.outputs()]

    def predict(self, x: Union[Data, Iterator[Data]], verbose: int = 0) -> Data:
        """Generates output predictions for the input samples (leading dim = samples)."""
        if isinstance(x, np.ndarray):
            inputs = [x]
        elif isinstance(x, list):
            inputs = list(x)
        elif hasattr(x, "__next__") and hasattr(x, "__iter__"):
            items = list(x)
            if items and isinstance(items[0], np.ndarray):
                inputs = [np.stack(items)]
            else:
                inputs = [np.stack(col) for col in zip(*items)]
        else:
            raise ValueError(
                "Expected either a list of inputs or a Numpy array with implicit initial "
                f"batch dimension or an iterator yielding one of the above. Received: {x}")
        n = inputs[0].shape[0]
        bs = self.batch_size or n
        chunks = []
        for lo in range(0, n, bs):
            chunks.append(self._run([a[lo:lo + bs] for a in inputs]))
        outputs = [np.concatenate(parts) for parts in zip(*chunks)]
        if len(self._g.outputs()) == 1:
            return outputs[0]
        return outputs

    # device-resident API used by bench.py: no host copies
    @property
    def graph(self):
        return self._g

    def close(self):
        self._g.close()
