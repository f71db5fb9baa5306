"""`onnxruntime.InferenceSession`-shaped front end of the B200 forecast step.

The reference's ONNX scripts drive the 24 h model through

    output, output_surface = ort_session.run(None, {'input': input_24, 'input_surface': input_surface_24})

(`inference/inference_singleOutput.py:146-147`, `inference_multiOutput.py:177-178`, `inference_iterative.py:140-141`):
numpy float32 fields in physical units, `input [5,13,721,1440]`, `input_surface [4,721,1440]`, the same shapes back,
and they chain forecasts by feeding the returned arrays into the next call.  `InferenceSession` below takes that call
unchanged, so those loops run on the B200 path by replacing the `ort.InferenceSession(...)` constructor:

    sess = InferenceSession(model, statistics, statistics_last, maps, const_h)      # model: models.pangu_model.PanguModel on cuda
    output, output_surface = sess.run(None, {'input': x, 'input_surface': xs})

What happens per call: pinned-host staging -> H2D -> one CUDA-graph replay of the forward with the de-normalisation
(`normBackData`, era5_data/utils_data.py:540-546) fused into the patch-recover scatter -> D2H -> numpy.

Chained calls (`chain_on_device=True`, the default): when a call is fed exactly the array OBJECTS the previous call
returned (the chained-forecast loops above), the step starts from the device copy of that state instead: no H2D.  For
this to be bit-identical to re-uploading, the returned arrays must still hold what was returned, so they are handed
out READ-ONLY (`flags.writeable = False`): an in-place edit between calls (clipping humidity, masking, bias
correction) raises numpy's "assignment destination is read-only" instead of being silently dropped; edit a copy
(`x = x.copy()` or `np.clip(x, ...)` without `out=`) and pass that -- a different object is uploaded like any fresh
input.  `chain_on_device=False` gives plain ORT behaviour: writeable arrays, every call uploads its feeds.
There is no CPU path: a model that is not on a CUDA device raises `PanguError`.
"""
import numpy as np
import torch

from .abi import PanguError
from .rollout import Rollout

_UPPER, _SURFACE = (5, 13, 721, 1440), (4, 721, 1440)


class _NodeArg:
    """The three attributes ORT callers read from `get_inputs()` / `get_outputs()` entries."""

    def __init__(self, name, shape):
        self.name, self.shape, self.type = name, list(shape), "tensor(float)"


class InferenceSession:
    def __init__(self, model, statistics, statistics_last, maps, const_h, graph=True, copy_outputs=True,
                 chain_on_device=True):
        """`copy_outputs=False` returns views of the session's pinned host buffers (valid until the next `run`) and saves
        two 287 MB host copies per call; the default hands out fresh arrays, like ORT.  `chain_on_device`: see the module
        docstring (read-only outputs + no re-upload when they are fed back)."""
        self._ro = Rollout(model, statistics, statistics_last, maps, const_h, graph=graph)
        self._copy = copy_outputs
        self._chain = bool(chain_on_device)
        self._host_in = None             # pinned staging, allocated on first use
        self._host_out = None
        self._last = None                # (ndarray, ndarray) returned by the previous run, for the chained fast path
        self._state = None               # device tensors holding the same values
        self.h2d_bytes = 0               # bytes copied host -> device by the last run (0 on the chained path)

    # -- onnxruntime surface --------------------------------------------------------------------------------------
    def get_inputs(self):
        return [_NodeArg("input", _UPPER), _NodeArg("input_surface", _SURFACE)]

    def get_outputs(self):
        return [_NodeArg("output", _UPPER), _NodeArg("output_surface", _SURFACE)]

    def get_providers(self):
        return ["PanguB200ExecutionProvider"]

    def run(self, output_names, input_feed, run_options=None):
        if set(input_feed) != {"input", "input_surface"}:
            raise PanguError("InferenceSession.run: feeds must be exactly {'input', 'input_surface'}, got %s" % sorted(input_feed))
        x, xs = input_feed["input"], input_feed["input_surface"]
        dev = self._ro.dev
        chained = (self._chain and self._last is not None and x is self._last[0] and xs is self._last[1]
                   and not x.flags.writeable and not xs.flags.writeable)
        if chained:
            inp, inp_s = self._state
            self.h2d_bytes = 0
        else:
            x, xs = self._as_f32(x, _UPPER, "input"), self._as_f32(xs, _SURFACE, "input_surface")
            if self._host_in is None:
                self._host_in = (torch.empty(_UPPER, dtype=torch.float32).pin_memory(),
                                 torch.empty(_SURFACE, dtype=torch.float32).pin_memory())
            self._host_in[0].numpy()[...] = x
            self._host_in[1].numpy()[...] = xs
            inp = self._host_in[0].to(dev, non_blocking=True)
            inp_s = self._host_in[1].to(dev, non_blocking=True)
            self.h2d_bytes = x.nbytes + xs.nbytes
        out, out_s = self._step(inp, inp_s)
        if self._host_out is None:
            self._host_out = (torch.empty(_UPPER, dtype=torch.float32).pin_memory(),
                              torch.empty(_SURFACE, dtype=torch.float32).pin_memory())
        self._host_out[0].copy_(out.reshape(_UPPER), non_blocking=True)
        self._host_out[1].copy_(out_s.reshape(_SURFACE), non_blocking=True)
        # keep the state for a chained call; with graph replay `out` is a static buffer the next replay overwrites
        self._state = (out.reshape(_UPPER).clone(), out_s.reshape(_SURFACE).clone())
        torch.cuda.current_stream(dev).synchronize()
        res = [self._host_out[0].numpy(), self._host_out[1].numpy()]
        if self._copy:
            res = [r.copy() for r in res]
        if self._chain:
            for r in res:
                r.flags.writeable = False         # the device state mirrors these values: keep them what they are
            self._last = (res[0], res[1])
        if output_names:
            by_name = {"output": res[0], "output_surface": res[1]}
            try:
                return [by_name[n] for n in output_names]
            except KeyError as e:
                raise PanguError("InferenceSession.run: unknown output %s" % e) from None
        return res

    # -- internals ------------------------------------------------------------------------------------------------
    @staticmethod
    def _as_f32(a, shape, name):
        a = np.asarray(a)
        if a.dtype != np.float32:
            raise PanguError("InferenceSession.run: '%s' must be float32, got %s" % (name, a.dtype))   # ORT raises too
        if a.shape != shape and a.shape != (1,) + shape:
            raise PanguError("InferenceSession.run: '%s' must have shape %s, got %s" % (name, shape, a.shape))
        return a.reshape(shape)

    def _step(self, inp, inp_s):
        ro = self._ro
        if not ro.use_graph:
            return ro.step(inp, inp_s)
        if ro._graphed is None:
            from .graph import GraphedForward
            ro._graphed = GraphedForward(ro.step, (inp, inp_s))
        ro._graphed.load(inp, inp_s)
        return ro._graphed.replay()
