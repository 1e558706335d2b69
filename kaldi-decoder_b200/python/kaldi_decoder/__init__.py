"""Drop-in `kaldi_decoder` package backed by the B200-native decoder.

The four hot-path names of the reference (kaldi-decoder/python/kaldi_decoder/
__init__.py:1-9) keep their import path, signatures and defaults:

    from kaldi_decoder import DecodableCtc, DecodableInterface, FasterDecoder, FasterDecoderOptions

`SimpleDecoder` (beam-only search) runs on the same kernels.
`LatticeSimpleDecoder` and `LatticeSimpleDecoderConfig` are not part of the
accelerated path and are not provided.

Additions: `StdFst`, `StdVectorFst`, `StdConstFst`, `Lattice`, `get_linear_symbol_sequence` (stand-ins for
the kaldifst types the reference's bindings exchange -- kaldifst is a separate
package), `LatticeFasterDecoderConfig` (the options struct only), `BatchFasterDecoder` (many
lanes per call; `decode_async` / `wait` / `get_results` for deferred, overlapped batches),
`DeviceGraph`, `DeviceConfig`.

The extension module needs the CUDA library built for sm_100a and a GPU at run
time; there is no CPU fallback.
"""
try:
    from kaldi_decoder.lib._kaldi_decoder import (  # noqa: F401
        BatchFasterDecoder,
        DecodableCtc,
        DecodableInterface,
        DeviceConfig,
        DeviceGraph,
        FasterDecoder,
        FasterDecoderOptions,
        Lattice,
        LatticeFasterDecoderConfig,
        SimpleDecoder,
        StdConstFst,
        StdFst,
        StdVectorFst,
        clear_graph_cache,
        device_count,
        get_linear_symbol_sequence,
        graph_uploads,
    )
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "kaldi_decoder: the native module kaldi_decoder/lib/_kaldi_decoder is missing or failed "
        "to load (build it with `python kaldi-decoder_b200/build.py`; it links "
        "kaldi-decoder_b200/lib/libkd_b200.so, sm_100a, no CPU fallback): " + str(e)) from e

__version__ = "0.3.0+b200.r2"


def fst_from_kaldifst(fst) -> "StdVectorFst":
    """Converts a kaldifst/OpenFst-python style FST (``start``, ``num_states``,
    ``final(s)``, arc iteration) into this package's StdVectorFst by duck typing."""
    import numpy as np
    n = int(fst.num_states)
    off = [0]
    il, ol, w, ns, fin = [], [], [], [], []
    try:
        import kaldifst  # type: ignore
        arc_iter = lambda s: _iter_kaldifst(kaldifst, fst, s)  # noqa: E731
    except ImportError:
        arc_iter = lambda s: fst.arcs(s)  # noqa: E731
    for s in range(n):
        for a in arc_iter(s):
            if isinstance(a, tuple):
                i, o, ww, nn = a
            else:
                i, o, nn = a.ilabel, a.olabel, a.nextstate
                ww = getattr(a.weight, "value", a.weight)
            il.append(i); ol.append(o); w.append(float(ww)); ns.append(nn)
        off.append(len(il))
        f = fst.final(s)
        fin.append(float(getattr(f, "value", f)))
    return StdVectorFst.from_arrays(n, int(fst.start), np.asarray(off, np.int64),
                                    np.asarray(il, np.int32), np.asarray(ol, np.int32),
                                    np.asarray(w, np.float32), np.asarray(ns, np.int32),
                                    np.asarray(fin, np.float32))


def _iter_kaldifst(kaldifst, fst, s):
    it = kaldifst.ArcIterator(fst, s)
    while not it.done():
        yield it.value
        it.next()


def _cuda_matrices(tensors):
    """(ptrs, rows, cols, producer stream handle) of contiguous float32 [T, V] CUDA arrays."""
    ptrs, rows, cols, stream, device = [], [], None, 0, None
    for t in tensors:
        if hasattr(t, "data_ptr"):  # torch
            if not t.is_cuda or str(t.dtype) != "torch.float32" or not t.is_contiguous() or t.dim() != 2:
                raise ValueError("expected contiguous float32 [T, V] CUDA tensors")
            ptr, shape = int(t.data_ptr()), tuple(t.shape)
            dev = int(t.device.index or 0)
            if not stream:
                import torch
                # the stream the tensors were (or are being) produced on: the search is ordered
                # behind it -- the decoder's own streams do not synchronise with torch's
                stream = int(torch.cuda.current_stream(t.device).cuda_stream)
        elif not hasattr(t, "__cuda_array_interface__") and hasattr(t, "__dlpack__"):
            ptr, shape, dev = _from_dlpack(t)
        else:
            cai = t.__cuda_array_interface__
            if cai["typestr"] not in ("<f4", "=f4") or len(cai["shape"]) != 2 or cai.get("strides"):
                raise ValueError("expected contiguous float32 [T, V] CUDA arrays")
            ptr, shape = int(cai["data"][0]), tuple(cai["shape"])
            dev = None
            if cai.get("stream") not in (None, 0, 1, 2) and not stream:
                stream = int(cai["stream"])
        if dev is not None:
            if device is None:
                device = dev
            elif device != dev:
                raise ValueError("all matrices must live on the same CUDA device")
        if cols is None:
            cols = int(shape[1])
        elif cols != int(shape[1]):
            raise ValueError("all matrices must have the same number of columns")
        ptrs.append(ptr)
        rows.append(int(shape[0]))
    return ptrs, rows, int(cols or 0), stream, device


def _from_dlpack(t):
    """(data pointer, shape, device index) of a DLPack exporter holding a compact float32
    [T, V] CUDA matrix.  The capsule is consumed (its deleter runs here); the exporter `t`
    keeps the memory alive, so `t` has to outlive the decoding call -- as for every other
    accepted array type."""
    import ctypes as C

    class DLDevice(C.Structure):
        _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]

    class DLDataType(C.Structure):
        _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]

    class DLTensor(C.Structure):
        _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int),
                    ("dtype", DLDataType), ("shape", C.POINTER(C.c_int64)),
                    ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]

    class DLManagedTensor(C.Structure):
        pass

    DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p),
                                ("deleter", C.CFUNCTYPE(None, C.POINTER(DLManagedTensor)))]
    cap = t.__dlpack__()
    api = C.pythonapi
    api.PyCapsule_GetPointer.restype = C.c_void_p
    api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]
    mt = C.cast(api.PyCapsule_GetPointer(cap, b"dltensor"), C.POINTER(DLManagedTensor))
    d = mt.contents.dl_tensor
    try:
        kDLCUDA, kDLFloat = 2, 2
        if d.device.device_type != kDLCUDA:
            raise ValueError("expected a CUDA array (DLPack device type kDLCUDA)")
        if (d.dtype.code, d.dtype.bits, d.dtype.lanes) != (kDLFloat, 32, 1) or d.ndim != 2:
            raise ValueError("expected a float32 [T, V] array")
        shape = (int(d.shape[0]), int(d.shape[1]))
        if d.strides and shape[0] > 0 and (int(d.strides[1]) != 1 or int(d.strides[0]) != shape[1]):
            raise ValueError("expected a compact row-major array")
        return int(d.data or 0) + int(d.byte_offset), shape, int(d.device.device_id)
    finally:
        api.PyCapsule_SetName(cap, b"used_dltensor")  # consumed: the capsule must not free it again
        if mt.contents.deleter:
            mt.contents.deleter(mt)


def advance_decoding_cuda(decoder: "BatchFasterDecoder", lanes, tensors, offsets=None,
                          max_num_frames: int = -1, device: int = None) -> None:
    """`BatchFasterDecoder.advance_decoding` for log-probs that already live on the GPU.

    `tensors[i]` is a contiguous float32 `[T_i, V]` CUDA array for lane `lanes[i]`: a torch
    tensor, anything exposing `__cuda_array_interface__` (CuPy, Numba) or `__dlpack__`.  No copy is made and
    the call returns when the frames are decoded, so the arrays only have to outlive the call.

    Stream ordering: the decoder launches on its own non-blocking streams.  For torch tensors
    the current stream of their device is synchronised first, so work enqueued on it (the
    log_softmax that produced the tensors) is complete before the search reads them; arrays
    produced on other streams must be synchronised by the caller.  `device` (the decoder's
    device index), when given, is checked against the tensors' device.
    """
    ptrs, rows, cols, stream, dev = _cuda_matrices(tensors)
    if device is not None and dev is not None and int(device) != dev:
        raise ValueError(f"tensors live on cuda:{dev}, the decoder on cuda:{device}")
    if stream:
        import torch
        torch.cuda.current_stream(dev).synchronize()
    decoder.advance_decoding_ptrs(list(lanes), ptrs, rows, cols,
                                  list(offsets) if offsets is not None else [],
                                  int(max_num_frames), True)


def decode_cuda_async(decoder: "BatchFasterDecoder", lanes, tensors) -> int:
    """Deferred `decode` of CUDA log-prob matrices (see `advance_decoding_cuda` for the accepted
    arrays): InitDecoding + all frames + best-path selection in one kernel launch, ordered
    behind the stream the tensors were produced on (no host synchronisation).  Returns a ticket
    for `decoder.wait(ticket)` / `decoder.get_results(ticket)`; keep the tensors alive until then."""
    ptrs, rows, cols, stream, _ = _cuda_matrices(tensors)
    return decoder.decode_async(list(lanes), ptrs, rows, cols, True, stream)
