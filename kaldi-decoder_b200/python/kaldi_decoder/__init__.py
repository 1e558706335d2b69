"""Drop-in `kaldi_decoder` package backed by the B200-native decoder.

The four hot-path names of the reference (kaldi-decoder/python/kaldi_decoder/
__init__.py:1-9) keep their import path, signatures and defaults:

    from kaldi_decoder import DecodableCtc, DecodableInterface, FasterDecoder, FasterDecoderOptions

`SimpleDecoder` (beam-only search) runs on the same kernels.
`LatticeSimpleDecoder` and `LatticeSimpleDecoderConfig` are not part of the
accelerated path and are not provided.

Additions: `StdVectorFst`, `Lattice`, `get_linear_symbol_sequence` (stand-ins for
the kaldifst types the reference's bindings exchange -- kaldifst is a separate
package), `BatchFasterDecoder`, `DeviceGraph`, `DeviceConfig`.

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
        SimpleDecoder,
        StdVectorFst,
        device_count,
        get_linear_symbol_sequence,
    )
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "kaldi_decoder: the native module kaldi_decoder/lib/_kaldi_decoder is missing or failed "
        "to load (build it with `python kaldi-decoder_b200/build.py`; it links "
        "kaldi-decoder_b200/lib/libkd_b200.so, sm_100a, no CPU fallback): " + str(e)) from e

__version__ = "0.3.0+b200.r1"


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


def advance_decoding_cuda(decoder: "BatchFasterDecoder", lanes, tensors, offsets=None,
                          max_num_frames: int = -1) -> None:
    """`BatchFasterDecoder.advance_decoding` for log-probs that already live on the GPU.

    `tensors[i]` is a contiguous float32 `[T_i, V]` CUDA array for lane `lanes[i]`: a torch
    tensor, or anything exposing `__cuda_array_interface__` (CuPy, Numba).  No copy is made and
    the call returns when the frames are decoded, so the arrays only have to outlive the call.
    """
    ptrs, rows, cols = [], [], None
    for t in tensors:
        if hasattr(t, "data_ptr"):  # torch
            if not t.is_cuda or str(t.dtype) != "torch.float32" or not t.is_contiguous() or t.dim() != 2:
                raise ValueError("expected contiguous float32 [T, V] CUDA tensors")
            ptr, shape = int(t.data_ptr()), tuple(t.shape)
        else:
            cai = t.__cuda_array_interface__
            if cai["typestr"] not in ("<f4", "=f4") or len(cai["shape"]) != 2 or cai.get("strides"):
                raise ValueError("expected contiguous float32 [T, V] CUDA arrays")
            ptr, shape = int(cai["data"][0]), tuple(cai["shape"])
        if cols is None:
            cols = int(shape[1])
        elif cols != int(shape[1]):
            raise ValueError("all matrices must have the same number of columns")
        ptrs.append(ptr)
        rows.append(int(shape[0]))
    decoder.advance_decoding_ptrs(list(lanes), ptrs, rows, int(cols or 0),
                                  list(offsets) if offsets is not None else [],
                                  int(max_num_frames), True)
