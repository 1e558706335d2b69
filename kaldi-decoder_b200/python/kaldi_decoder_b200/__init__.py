"""ctypes binding of the C ABI (capi), synthetic workloads (synth) and multi-GPU helpers (parallel) of the B200 decoder."""
