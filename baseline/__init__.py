"""Reference-side driver code: runs the UNMODIFIED reference (eth-easl/Scratchpad) attention path -- its
FlashInferAttnBackend over the installed FlashInfer, and its in-tree Triton kernels -- on the same tensors the
cuda_b200 backend sees.  Comparator and parity-pin infrastructure only: nothing under ``scratchpad_b200/`` imports
this package, and nothing here is on the product path."""
