"""dftbplus_b200 -- B200 (sm_100a) implementation of DFTB+'s per-SCC-step hot path.

`csrc/` holds the CUDA kernels and the C ABI (include/dftbp_b200.h); `api.HotPath` is the host
mirror of the reference routines on top of it.  Importing `api` needs the built library
(`python -m dftbplus_b200.build`); there is no CPU fallback.
"""
__version__ = "0.1.0"
