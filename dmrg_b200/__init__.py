"""dmrg_b200 — B200-native matrix-product-state two-site update behind the API of peijunz-archive/DMRG.

Importing the package is free of side effects; the CUDA library (`libmps_b200.so`, built in-tree by
`__graft_entry__.build()`) is loaded on first use and there is no CPU fallback.
"""
__all__ = ['MPS', 'spin', 'Ising', 'ST', 'iDMRG', 'DMRG', 'MPO', 'AKLT', 'batched', 'transmit']
