"""gmc_b200 -- host side of the B200-native GCN max-cut hot path.

    _lib      ctypes binding of lib/libgcnmaxcut.so (C ABI: include/gcnmaxcut.h)
    graph     CSRGraph (DGLGraph stand-in in dataset tuples), GraphBatch (block-diagonal device batch)
    ops       one wrapper per gmc_* entry point, torch tensors in, current stream
    model     GraphConv / GCNSoftmax modules (forward and backward on the library)
    engine    GCNEngine: the fused train / eval step (no autograd)
    optim     FusedAdam (torch.optim.Adam-compatible state)
    synth     vectorised random regular graph batches (BASELINE configs 3-5)
    dist      data-parallel sharding helpers (NCCL all-reduce of the weight gradients)

Importing this package never touches CUDA; the first compute call does, and raises GmcError when
no device or no built library is present (there is no CPU fallback by design).
"""
__version__ = "0.1.0"
