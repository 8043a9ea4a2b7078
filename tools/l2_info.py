#!/usr/bin/env python
"""Print the L2 configuration of cuda:0 (round-2 input for the tick's box-to-box spread)."""
from cuda import cudart

def ck(r):
    assert r[0] == cudart.cudaError_t.cudaSuccess, r[0]
    return r[1] if len(r) == 2 else r[1:]

ck(cudart.cudaSetDevice(0) + (None,))
p = ck(cudart.cudaGetDeviceProperties(0))
A = cudart.cudaDeviceAttr
info = {
    "name": p.name.decode() if isinstance(p.name, bytes) else str(p.name),
    "l2CacheSize": p.l2CacheSize, "persistingL2CacheMaxSize": p.persistingL2CacheMaxSize,
    "accessPolicyMaxWindowSize": p.accessPolicyMaxWindowSize, "multiProcessorCount": p.multiProcessorCount,
    "memoryBusWidth": p.memoryBusWidth, "memoryClockRate_kHz": ck(cudart.cudaDeviceGetAttribute(A.cudaDevAttrMemoryClockRate, 0)),
    "limitPersistingL2CacheSize": ck(cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitPersistingL2CacheSize)),
    "totalGlobalMem": p.totalGlobalMem,
}
print(info)
