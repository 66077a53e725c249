// Device memory of the library.  Default: cudaMalloc / cudaFree.  With memory reuse on (at3d_set_memory_reuse, or
// AT3D_B200_POOL_GB > 0 in the environment): the stream-ordered pool of the CUDA driver (cudaMallocAsync on the device's
// default memory pool) with a release threshold, so that the buffers of a destroyed state / solver object / per-call arena
// are handed to the next one without a round trip to the OS.  An optimisation loop creates and destroys a render state and
// its derivative tables every evaluation (at3d/medium.py:1813-1831 rebuilds the solvers): with cudaMalloc / cudaFree that
// is ~150 ms of host time per evaluation at BASELINE configs[1].
//   at3d_malloc : the pointer is valid on every stream when the call returns (pooled: the allocation is completed on the
//                 legacy default stream and that stream is synchronised).
//   at3d_free   : cudaFree semantics -- the device is idle when the memory goes back to the pool.
// Off by default only because the library would keep freed memory away from the other allocators of the process (torch's);
// the kernels run at the same speed on recycled and on fresh memory.
#pragma once
#include <cuda_runtime.h>

cudaError_t at3d_pool_alloc(void **p, size_t bytes);
cudaError_t at3d_pool_free(void *p);

template <class T>
inline cudaError_t at3d_malloc(T **p, size_t bytes) { return at3d_pool_alloc((void **)p, bytes); }
inline cudaError_t at3d_free(void *p) { return at3d_pool_free(p); }
