// Block-phase execution helpers.
//
// Kernels in this library are written as "block programs": a sequence of phases,
// each phase a per-thread body, with a block-wide barrier between phases and all
// cross-phase state in shared memory.  Under nvcc a phase body runs once for
// tid = threadIdx.x followed by __syncthreads(); under a plain host compiler the
// same source runs the body for tid = 0..nthreads-1 sequentially, which lets the
// index / twiddle / bucket logic of every kernel be checked against the oracle on
// a CPU-only box (tests/emu).  The host build is a test aid only: the shipped
// library contains no host implementation of any kernel.
#pragma once
#include "ptx_chain.cuh"

#if defined(__CUDACC__)
#define ZK_PHASE_BEGIN(tid, nthreads) { const uint32_t tid = threadIdx.x; (void)(nthreads);
#define ZK_PHASE_END } __syncthreads();
#define ZK_GLOBAL __global__
#else
#define ZK_PHASE_BEGIN(tid, nthreads) for (uint32_t tid = 0; tid < (nthreads); ++tid) {
#define ZK_PHASE_END }
#define ZK_GLOBAL
#endif
