// Compile-time switches shared by the generated model tables and the simulator core.
// The SAME source is compiled (a) by nvcc for sm_100a -- the product -- and (b) by g++ as a lane-loop
// emulation used ONLY by tests/emu (debugging the kernel source on a box without a GPU).
#pragma once
#ifdef __CUDACC__
#define MB_HD __device__ __forceinline__
#define MB_TABLE static __device__ const
#define MB_CTABLE static __constant__ const
#else
#define MB_HD inline
#define MB_TABLE static const
#define MB_CTABLE static const
#endif
