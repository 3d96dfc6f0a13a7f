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

// one step of the chain-walk kinematics (codegen.py: kin[step][chain]); 32 bytes = two 128-bit loads
struct alignas(16) MbKinRec {
  int j;          // joint met at this step of the chain
  float off[3];   // pivot offset in the parent joint frame
  float ax[3];    // joint axis as it points at q = 0 (joint frames are parallel to the base frame at q = 0)
  int store;      // 1: this chain stores the joint's kinematics (0: another chain does, or the chain has ended)
};
