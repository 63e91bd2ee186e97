// cpfft_b200: the per-voxel material code (kin.cuh, mm01.cuh, mm10.cuh, update.cuh) is written
// once and compiled twice: by nvcc for sm_100a (the product) and by plain g++ for the host
// (tests/native/material_host.cpp), where the `-m "not gpu"` suite runs the very same source
// voxel by voxel against the CPU oracle.  The host build is test infrastructure only; nothing
// in the product path links or calls it.  These macros are the whole difference between the two.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define CPF_DI __device__ __forceinline__
#define CPF_DNOINLINE static __device__ __noinline__
#define CPF_LDG(p) __ldg(p)
// two consecutive doubles with one 16-byte load (p 16-byte aligned)
#define CPF_LDG2(p, a, b) do { const double2 v2_ = __ldg(reinterpret_cast<const double2*>(p)); (a) = v2_.x; (b) = v2_.y; } while (0)
#define CPF_ANY_SYNC(mask, pred) __any_sync(mask, pred)
#define CPF_ACTIVEMASK() __activemask()
#define CPF_ATOMIC_INC(p) atomicAdd(p, 1)
#define CPF_D2LL(d) __double_as_longlong(d)
#define CPF_LL2D(l) __longlong_as_double(l)
#else
#include <cmath>
#include <cstring>
#include <cstdint>
using std::fabs; using std::sqrt; using std::fmax; using std::isnan;
#define CPF_DI inline
#define CPF_DNOINLINE static
#define CPF_LDG(p) (*(p))
#define CPF_LDG2(p, a, b) do { (a) = (p)[0]; (b) = (p)[1]; } while (0)
#define CPF_ANY_SYNC(mask, pred) (pred)          // one "lane" per call on the host
#define CPF_ACTIVEMASK() 1u
#define CPF_ATOMIC_INC(p) __atomic_fetch_add(p, 1, __ATOMIC_RELAXED)
static inline long long CPF_D2LL(double d) { long long l; std::memcpy(&l, &d, 8); return l; }
static inline double CPF_LL2D(long long l) { double d; std::memcpy(&d, &l, 8); return d; }
#endif
