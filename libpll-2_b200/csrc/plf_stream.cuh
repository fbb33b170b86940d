/*
 * plf_stream.cuh -- shared-memory ring helpers of the streaming DNA kernels:
 * mbarrier + 1-D bulk async copies (cp.async.bulk, the non-tensor TMA path;
 * UBLKCP in SASS), 128-bit shared loads, and the 4-state tip lookup table.
 */
#pragma once
#include "plf_device.cuh"

__device__ __forceinline__ unsigned int smem_u32(const void * p)
{
  return (unsigned int)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long * bar, unsigned int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long * bar, unsigned int bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long * bar, unsigned int parity)
{
  unsigned int ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, unsigned int bytes, unsigned long long * bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

/* tip lookup: tab[code][rate][i] = sum of the columns of row i selected by the
 * 4-bit state mask, pairwise order (src/core_partials_avx.c:1336-1395) */
__device__ __forceinline__ void build_tip_table(double * tab, const double * __restrict__ matrix, int R)
{
  for (int e = threadIdx.x; e < 64 * R; e += blockDim.x)
  {
    const int code = e / (4 * R), r = (e >> 2) % R, i = e & 3;
    tab[e] = masked_sum4(matrix + r * 16 + i * 4, code);
  }
}

__device__ __forceinline__ dbl4 lds_dbl4(const double * p)
{
  const double2 a = *reinterpret_cast<const double2 *>(p);
  const double2 b = *reinterpret_cast<const double2 *>(p + 2);
  return dbl4{a.x, a.y, b.x, b.y};
}

