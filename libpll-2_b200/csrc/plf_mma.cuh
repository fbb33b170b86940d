/*
 * plf_mma.cuh -- FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) helpers shared by
 * the 20-state kernels (plf_partials_aa_mma.cu, plf_edge_aa.cu).
 */
#pragma once
#include "plf_device.cuh"

#define AAM_THREADS 256
#define AAM_TAB_STRIDE 22 /* doubles per tip-table row (16-byte aligned rows, codes spread over banks) */
#define AAM_FRAGS 15      /* 3 n-tiles x 5 k-tiles */

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void stg_v2(double * p, double x, double y)
{
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}

/* state a lane's slot q of k-tile kt stands for */
__device__ __forceinline__ int aam_state(int kt, int q) { return kt < 4 ? (kt >> 1) * 8 + 2 * q + (kt & 1) : 16 + q; }

