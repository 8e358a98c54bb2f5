// gsr_tf32.cuh — 3xTF32 helpers for the per-warp mma.sync products of the blend kernels (gsr_blend_fwd.cu, gsr_blend_bwd_mma.cu).
#pragma once
#include "gsr_common.cuh"

namespace gsr {

// x = hi + lo with hi a TF32 value.  cvt.rna.tf32.f32 is emulated on sm_100a (VIADD, FSETP, SEL, LOP3), so the rounding is done
// on the bit pattern directly: adding half a TF32 ulp and clearing the low 13 bits rounds to nearest (ties away); Inf
// becomes NaN, which is where such inputs end up anyway.  lo = x - hi is exact; the tensor core reads its top 19 bits.
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo)
{
	hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
	lo = __float_as_uint(x - __uint_as_float(hi));
}
// truncating variant for operands kept as raw fp32 in registers: hi = x with the low 13 bits cleared (1 op), lo = x - hi
__device__ __forceinline__ void tf32_split_trunc(float x, uint32_t &hi, uint32_t &lo)
{
	hi = __float_as_uint(x) & 0xffffe000u;
	lo = __float_as_uint(x - __uint_as_float(hi));
}
// D += A B, m16n8k8, A row-major (a0: row g col t, a1: row g+8 col t, a2: row g col t+4, a3: row g+8 col t+4),
// B column-major (b0: row t col g, b1: row t+4 col g), C/D (c0, c1: row g cols 2t, 2t+1; c2, c3: row g+8), g = lane >> 2, t = lane & 3
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
	             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

} // namespace gsr
