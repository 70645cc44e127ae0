// engine_mma.cuh -- warp-level tensor-core helpers shared by the wide-state tile kernels (engine_wide.cuh) and the
// weight-gradient reduction of the width-4 backward (engine_row4.cuh).  Included inside namespace eng.
#pragma once

// x = hi + lo for the 3xTF32 products.  The tensor core reads only the top 19 bits of a .tf32 operand (the low 13
// mantissa bits are ignored), so x itself serves as hi = trunc_tf32(x) and lo = x - trunc_tf32(x) is exact in fp32:
// two instructions per element (LOP3 + FADD).  cvt.rna.tf32.f32 is a 4-instruction sequence with Inf / NaN
// handling on sm_100a; with it the splits were 40 % of all executed instructions (profiles/README.md).
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x);
    lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
// c (16x8, fp32) += a (16x8, row) * b (8x8, col); fragment layouts of PTX mma.m16n8k8.tf32:
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);  b0 (k=t, n=g)  b1 (k=t+4, n=g);
//   c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)      with g = lane / 4, t = lane % 4
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// d = a * b (zero accumulator input)
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
// 3xTF32 steps.  The tensor core adds into its fp32 accumulator with truncation, which biases a long chain of
// accumulations (measured: 5e-4 relative on a bias gradient after 120 chained mma); so the large terms hi*hi start
// from a zero accumulator, at most two k-steps are chained, and the result is added to `acc` by an ordinary
// round-to-nearest FADD; the small terms (2^-11 of the large ones) chain in their own accumulator `sm`, added once
// at the end.
struct SplitB { uint32_t h0, l0, h1, l1; };
__device__ __forceinline__ SplitB split_b(float b0, float b1) {
    SplitB r;
    tf32_split(b0, r.h0, r.l0);
    tf32_split(b1, r.h1, r.l1);
    return r;
}
// one k-step
__device__ __forceinline__ void mma_3xtf32(float (&acc)[4], float (&sm)[4], const uint32_t (&ah)[4],
                                           const uint32_t (&al)[4], const SplitB& b) {
    mma_tf32(sm, al, b.h0, b.h1);
    mma_tf32(sm, ah, b.l0, b.l1);
    float t[4];
    mma_tf32_zero(t, ah, b.h0, b.h1);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += t[e];
}
// two k-steps (a / b and a2 / b2)
__device__ __forceinline__ void mma_3xtf32_pair(float (&acc)[4], float (&sm)[4], const uint32_t (&ah)[4],
                                                const uint32_t (&al)[4], const SplitB& b, const uint32_t (&ah2)[4],
                                                const uint32_t (&al2)[4], const SplitB& b2) {
    mma_tf32(sm, al, b.h0, b.h1);
    mma_tf32(sm, ah, b.l0, b.l1);
    mma_tf32(sm, al2, b2.h0, b2.h1);
    mma_tf32(sm, ah2, b2.l0, b2.l1);
    float t[4];
    mma_tf32_zero(t, ah, b.h0, b.h1);
    mma_tf32(t, ah2, b2.h0, b2.h1);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += t[e];
}

// two k-steps into a fresh accumulator (small terms first), then one FADD per element: no separate small-term
// accumulator (used where the accumulators of many blocks stay live, dW)
__device__ __forceinline__ void mma_3xtf32_fresh_pair(float (&acc)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                                      const SplitB& b, const uint32_t (&ah2)[4],
                                                      const uint32_t (&al2)[4], const SplitB& b2) {
    float t[4];
    mma_tf32_zero(t, al, b.h0, b.h1);
    mma_tf32(t, ah, b.l0, b.l1);
    mma_tf32(t, al2, b2.h0, b2.h1);
    mma_tf32(t, ah2, b2.l0, b2.l1);
    mma_tf32(t, ah, b.h0, b.h1);
    mma_tf32(t, ah2, b2.h0, b2.h1);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += t[e];
}

