// device_common.cuh — warp-level primitives shared by every kernel of libb200lda (sm_100a).
//
// Arithmetic contract ("sampling spec"): every fp32 operation that feeds a sampling
// decision goes through the __f*_rn intrinsics below, so nvcc can neither contract a*b+c into an
// FMA nor reassociate; prefix sums are built from the 32-lane Kogge-Stone scan a warp computes
// (per-word prior rows: 32-element tiles chained by a sequential carry; a document's row: each lane
// sums its own slots, one scan over the 32 lane totals). oracle/spec_sampler.c restates exactly
// these orders on the CPU, which is what makes the topic index of every token reproducible bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200lda {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

// One Kogge-Stone step: v += (value of lane l-d) for lanes l >= d. shfl.sync.up returns, besides the
// value, a predicate telling whether the source lane was in range, so the step is SHFL + a
// predicated add.rn.f32 (individually rounded, never contracted): no lane compare, no select.
__device__ __forceinline__ float scan_step(float v, int d) {
  float out;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .f32 y;\n\t"
      "shfl.sync.up.b32 y|p, %1, %2, 0x0, 0xffffffff;\n\t"
      "mov.f32 %0, %1;\n\t"
      "@p add.rn.f32 %0, %1, y;\n\t"
      "}"
      : "=f"(out)
      : "f"(v), "r"(d));
  return out;
}

// Inclusive Kogge-Stone scan over the 32 lanes of a warp (lane l adds lane l-d for d=1,2,4,8,16).
__device__ __forceinline__ float warp_scan_inclusive(float v, int /*lane*/) {
  v = scan_step(v, 1);
  v = scan_step(v, 2);
  v = scan_step(v, 4);
  v = scan_step(v, 8);
  v = scan_step(v, 16);
  return v;
}

// Value of lane l-1, +0 for lane 0 (turns an inclusive warp scan into the exclusive one).
__device__ __forceinline__ float shfl_up1_or_zero(float v) {
  float out;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .f32 y;\n\t"
      "shfl.sync.up.b32 y|p, %1, 1, 0x0, 0xffffffff;\n\t"
      "selp.f32 %0, y, 0f00000000, p;\n\t"
      "}"
      : "=f"(out)
      : "f"(v));
  return out;
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so a token's draw depends only on
// (seed, global token index, sweep, stream) and never on which warp / GPU handles it.
// Replaces cc.mallet.util.Randoms (reference: `new Randoms()` cmu_ron/TrainAndPredict.java:34).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ uint4 token_random(uint64_t seed, uint64_t global_token, uint32_t sweep,
                                              uint32_t stream) {
  return philox4x32_10(make_uint4((uint32_t)global_token, (uint32_t)(global_token >> 32), sweep, stream),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// 24-bit uniform in [0,1): exact in fp32.
__device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * 0x1p-24f; }

// Layout of one word's prior prefix table: level 0 = the K inclusive prefix sums, level i+1 =
// every 32nd entry of level i (last entry clamped), until a level has <= 32 entries. Each level
// is padded to a multiple of 32 floats so every 32-block is one aligned 128-byte line.
struct PriorLayout {
  int nlev;
  int stride;       // floats per word
  int off[5];       // offset of level i inside the word's block
  int size[5];      // valid entries of level i
};

}  // namespace b200lda
