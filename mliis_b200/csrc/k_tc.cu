// Dense contractions on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 issued by one
// thread, operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle, K-major),
// fp32 accumulator in TMEM, read back with tcgen05.ld for the epilogue (bias, optional accumulate).
//
//   C[m, n] = sum_{tap, c} A[pixel(m) + offset(tap), c] * Wt[n][tap][c]  (+ bias[n]) (+ C[m, n])
//
// * 3x3 (dilated) convolutions are implicit GEMMs: for every filter tap the producer issues one 4-D TMA
//   box {32 channels, W, BH rows, 1 image} whose start coordinate is shifted by the tap; out-of-bounds
//   rows/columns are zero-filled by the TMA unit, which IS TensorFlow's SAME zero padding - no im2col
//   buffer, no halo logic, no predicates in the kernel.
// * 1x1 convolutions / plain GEMMs are the taps == 1 case with a 2-D map {K, M} and a {32, 128} box.
// * dgrad is the same kernel on a gradient tensor and re-laid-out weights (tc_prep_weights).
//
// Kernels in this file:
//   tc_conv_kernel   generic implicit GEMM (1x1 layers, and 3x3 layers whose halo'd tile does not fit tc_conv3)
//   tc_conv3_kernel  3x3 / dilated 3x3: one halo'd A box per (filter row, K chunk) serves the three horizontal taps,
//                    two pixel tiles per CTA (forward and dgrad of the decoder convolutions)
//   tc_wgrad_kernel  weight gradients: pixels are the MMA K dimension, both operands MN-major
//                    (SWIZZLE_128B_BASE32B); 3x3 layers share one halo'd A box across the three horizontal taps
//   tc_prep_*        weight operand preparation (re-layout, round-to-nearest TF32, hi/lo planes for 3xTF32)
// 3xTF32 (fp32-class accuracy): a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo; where the hi and lo planes of the B
// operand are adjacent in shared memory one N = 2*BN MMA forms the first and third product together.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2.. = operand transform
// (TF32 rounding / hi-lo split / fused BN+swish+gate prologue, in shared memory, during the main loop) and then
// epilogue (each warp owns the TMEM lane quarter warp_id % 4).  4 such warps in the conv kernels (192 threads),
// 16 in tc_wgrad_kernel, which has to split both operands.
// Reference ops replaced: tf.layers.conv2d of models/efficientlab.py:185-188, :218-224 and their
// Conv2DBackpropInput [TF-ext].
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace mliis {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, M = 128, N from the instruction descriptor, K = 8
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: issue several, then tc_ld_wait() once, then tc_ld_fence() on every destination array
// (an empty asm that "modifies" the registers, so that no use of them can be scheduled above the wait)
__device__ __forceinline__ void tc_ld16_nw(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld_fence(uint32_t (&v)[16]) {
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                    "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (ignored for swizzled K-major, set to 1) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------
struct TcParams {
  int debug;         // MLIIS_TC_DEBUG bits (bottleneck experiments only): 1 skip operand transform, 2 skip MMA issue
  int conv;          // 0: plain [M,K] (2-D map), 1: NHWC taps (4-D map)
  int M;             // plain: number of rows
  int H, W, BH;      // conv: image size, image rows per tile (tile = BH x W pixels <= 128)
  int tiles_per_image;
  int C;             // K per tap
  int taps, dil;
  int N, BN, accumulate;
  int stages;
  int region_bytes;  // shared memory of the stage ring (>= 64 KB: the epilogue reuses it as four 16 KB staging tiles)
  int a_box_bytes;   // bytes one A box delivers
  int split;         // 1: TF32 (operands rounded to nearest)  3: 3xTF32 (hi/lo split, fp32-class accuracy)
  // A-operand prologue applied by the transform warps (plain mode): a <- swish(pa[k]*a + pb[k]) * gate[img][k]
  const float* pa; const float* pb; const float* gate;
  int HW;
};

constexpr int kTcXformWarps = 8;                       // operand transform during the main loop, then epilogue
constexpr int kTcXformThreads = 32 * kTcXformWarps;    // 256
constexpr int kTcThreads = 64 + kTcXformThreads;       // + TMA producer warp + MMA issuer warp
constexpr int kC3Threads = 192;                        // tc_conv3_kernel: 4 transform / epilogue warps
constexpr int kABytes = 128 * 128;   // 128 rows x 32 fp32

// round to nearest TF32 (ties away from zero, same result as cvt.rna.tf32.f32) with two full-rate integer ops
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 rn_tf32_4(float4 v) { return f4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One arrival per WARP: every lane has fenced its own shared-memory writes (fence.proxy.async) before the call, the warp
// barrier orders them before lane 0's arrive (release).  The barrier is initialised with the number of warps: 512
// per-thread arrivals on one barrier word per 600-clk wgrad stage were a serial cost of their own.
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// swish: x * 1/(1 + 2^(-x*log2 e)) with the MUFU reciprocal + one Newton step; the transform warps are ALU-bound
__device__ __forceinline__ float swish_fast(float x) { return x * sigmoid_f(x); }      // common.cuh: MUFU rcp + Newton
__device__ __forceinline__ float4 swish_fast4(float4 v) {
  return f4(swish_fast(v.x), swish_fast(v.y), swish_fast(v.z), swish_fast(v.w));
}
// barrier ids are immediates so that ptxas reserves only the barriers actually used
__device__ __forceinline__ void named_bar_sync_half(int half) {
  if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
// TMA stores of one shared-memory tile (dense [rows][128 B], SWIZZLE_128B) into a global tensor; rows / columns
// outside the tensor are clipped by the TMA unit.  `add`: out += tile (element-wise reduction in L2; every output
// element is produced by exactly one CTA, so the result is deterministic).
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Shared memory: [stage ring: S x {A (hi) 16 KB | A lo 16 KB if split==3 | B hi BN*128 | B lo BN*128 if split==3}]
//                [barriers][prologue coefficients: pa | pb | gate(image 0) | gate(image 1), Kp floats each]
// Generic 1x1 / small 3x3 implicit GEMM, one 128-row tile per CTA.  Round-2 changes (round-1 ncu: 3.6 % of HBM on the
// MBConv project convs, bound by 4 transform warps doing expf + a float division + an integer division + three global
// coefficient loads per float4): 8 transform warps, every per-thread constant hoisted (a thread always owns the same
// channel group and the same 4 rows, so the image index of its rows is computed once), coefficients and gates staged
// in shared memory once per CTA, MUFU reciprocal; shared memory sized for 2 CTAs per SM where the stage is small; the
// epilogue goes TMEM -> registers -> swizzled shared-memory tile -> one TMA store per 32 output columns (coalesced
// 128-byte rows instead of 32 scattered 16-byte stores per warp instruction).
__global__ void __launch_bounds__(kTcThreads, 2)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, TcParams p, long long zs) {
  extern __shared__ uint8_t smem_raw[];
  if (p.debug & 16) return;   // experiment: cost of everything except the tensor-core kernels
  const int slot = blockIdx.z;      // task-batched launch: every tensor map has the slot as its outermost dimension
  { const size_t zo = (size_t)slot * zs; bias = zp(bias, zo); p.pa = zp(p.pa, zo); p.pb = zp(p.pb, zo); p.gate = zp(p.gate, zo); }
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;              // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  const bool x3 = p.split == 3;
  const bool xform = x3 || p.pa != nullptr;     // single-pass TF32 without a prologue: operands go TMA -> MMA directly
  const int b_bytes = p.BN * 128;
  const int a_off_lo = kABytes;
  const int b_off = x3 ? 2 * kABytes : kABytes;
  const int stage_bytes = b_off + (x3 ? 2 : 1) * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.region_bytes);
  const uint32_t bar0 = base + (uint32_t)p.region_bytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                      // TMA -> transform warps
  auto empty_bar = [&](int s) { return bar0 + 8u * (p.stages + s); };        // MMA -> TMA
  auto ready_bar = [&](int s) { return bar0 + 8u * (2 * p.stages + s); };    // transform warps -> MMA
  const uint32_t tmem_full_bar = bar0 + 8u * (3 * p.stages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.stages + 1);
  const int Kp = (p.C + 31) & ~31;
  float* coef = reinterpret_cast<float*>(smem + p.region_bytes + ((8 * (3 * p.stages + 2) + 15) & ~15));
  float* s_pa = coef;
  float* s_pb = coef + Kp;
  float* s_g0 = coef + 2 * Kp;
  float* s_g1 = coef + 3 * Kp;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // MLIIS_TC_DEBUG bit 32: phase timestamps (SM clocks since kernel entry) of one CTA, printed by transform thread 0
  const bool prof = (p.debug & 32) && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
  const long long t_entry = prof ? clock64() : 0;
  // wide: hi and lo weight planes are adjacent in the stage, so one N = 2*BN MMA forms a_hi*b_hi | a_hi*b_lo in two
  // accumulator column ranges that the epilogue adds (A is fetched from shared memory twice per k-step, not three times)
  // (short-K layers are latency-bound, not MMA-bound: they keep the narrow accumulator so that more CTAs fit the
  //  512 TMEM columns of an SM)
  const bool wide = x3 && 2 * p.BN <= 256 && !(p.debug & 8) && (p.taps * ((p.C + 31) / 32) >= 4 || 2 * p.BN <= 128);
  uint32_t ncols = 32;
  while ((int)ncols < (wide ? 2 : 1) * p.BN) ncols <<= 1;

  // tile coordinates
  int img = 0, y0 = 0, m0 = 0;
  if (p.conv) {
    img = blockIdx.x / p.tiles_per_image;
    y0 = (blockIdx.x - img * p.tiles_per_image) * p.BH;
  } else {
    m0 = blockIdx.x * 128;
  }
  const int n0 = blockIdx.y * p.BN;
  const int kchunks = (p.C + 31) / 32;
  const int KB = p.taps * kchunks;

  // TMA producer step kb.  The first min(stages, KB) steps are issued by the thread that initialises the barriers,
  // before the block-wide sync: the first global-memory round trip overlaps the TMEM allocation and the sync.
  auto produce = [&](int kb) {
    const int s = kb % p.stages;
    mbar_wait(empty_bar(s), ((kb / p.stages) & 1) ^ 1);
    const int tap = kb / kchunks, kc = kb - tap * kchunks;
    const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + b_off;
    mbar_expect_tx(full_bar(s), (uint32_t)(p.a_box_bytes + (x3 ? 2 : 1) * b_bytes));
    if (p.conv) {
      const int dy = (tap / 3 - 1) * p.dil, dx = (tap % 3 - 1) * p.dil;
      tma_load_5d(sa, &tmA, full_bar(s), kc * 32, dx, y0 + dy, img, slot);
    } else {
      tma_load_3d(sa, &tmA, full_bar(s), kc * 32, m0, slot);
    }
    tma_load_5d(sb, &tmB, full_bar(s), kc * 32, tap, n0, 0, slot);
    if (x3) tma_load_5d(sb + b_bytes, &tmB, full_bar(s), kc * 32, tap, n0, 1, slot);
  };
  const int n_pre = min(p.stages, KB);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(ready_bar(s), kTcXformThreads / 32);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // barrier words (generic proxy) -> TMA (async proxy)
    for (int kb = 0; kb < n_pre; ++kb) produce(kb);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // prologue coefficients -> shared memory (zero beyond C: swish(0*x + 0) = 0 keeps the TMA zero fill).  A tile
  // of 128 rows spans at most two images when HW >= 127; smaller images (tests) read their gates from global memory.
  const int im0 = p.pa ? m0 / p.HW : 0;
  const bool gate_smem = p.gate != nullptr && (min(m0 + 127, p.M - 1) / p.HW) <= im0 + 1;
  if (p.pa) {
    const int im1 = ((im0 + 1) * p.HW < p.M) ? im0 + 1 : im0;
    for (int i = threadIdx.x; i < Kp; i += kTcThreads) {
      const bool in = i < p.C;
      s_pa[i] = in ? p.pa[i] : 0.f;
      s_pb[i] = in ? p.pb[i] : 0.f;
      if (gate_smem) {
        s_g0[i] = in ? p.gate[(size_t)im0 * p.C + i] : 0.f;
        s_g1[i] = in ? p.gate[(size_t)im1 * p.C + i] : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      for (int kb = n_pre; kb < KB; ++kb) produce(kb);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format TF32 [7,10)=[10,13)=2,
      // K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_w = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 2) << 17) | ((128u >> 4) << 24);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % p.stages;
        mbar_wait(xform ? ready_bar(s) : full_bar(s), (kb / p.stages) & 1);
        tc_fence_after();
        const int kc = kb % kchunks;
        const int rem = p.C - kc * 32;
        const int nk = rem >= 32 ? 4 : (rem + 7) / 8;
        const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + b_off;
        const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sb);
        const uint64_t dal = make_kmajor_sw128_desc(sa + a_off_lo), dbl = make_kmajor_sw128_desc(sb + b_bytes);
        for (int k = 0; k < nk && !(p.debug & 2); ++k) {  // advance 32 bytes (8 tf32) along K inside the swizzle atom
          const uint64_t adv = (uint64_t)(2 * k);
          // a*b ~= ah*bh + al*bh + ah*bl   (al*bl ~ 2^-22 relative, dropped)
          if (wide) {
            tc_mma_tf32(tmem_acc, da + adv, db + adv, idesc_w, (kb | k) ? 1u : 0u);
            tc_mma_tf32(tmem_acc, dal + adv, db + adv, idesc, 1u);
          } else {
            tc_mma_tf32(tmem_acc, da + adv, db + adv, idesc, (kb | k) ? 1u : 0u);
            if (x3) {
              tc_mma_tf32(tmem_acc, dal + adv, db + adv, idesc, 1u);
              tc_mma_tf32(tmem_acc, da + adv, dbl + adv, idesc, 1u);
            }
          }
        }
        tc_commit(empty_bar(s));      // frees the stage once these MMAs have read it
      }
      tc_commit(tmem_full_bar);       // accumulator complete
    }
  } else {
    // ---------------- operand transform (during the main loop), then epilogue ----------------
    const int t = threadIdx.x - 64;   // 0..255
    const long long t_sync = prof ? clock64() - t_entry : 0;
    long long t_first = 0, t_xf = 0;
    // float4 i = t + 256*j of the 16 KB A tile: row = i/8 = (t>>3) + 32*j, physical 16-byte chunk = t&7.
    // SWIZZLE_128B: logical chunk = physical chunk XOR (row & 7); 32*j does not touch the low 3 row bits, so a
    // thread owns ONE channel group for the whole kernel and the same four rows in every stage.
    const int rl = t >> 3;
    const int lc4 = (((t & 7) ^ (rl & 7)) << 2);
    int sel[4];
    size_t goff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + rl + 32 * j;
      const int im = (p.gate && m < p.M) ? m / p.HW : im0;
      sel[j] = im - im0;
      goff[j] = (size_t)im * p.C;
    }
    for (int kb = 0; kb < KB && xform; ++kb) {
      const int s = kb % p.stages;
      const int k = (kb % kchunks) * 32 + lc4;
      float4 a4 = f4s(0.f), b4 = f4s(0.f), g4[4];
      if (p.pa) {
        a4 = ld4(s_pa + k);
        b4 = ld4(s_pb + k);
        if (p.gate) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (gate_smem) g4[j] = ld4((sel[j] ? s_g1 : s_g0) + k);
            else g4[j] = k < p.C ? ld4(p.gate + goff[j] + k) : f4s(0.f);
          }
        }
      }
      mbar_wait(full_bar(s), (kb / p.stages) & 1);
      if (prof && kb == 0) t_first = clock64() - t_entry;
      float4* a_hi = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
      float4* a_lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + a_off_lo);
      if (!(p.debug & 1)) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = a_hi[t + kTcXformThreads * j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 x = v[j];
          if (p.pa) {
            x = swish_fast4(affine4(x, a4, b4));
            if (p.gate) x = x * g4[j];
          }
          const float4 h = rn_tf32_4(x);
          a_hi[t + kTcXformThreads * j] = h;
          if (x3) a_lo[t + kTcXformThreads * j] = rn_tf32_4(x - h);
        }
      }
      // generic-proxy writes must be visible to the async proxy (tcgen05.mma reads smem through it)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_warp(ready_bar(s));
    }
    // ---- epilogue: TMEM lane quarter = warp & 3 (hardware rule); the two warps of a quarter split the 32-column
    // chunks (half 0: even chunks, half 1: odd chunks); each half stages through its own 16 KB tile of the ring ----
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;          // accumulator row == TMEM lane == row of the staging tile
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
    if (prof) t_xf = clock64() - t_entry;
    mbar_wait(tmem_full_bar, 0);                // every MMA has completed: the stage ring is free
    tc_fence_after();
    const long long t_acc = prof ? clock64() - t_entry : 0;
    const uint32_t tbase = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const int nchunks = (p.BN + 31) / 32;
    int nst = 0;                                // stores issued by this half: two staging tiles alternate
    for (int ch = half; ch < nchunks; ch += 2, ++nst) {
      const int c = ch * 32, n = n0 + c;
      if (n >= p.N) break;                      // uniform over the half: whole chunk outside the tensor
      uint8_t* stg = smem + (half * 2 + (nst & 1)) * 16384;
      const uint32_t stg_addr = base + (uint32_t)(half * 2 + (nst & 1)) * 16384u;
      // the store that last read THIS tile (two stores ago) is done; the previous one may still be in flight
      if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      named_bar_sync_half(half);
      // two 16-column halves: per half the accumulator (and, for 3xTF32 "wide", its a_hi*b_lo partner range) leaves
      // TMEM with ONE wait; the bias loads are issued first.  Register budget: the kernel must keep 2 CTAs per SM.
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int cc = c + 16 * h2;
        const bool live = cc < p.BN;              // uniform
        float4 add[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) add[q] = (bias && n + 16 * h2 + q * 4 < p.N) ? ld4(bias + n + 16 * h2 + q * 4) : f4s(0.f);
        uint32_t v0[16], w0[16];
        __syncwarp();
        if (live) {
          tc_ld16_nw(tbase + (uint32_t)cc, v0);   // warp-collective: executed by all 32 lanes, converged
          if (wide) tc_ld16_nw(tbase + (uint32_t)(p.BN + cc), w0);
          tc_ld_wait();
        }
        tc_ld_fence(v0); tc_ld_fence(w0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o = f4s(0.f);
          if (live) {
            o = f4(__uint_as_float(v0[q * 4 + 0]), __uint_as_float(v0[q * 4 + 1]), __uint_as_float(v0[q * 4 + 2]),
                   __uint_as_float(v0[q * 4 + 3]));
            if (wide) o = o + f4(__uint_as_float(w0[q * 4 + 0]), __uint_as_float(w0[q * 4 + 1]), __uint_as_float(w0[q * 4 + 2]),
                                 __uint_as_float(w0[q * 4 + 3]));
          }
          *reinterpret_cast<float4*>(stg + r * 128 + (((4 * h2 + q) ^ (r & 7)) << 4)) = o + add[q];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      named_bar_sync_half(half);
      if (leader) {
        if (p.conv) tma_store_4d(&tmC, stg_addr, n, y0 * p.W, img, slot, p.accumulate != 0);
        else tma_store_3d(&tmC, stg_addr, n, m0, slot, p.accumulate != 0);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_read();          // shared memory must stay valid until the TMA unit has read it
    if (prof && t == 0)
      printf("[tc_conv] grid %d x %d  KB %d stages %d BN %d wide %d | after sync %lld | first stage landed %lld | transforms "
             "done %lld | accumulator complete %lld | epilogue done %lld\n", gridDim.x, gridDim.y, KB, p.stages, p.BN,
             (int)wide, t_sync, t_first, t_xf, t_acc, clock64() - t_entry);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(ncols) : "memory");
  }
}

// =================================================================================================
// PERSISTENT pointwise convolution (round 2; VERDICT r1 "kernel furthest below its roofline").  The MBConv 1x1 layers at
// 112x112 / 56x56 have 196-784 row tiles of 128 pixels and K <= 160: in tc_conv_kernel every tile is its own CTA and
// pays 5-7 us of SERIAL fixed work (barrier init, TMEM allocation, first TMA round trip, transform, MMA, TMEM -> smem ->
// TMA-store epilogue) for < 1 us of streaming.  Here one CTA per SM walks the tiles:
//   * the weight operand (all K, hi and lo planes) is loaded ONCE and stays resident in shared memory;
//   * the ring holds A stages only; the producer runs ahead across tile boundaries;
//   * TWO TMEM accumulators: the epilogue warps drain tile i while the transform warps and the MMA work on tile i+1;
//   * roles: warp 0 TMA producer, warp 1 MMA issuer, 8 transform warps, 8 epilogue warps (own staging tiles, outside
//     the ring).  Barriers: b_full | a_full / a_ready / a_empty per stage | acc_full / acc_empty per accumulator.
// Same arithmetic, same order of accumulation per output as tc_conv_kernel (bit-identical results).
// =================================================================================================
struct TcPwParams {
  int M, C, N, BN, n_tiles, KB, SA, split, wide, ncol_acc, ncols_alloc, accumulate;
  int b_res_bytes;      // resident weight operand: KB x planes x BN x 128
  const float* pa; const float* pb; const float* gate;
  int HW;
  int debug;
};
constexpr int kPwXformThreads = 512, kPwEpiThreads = 256;   // 16 transform warps: the fused BN+swish+gate prologue is ALU-bound
constexpr int kPwNJ = 1024 / kPwXformThreads;                // float4 of a 16 KB A tile per transform thread
constexpr int kPwThreads = 64 + kPwXformThreads + kPwEpiThreads;

__global__ void __launch_bounds__(kPwThreads, 1)
tc_pw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, TcPwParams p, long long zs) {
  extern __shared__ uint8_t smem_raw[];
  if (p.debug & 16) return;
  const int slot = blockIdx.z;
  { const size_t zo = (size_t)slot * zs; bias = zp(bias, zo); p.pa = zp(p.pa, zo); p.pb = zp(p.pb, zo); p.gate = zp(p.gate, zo); }
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const bool x3 = p.split == 3;
  const int planes = x3 ? 2 : 1;
  const int b_bytes = p.BN * 128;
  const int a_stage = planes * kABytes;
  // [weights KB x (hi|lo)] [A ring SA x (hi|lo)] [staging 4 x 16 KB] [barriers] [pa | pb]
  const uint32_t b_base = base, a_base = base + (uint32_t)p.b_res_bytes;
  const uint32_t stg_base = a_base + (uint32_t)p.SA * a_stage;
  const uint32_t bar0 = stg_base + 4u * 16384u;
  uint8_t* bars_ptr = smem + p.b_res_bytes + (size_t)p.SA * a_stage + 4 * 16384;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_ready = [&](int s) { return bar0 + 8u * (p.SA + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * p.SA + s); };
  const uint32_t b_full = bar0 + 8u * (3 * p.SA);
  auto acc_full = [&](int b) { return bar0 + 8u * (3 * p.SA + 1 + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (3 * p.SA + 3 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * (3 * p.SA + 5));
  const int Kp = (p.C + 31) & ~31;
  float* s_pa = reinterpret_cast<float*>(bars_ptr + ((8 * (3 * p.SA + 5) + 4 + 15) & ~15));
  float* s_pb = s_pa + Kp;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ncols = (uint32_t)p.ncols_alloc;
  const int KB = p.KB;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < p.SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_ready(s), kPwXformThreads / 32); mbar_init(a_empty(s), 1); }
    mbar_init(b_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), kPwEpiThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // the resident weight operand: every k-block, hi plane then lo plane, back to back
    mbar_expect_tx(b_full, (uint32_t)p.b_res_bytes);
    for (int kb = 0; kb < KB; ++kb)
      for (int pl = 0; pl < planes; ++pl)
        tma_load_5d(b_base + (uint32_t)((kb * planes + pl) * b_bytes), &tmB, b_full, kb * 32, 0, 0, pl, slot);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.pa) {
    for (int i = threadIdx.x; i < Kp; i += kPwThreads) {
      const bool in = i < p.C;
      s_pa[i] = in ? p.pa[i] : 0.f;        // zero beyond C: swish(0*x + 0) = 0 keeps the TMA zero fill
      s_pb[i] = in ? p.pb[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: A stages, across tile boundaries ----------------
      int it = 0;
      for (int tl = 0; tl < n_my; ++tl) {
        const int m0 = ((int)blockIdx.x + tl * (int)gridDim.x) * 128;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % p.SA;
          mbar_wait(a_empty(s), ((it / p.SA) & 1) ^ 1);
          mbar_expect_tx(a_full(s), (uint32_t)kABytes);
          tma_load_3d(a_base + (uint32_t)s * a_stage, &tmA, a_full(s), kb * 32, m0, slot);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_w = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 2) << 17) | ((128u >> 4) << 24);
      mbar_wait(b_full, 0);
      int it = 0;
      for (int tl = 0; tl < n_my; ++tl) {
        const int buf = tl & 1;
        mbar_wait(acc_empty(buf), ((tl >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_acc + (uint32_t)(buf * p.ncol_acc);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % p.SA;
          mbar_wait(a_ready(s), (it / p.SA) & 1);
          tc_fence_after();
          const int rem = p.C - kb * 32;
          const int nk = rem >= 32 ? 4 : (rem + 7) / 8;
          const uint32_t sa = a_base + (uint32_t)s * a_stage, sb = b_base + (uint32_t)(kb * planes * b_bytes);
          const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sb);
          const uint64_t dal = make_kmajor_sw128_desc(sa + kABytes), dbl = make_kmajor_sw128_desc(sb + b_bytes);
          for (int k = 0; k < nk; ++k) {
            const uint64_t adv = (uint64_t)(2 * k);
            if (p.wide) {
              tc_mma_tf32(acc, da + adv, db + adv, idesc_w, (kb | k) ? 1u : 0u);
              tc_mma_tf32(acc, dal + adv, db + adv, idesc, 1u);
            } else {
              tc_mma_tf32(acc, da + adv, db + adv, idesc, (kb | k) ? 1u : 0u);
              if (x3) {
                tc_mma_tf32(acc, dal + adv, db + adv, idesc, 1u);
                tc_mma_tf32(acc, da + adv, dbl + adv, idesc, 1u);
              }
            }
          }
          tc_commit(a_empty(s));
        }
        tc_commit(acc_full(buf));
      }
    }
  } else if (warp < 2 + kPwXformThreads / 32) {
    // ---------------- operand transform (8 warps): TF32 rounding, hi/lo split, fused BN + swish + gate ----------------
    const int t = threadIdx.x - 64;   // float4 i = t + kPwXformThreads*j: row (t>>3) + (kPwXformThreads/8)*j, 16-byte chunk t&7
    const int rl = t >> 3;
    const int lc4 = (((t & 7) ^ (rl & 7)) << 2);
    int it = 0;
    for (int tl = 0; tl < n_my; ++tl) {
      const int m0 = ((int)blockIdx.x + tl * (int)gridDim.x) * 128;
      size_t goff[kPwNJ];
#pragma unroll
      for (int j = 0; j < kPwNJ; ++j) {
        const int m = min(m0 + rl + (kPwXformThreads / 8) * j, p.M - 1);
        goff[j] = p.gate ? (size_t)(m / p.HW) * p.C : 0;
      }
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % p.SA;
        const int k = kb * 32 + lc4;
        float4 a4 = f4s(0.f), b4 = f4s(0.f), g4[kPwNJ];
        if (p.pa) {
          a4 = ld4(s_pa + k);
          b4 = ld4(s_pb + k);
          if (p.gate) {
#pragma unroll
            for (int j = 0; j < kPwNJ; ++j) g4[j] = k < p.C ? ld4(p.gate + goff[j] + k) : f4s(0.f);
          }
        }
        mbar_wait(a_full(s), (it / p.SA) & 1);
        float4* a_hi = reinterpret_cast<float4*>(smem + p.b_res_bytes + (size_t)s * a_stage);
        float4* a_lo = a_hi + kABytes / 16;
        float4 v[kPwNJ];
#pragma unroll
        for (int j = 0; j < kPwNJ; ++j) v[j] = a_hi[t + kPwXformThreads * j];
#pragma unroll
        for (int j = 0; j < kPwNJ; ++j) {
          float4 x = v[j];
          if (p.pa) {
            x = swish_fast4(affine4(x, a4, b4));
            if (p.gate) x = x * g4[j];
          }
          const float4 h = rn_tf32_4(x);
          a_hi[t + kPwXformThreads * j] = h;
          if (x3) a_lo[t + kPwXformThreads * j] = rn_tf32_4(x - h);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_warp(a_ready(s));
      }
    }
  } else {
    // ---------------- epilogue (8 warps): TMEM -> registers (+ bias) -> swizzled smem tile -> TMA store ----------------
    const int ew = warp - (2 + kPwXformThreads / 32);      // 0..7
    const int quarter = warp & 3, half = ew >> 2;          // TMEM lane quarter = warp % 4 (hardware rule)
    const int r = quarter * 32 + lane;
    const bool leader = (ew & 3) == 0 && lane == 0;
    const int nchunks = (p.BN + 31) / 32;
    int nst = 0;                                           // stores issued by this half: staging tile = nst & 1
    for (int tl = 0; tl < n_my; ++tl) {
      const int buf = tl & 1;
      const int m0 = ((int)blockIdx.x + tl * (int)gridDim.x) * 128;
      mbar_wait(acc_full(buf), (tl >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_acc + (uint32_t)(buf * p.ncol_acc) + ((uint32_t)(quarter * 32) << 16);
      for (int ch = half; ch < nchunks; ch += 2) {
        const int c = ch * 32, n = c;
        if (n >= p.N) break;                               // uniform over the half
        const uint32_t stg_off = (uint32_t)((half * 2 + (nst & 1)) * 16384);
        uint8_t* stg = smem + p.b_res_bytes + (size_t)p.SA * a_stage + stg_off;
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this tile
        named_bar_sync_half(half);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int cc = c + 16 * h2;
          const bool live = cc < p.BN;                     // uniform
          float4 add[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) add[q] = (bias && n + 16 * h2 + q * 4 < p.N) ? ld4(bias + n + 16 * h2 + q * 4) : f4s(0.f);
          uint32_t v0[16], w0[16];
          __syncwarp();
          if (live) {
            tc_ld16_nw(tbase + (uint32_t)cc, v0);
            if (p.wide) tc_ld16_nw(tbase + (uint32_t)(p.BN + cc), w0);
            tc_ld_wait();
          }
          tc_ld_fence(v0); tc_ld_fence(w0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o = f4s(0.f);
            if (live) {
              o = f4(__uint_as_float(v0[q * 4 + 0]), __uint_as_float(v0[q * 4 + 1]), __uint_as_float(v0[q * 4 + 2]),
                     __uint_as_float(v0[q * 4 + 3]));
              if (p.wide) o = o + f4(__uint_as_float(w0[q * 4 + 0]), __uint_as_float(w0[q * 4 + 1]),
                                     __uint_as_float(w0[q * 4 + 2]), __uint_as_float(w0[q * 4 + 3]));
            }
            *reinterpret_cast<float4*>(stg + r * 128 + (((4 * h2 + q) ^ (r & 7)) << 4)) = o + add[q];
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync_half(half);
        if (leader) {
          tma_store_3d(&tmC, stg_base + stg_off, n, m0, slot, p.accumulate != 0);
          tma_store_commit();
        }
        ++nst;
      }
      // every tcgen05.ld of this warp on the accumulator has completed (wait::ld above): hand it back to the MMA
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
    }
    if (leader) tma_store_wait_read();          // shared memory must stay valid until the TMA unit has read it
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(ncols) : "memory");
  }
}

// =================================================================================================
// 3x3 (dilated) implicit GEMM, second generation.  Measured limiter of tc_conv_kernel on the big decoder convs: the
// per-SM TMA engine (~one 128-byte row per ~5 clk), not the tensor pipe.  So this kernel moves fewer rows per MMA:
//  * ONE halo'd A box {32 ch, W+2d, MT*BH rows} per (filter row dy, K chunk) serves the three horizontal taps: tap dx
//    reads the same shared-memory slab through a descriptor whose start address is shifted by dx*d rows (128 B each).
//    Measured on B200: the SWIZZLE_128B XOR phase is taken from the absolute shared-memory address bits, so a
//    row-shifted start needs NO base_offset (setting it double-applies the phase and scrambles the K chunks);
//  * MT = 2 pixel tiles per CTA share every weight tile (two TMEM accumulators);
//  * A and B live in separate rings (A: 2-3 slots of 32 KB, B: up to 8 slots), so the pipeline is deeper.
// Tile rows are r = ly*(W+2d) + lx; rows with lx >= W are halo pixels whose outputs are discarded.
// =================================================================================================
// which filter taps read inside the image for an output pixel of border class cls = 3*ry + rx (k_pool.cu)
__device__ __forceinline__ bool c3_tap_valid(int tap, int cls) {
  const int ty = tap / 3, tx = tap - ty * 3, ry = cls / 3, rx = cls - ry * 3;
  return !((ty == 0 && ry == 0) || (ty == 2 && ry == 2) || (tx == 0 && rx == 0) || (tx == 2 && rx == 2));
}

struct TcC3Params {
  int H, W, RW, BH, MT, tiles_per_image, groups_per_image;
  int C, dil, N, BN, ldc, accumulate;
  int SA, SB, split, ncol_acc, ncols_alloc, wide;
  int a_box_bytes, a_slot_bytes, b_plane_bytes;
  int boff;          // experiment knob: 1 = set the descriptor base_offset field for shifted starts, 0 = leave it 0
  int debug;         // MLIIS_TC_DEBUG bits: 1 skip operand transform, 2 skip MMA issue, 4 one tile per CTA, 8 no wide-B
  const float* bias9;   // [B][9 taps][N] per-image tap vectors of the folded pooled branch (pool_taps_kernel) or null
};

__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_off(uint32_t smem_addr) {
  // start address not 1024-byte aligned: base_offset (bits 49..51) = (address >> 7) & 7
  return make_kmajor_sw128_desc(smem_addr) | ((uint64_t)((smem_addr >> 7) & 7) << 49);
}

__global__ void __launch_bounds__(kC3Threads)
tc_conv3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, TcC3Params p, long long zs) {
  extern __shared__ uint8_t smem_raw[];
  if (p.debug & 16) return;   // experiment: cost of everything except the tensor-core kernels
  const int slot = blockIdx.z;
  { const size_t zo = (size_t)slot * zs; bias = zp(bias, zo); p.bias9 = zp(p.bias9, zo); }
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const bool x3 = p.split == 3;
  const int a_stage = (x3 ? 2 : 1) * p.a_slot_bytes;          // [hi][lo]
  const int b_stage = (x3 ? 2 : 1) * p.b_plane_bytes;         // [hi][lo]
  const uint32_t a_base = base, b_base = base + (uint32_t)p.SA * a_stage;
  const uint32_t bar0 = b_base + (uint32_t)p.SB * b_stage;
  uint8_t* bars_ptr = smem + (size_t)p.SA * a_stage + (size_t)p.SB * b_stage;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_ready = [&](int s) { return bar0 + 8u * (p.SA + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * p.SA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (3 * p.SA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (3 * p.SA + p.SB + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (3 * p.SA + 2 * p.SB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * (3 * p.SA + 2 * p.SB + 1));
  float* b9s = reinterpret_cast<float*>(bars_ptr + ((8 * (3 * p.SA + 2 * p.SB + 1) + 4 + 15) & ~15));   // [9 classes][BN] (bias9 only), 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ncols = (uint32_t)p.ncols_alloc;
  // MLIIS_TC_DEBUG bit 32: phase timestamps of CTA 0 (SM clocks since kernel entry), printed by each role's lane 0
  const bool prof = (p.debug & 32) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  const long long t_entry = prof ? clock64() : 0;
  const int img = blockIdx.x / p.groups_per_image;
  const int y0 = (blockIdx.x - img * p.groups_per_image) * p.MT * p.BH;
  const int n0 = blockIdx.y * p.BN;
  const int KC = (p.C + 31) / 32;
  const int NA = 3 * KC;                 // A stages: (dy, kc)
  // producer work item j = 4*ia + w: w == 0 loads the halo'd A box of stage ia = (dy, kc), w = 1..3 the weight tile of
  // tap (dy, dx = w-1).  Items [0, n_pre) are issued by the thread that initialises the barriers, BEFORE the block-wide
  // sync (every slot is free then): the first global-memory round trip overlaps the TMEM allocation and the sync.
  long long w_pa = 0, w_pb = 0;
  auto produce = [&](int j) {
    const int ia = j >> 2, w = j & 3;
    const int dy = ia / KC, kc = ia - dy * KC;
    if (w == 0) {
      const int sa = ia % p.SA;
      const long long t0 = prof ? clock64() : 0;
      mbar_wait(a_empty(sa), ((ia / p.SA) & 1) ^ 1);
      if (prof) w_pa += clock64() - t0;
      mbar_expect_tx(a_full(sa), (uint32_t)p.a_box_bytes);
      tma_load_5d(a_base + (uint32_t)sa * a_stage, &tmA, a_full(sa), kc * 32, -p.dil, y0 + (dy - 1) * p.dil, img, slot);
    } else {
      const int dx = w - 1, ib = 3 * ia + dx, sb = ib % p.SB;
      const long long t0 = prof ? clock64() : 0;
      mbar_wait(b_empty(sb), ((ib / p.SB) & 1) ^ 1);
      if (prof) w_pb += clock64() - t0;
      mbar_expect_tx(b_full(sb), (uint32_t)b_stage);
      const uint32_t dst = b_base + (uint32_t)sb * b_stage;
      tma_load_5d(dst, &tmB, b_full(sb), kc * 32, dy * 3 + dx, n0, 0, slot);
      if (x3) tma_load_5d(dst + p.b_plane_bytes, &tmB, b_full(sb), kc * 32, dy * 3 + dx, n0, 1, slot);
    }
  };
  const int n_pre = min(4 * NA, 1 + min(p.SB, 3));
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < p.SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_ready(s), 128 / 32); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.SB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // barrier words (generic proxy) -> TMA (async proxy)
    for (int j = 0; j < n_pre; ++j) produce(j);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int j = n_pre; j < 4 * NA; ++j) produce(j);
      const long long w_a = w_pa, w_b = w_pb;
      if (prof) printf("[c3 producer] NA %d SA %d SB %d | done issuing at %lld | waited a_empty %lld b_empty %lld\n", NA, p.SA,
                       p.SB, clock64() - t_entry, w_a, w_b);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      long long w_a = 0, w_b = 0, t_first = 0;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      // wide: the hi and lo weight planes sit back to back in the B slot, so one N = 2*BN MMA forms a_hi*b_hi and
      // a_hi*b_lo in adjacent accumulator column ranges (A is fetched from shared memory once for both products)
      const uint32_t idesc_w = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 2) << 17) | ((128u >> 4) << 24);
      int ib = 0;
      for (int ia = 0; ia < NA; ++ia) {
        const int kc = ia % KC;
        const int sa = ia % p.SA;
        long long t0 = prof ? clock64() : 0;
        mbar_wait(x3 ? a_ready(sa) : a_full(sa), (ia / p.SA) & 1);
        if (prof) { const long long t1 = clock64(); if (ia == 0) t_first = t1 - t_entry; else w_a += t1 - t0; }
        const int rem = p.C - kc * 32;
        const int nk = rem >= 32 ? 4 : (rem + 7) / 8;
        const uint32_t a_hi = a_base + (uint32_t)sa * a_stage, a_lo = a_hi + p.a_slot_bytes;
        for (int dx = 0; dx < 3; ++dx, ++ib) {
          const int sb = ib % p.SB;
          t0 = prof ? clock64() : 0;
          mbar_wait(b_full(sb), (ib / p.SB) & 1);
          if (prof) w_b += clock64() - t0;
          tc_fence_after();
          const uint32_t b_hi = b_base + (uint32_t)sb * b_stage, b_lo = b_hi + p.b_plane_bytes;
          const uint64_t db = make_kmajor_sw128_desc(b_hi), dbl = make_kmajor_sw128_desc(b_lo);
          for (int t = 0; t < p.MT; ++t) {
            const uint32_t row_off = (uint32_t)((t * p.BH * p.RW + dx * p.dil) * 128);
            const uint64_t da = p.boff ? make_kmajor_sw128_desc_off(a_hi + row_off) : make_kmajor_sw128_desc(a_hi + row_off);
            const uint64_t dal = p.boff ? make_kmajor_sw128_desc_off(a_lo + row_off) : make_kmajor_sw128_desc(a_lo + row_off);
            const uint32_t acc = tmem_acc + (uint32_t)(t * p.ncol_acc);
            for (int k = 0; k < nk && !(p.debug & 2); ++k) {
              const uint64_t adv = (uint64_t)(2 * k);
              if (p.wide) {
                tc_mma_tf32(acc, da + adv, db + adv, idesc_w, (ia | dx | k) ? 1u : 0u);
                tc_mma_tf32(acc, dal + adv, db + adv, idesc, 1u);
              } else {
                tc_mma_tf32(acc, da + adv, db + adv, idesc, (ia | dx | k) ? 1u : 0u);
                if (x3) {
                  tc_mma_tf32(acc, dal + adv, db + adv, idesc, 1u);
                  tc_mma_tf32(acc, da + adv, dbl + adv, idesc, 1u);
                }
              }
            }
          }
          tc_commit(b_empty(sb));
        }
        tc_commit(a_empty(sa));
      }
      tc_commit(tmem_full_bar);
      if (prof) {
        const long long t_iss = clock64() - t_entry;
        mbar_wait(tmem_full_bar, 0);
        printf("[c3 mma] first a_ready at %lld | all issued at %lld | accumulators complete at %lld | waited a_ready %lld "
               "b_full %lld (after the first stage)\n", t_first, t_iss, clock64() - t_entry, w_a, w_b);
      }
    }
  } else {
    const int t = threadIdx.x - 64;
    long long w_x = 0, t_x = 0;
    if (p.bias9) {
      // border-class biases of this image into shared memory while the first loads are in flight: thread j owns output
      // column n0 + j, reads its 9 tap values (one coalesced round trip) and forms the 9 class sums in tap order
      for (int j = t; j < p.BN; j += 128) {
        float tv[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) tv[tap] = n0 + j < p.N ? p.bias9[((size_t)img * 9 + tap) * p.N + n0 + j] : 0.f;
#pragma unroll
        for (int cls = 0; cls < 9; ++cls) {
          float acc = 0.f;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            if (c3_tap_valid(tap, cls)) acc += tv[tap];
          b9s[cls * p.BN + j] = acc;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    if (x3) {
      const int n4 = p.a_box_bytes / 16;
      for (int ia = 0; ia < NA; ++ia) {
        const int sa = ia % p.SA;
        const long long t0 = prof ? clock64() : 0;
        mbar_wait(a_full(sa), (ia / p.SA) & 1);
        const long long t1 = prof ? clock64() : 0;
        float4* hi = reinterpret_cast<float4*>(smem + (size_t)sa * a_stage);
        float4* lo = reinterpret_cast<float4*>(smem + (size_t)sa * a_stage + p.a_slot_bytes);
        for (int i = t; i < n4 && !(p.debug & 1); i += 128) {
          const float4 v = hi[i];
          const float4 h = rn_tf32_4(v);
          hi[i] = h;
          lo[i] = rn_tf32_4(v - h);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_warp(a_ready(sa));
        if (prof) { w_x += t1 - t0; t_x += clock64() - t1; }
      }
    }
    // ---- epilogue: TMEM -> registers (+ bias, + border-class bias) -> swizzled shared-memory tile holding only the
    // tile's real pixels (halo rows dropped: packed row = ly * W + lx) -> one TMA store per 32 output columns.  Two
    // staging tiles alternate in the (now idle) stage ring; the TMA unit clips rows past the image and columns past N.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool leader = threadIdx.x == 64;
    const long long t_w0 = prof ? clock64() : 0;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const long long t_e0 = prof ? clock64() : 0;
    const int ly = r / p.RW, lx = r - ly * p.RW;
    const bool rowok = ly < p.BH && lx < p.W;
    const int pr = ly * p.W + lx;                       // packed pixel row of the staging tile
    const int nchunks = (p.BN + 31) / 32;
    int it = 0;
    for (int tile = 0; tile < p.MT; ++tile) {
      const int yt = y0 + tile * p.BH;
      if (yt >= p.H) break;                             // uniform: the whole tile lies below the image
      const int y = yt + ly;
      const float* b9 = nullptr;
      if (p.bias9) {   // border class of this output pixel: which filter taps read inside the image (k_pool.cu)
        const int yy = min(y, p.H - 1), xx = min(lx, p.W - 1);
        const int cls = (yy < p.dil ? 0 : (yy >= p.H - p.dil ? 2 : 1)) * 3 + (xx < p.dil ? 0 : (xx >= p.W - p.dil ? 2 : 1));
        b9 = b9s + cls * p.BN;     // shared memory, indexed by the column inside this CTA's N tile
      }
      const uint32_t tbase = tmem_acc + (uint32_t)(tile * p.ncol_acc) + ((uint32_t)(quarter * 32) << 16);
      for (int ch = 0; ch < nchunks; ++ch, ++it) {
        const int c = ch * 32, n = n0 + c;
        if (n >= p.N) break;                            // uniform
        // per-column additive terms first (global loads in flight while the accumulator comes out of TMEM)
        float4 add[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          add[q] = f4s(0.f);
          if (n + q * 4 < p.N) {
            if (bias) add[q] = ld4(bias + n + q * 4);
            if (b9 && c + q * 4 < p.BN) add[q] = add[q] + ld4(b9 + c + q * 4);
          }
        }
        uint32_t v0[16], v1[16], w0[16], w1[16];
        const bool second = c + 16 < p.BN;       // uniform
        __syncwarp();
        tc_ld16_nw(tbase + (uint32_t)c, v0);     // warp-collective: executed by all 32 lanes, converged
        if (second) tc_ld16_nw(tbase + (uint32_t)(c + 16), v1);
        if (p.wide) {
          tc_ld16_nw(tbase + (uint32_t)(p.BN + c), w0);
          if (second) tc_ld16_nw(tbase + (uint32_t)(p.BN + c + 16), w1);
        }
        tc_ld_wait();
        tc_ld_fence(v0); tc_ld_fence(v1); tc_ld_fence(w0); tc_ld_fence(w1);
        float o[32];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          o[q] = __uint_as_float(v0[q]) + (p.wide ? __uint_as_float(w0[q]) : 0.f);
          o[16 + q] = second ? __uint_as_float(v1[q]) + (p.wide ? __uint_as_float(w1[q]) : 0.f) : 0.f;
        }
        uint8_t* stg = smem + (it & 1) * 16384;
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this tile is done
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (rowok) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stg + pr * 128 + ((q ^ (pr & 7)) << 4)) =
                f4(o[q * 4 + 0], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]) + add[q];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          tma_store_4d(&tmC, base + (uint32_t)(it & 1) * 16384u, n, yt * p.W, img, slot, p.accumulate != 0);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_read();          // shared memory must stay valid until the TMA unit has read it
    if (prof && t == 0)
      printf("[c3 xform/epilogue] transform: waited a_full %lld, worked %lld | waited for the accumulators %lld | epilogue "
             "%lld .. %lld (%lld clk)\n", w_x, t_x, t_e0 - t_w0, t_e0 - t_entry, clock64() - t_entry, clock64() - t_e0);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(ncols) : "memory");
  }
}

// =================================================================================================
// CTA-PAIR variant of tc_conv3_kernel (tcgen05 cta_group::2).  Round-1 finding: in 3xTF32 the 3x3 kernels are bound by
// the MMA issue rate of ONE SM's tensor pipe at cta_group::1 (a 128 x 112 x 8 TF32 MMA retires in ~113 clk = half the
// pipe's rate).  Here two CTAs of a cluster (two SMs of a TPC) form one MMA of M = 256: CTA r owns the accumulator rows
// of ITS pixel tiles (its own A ring, its own TMEM) and loads only HALF of every weight tile (rows
// [r*BN/2, (r+1)*BN/2) of the hi and lo planes) - the tensor pipes of both SMs read A from their own shared memory and
// B from both.  The leader (cluster rank 0) issues every MMA; completion is multicast to the barriers of both CTAs.
//   a_full (TMA -> transform, local) ; a_ready (256 arrivals on the LEADER's barrier: 128 local + 128 remote transform
//   threads) ; b_full (TMA -> local) ; b_peer (the peer's relay thread -> leader: "my half of B slot sb has landed") ;
//   a_empty / b_empty / tmem_full: tcgen05.commit.cta_group::2 ... multicast::cluster to both CTAs.
// The wide-B trick (one N = 2*BN MMA for hi|lo) does not combine with the N split and is not used: 3 MMAs per k-step.
// =================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {     // arrives on `bar` of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kC3Threads)
tc_conv3_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const float* __restrict__ bias, float* __restrict__ out, TcC3Params p, long long zs, int n_ctas) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int slot = blockIdx.z;
  { const size_t zo = (size_t)slot * zs; bias = zp(bias, zo); out += zo; p.bias9 = zp(p.bias9, zo); }
  const uint32_t rank = cluster_ctarank();               // 0 = leader (issues the MMAs)
  const int a_stage = 2 * p.a_slot_bytes;                // [hi][lo]
  const int b_stage = 2 * p.b_plane_bytes;               // [hi][lo], BN/2 rows each
  const uint32_t a_base = base, b_base = base + (uint32_t)p.SA * a_stage;
  const uint32_t bar0 = b_base + (uint32_t)p.SB * b_stage;
  uint8_t* bars_ptr = smem + (size_t)p.SA * a_stage + (size_t)p.SB * b_stage;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_ready = [&](int s) { return bar0 + 8u * (p.SA + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * p.SA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (3 * p.SA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (3 * p.SA + p.SB + s); };
  auto b_peer = [&](int s) { return bar0 + 8u * (3 * p.SA + 2 * p.SB + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (3 * p.SA + 3 * p.SB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * (3 * p.SA + 3 * p.SB + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ncols = (uint32_t)p.ncols_alloc;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_ready(s), 256); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.SB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); mbar_init(b_peer(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {     // one warp of EACH CTA of the pair takes part in the cta_group::2 allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  // this CTA's pixel tiles (CTAs past the end of the problem load zero-filled boxes and store nothing)
  const int cta = blockIdx.x;
  const bool live = cta < n_ctas;
  const int img = live ? cta / p.groups_per_image : 0;
  const int y0 = live ? (cta - img * p.groups_per_image) * p.MT * p.BH : 0;
  const int n0 = blockIdx.y * p.BN;
  const int nh = p.BN / 2;                 // weight rows this CTA loads
  const int KC = (p.C + 31) / 32;
  const int NA = 3 * KC;                   // A stages: (dy, kc)

  if (warp == 0) {
    if (lane == 0) {
      int ib = 0;
      for (int ia = 0; ia < NA; ++ia) {
        const int dy = ia / KC, kc = ia - dy * KC;
        const int sa = ia % p.SA;
        mbar_wait(a_empty(sa), ((ia / p.SA) & 1) ^ 1);
        mbar_expect_tx(a_full(sa), (uint32_t)p.a_box_bytes);
        // a CTA past the end reads image index B (out of bounds -> zero fill): it still feeds the pair's MMA
        tma_load_5d(a_base + (uint32_t)sa * a_stage, &tmA, a_full(sa), kc * 32, -p.dil, y0 + (dy - 1) * p.dil,
                    live ? img : 0x3fffffff, slot);
        for (int dx = 0; dx < 3; ++dx, ++ib) {
          const int sb = ib % p.SB;
          mbar_wait(b_empty(sb), ((ib / p.SB) & 1) ^ 1);
          mbar_expect_tx(b_full(sb), (uint32_t)b_stage);
          const uint32_t dst = b_base + (uint32_t)sb * b_stage;
          tma_load_5d(dst, &tmB, b_full(sb), kc * 32, dy * 3 + dx, n0 + (int)rank * nh, 0, slot);
          tma_load_5d(dst + p.b_plane_bytes, &tmB, b_full(sb), kc * 32, dy * 3 + dx, n0 + (int)rank * nh, 1, slot);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ---------------- MMA issuer (leader) ----------------
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((256u >> 4) << 24);
      int ib = 0;
      for (int ia = 0; ia < NA; ++ia) {
        const int kc = ia % KC;
        const int sa = ia % p.SA;
        mbar_wait_cluster(a_ready(sa), (ia / p.SA) & 1);
        const int rem = p.C - kc * 32;
        const int nk = rem >= 32 ? 4 : (rem + 7) / 8;
        const uint32_t a_hi = a_base + (uint32_t)sa * a_stage, a_lo = a_hi + p.a_slot_bytes;
        for (int dx = 0; dx < 3; ++dx, ++ib) {
          const int sb = ib % p.SB;
          mbar_wait(b_full(sb), (ib / p.SB) & 1);
          mbar_wait_cluster(b_peer(sb), (ib / p.SB) & 1);
          tc_fence_after();
          const uint32_t b_hi = b_base + (uint32_t)sb * b_stage, b_lo = b_hi + p.b_plane_bytes;
          const uint64_t db = make_kmajor_sw128_desc(b_hi), dbl = make_kmajor_sw128_desc(b_lo);
          for (int t = 0; t < p.MT; ++t) {
            const uint32_t row_off = (uint32_t)((t * p.BH * p.RW + dx * p.dil) * 128);
            const uint64_t da = make_kmajor_sw128_desc(a_hi + row_off), dal = make_kmajor_sw128_desc(a_lo + row_off);
            const uint32_t acc = tmem_acc + (uint32_t)(t * p.ncol_acc);
            for (int k = 0; k < nk; ++k) {
              const uint64_t adv = (uint64_t)(2 * k);
              tc_mma_tf32_pair(acc, da + adv, db + adv, idesc, (ia | dx | k) ? 1u : 0u);
              tc_mma_tf32_pair(acc, dal + adv, db + adv, idesc, 1u);
              tc_mma_tf32_pair(acc, da + adv, dbl + adv, idesc, 1u);
            }
          }
          tc_commit_pair(b_empty(sb));
        }
        tc_commit_pair(a_empty(sa));
      }
      tc_commit_pair(tmem_full_bar);
    } else if (lane == 0) {
      // ---------------- relay (peer): tell the leader when this CTA's half of a weight slot has landed ----------------
      const int NB = 3 * NA;
      for (int ib = 0; ib < NB; ++ib) {
        const int sb = ib % p.SB;
        mbar_wait(b_full(sb), (ib / p.SB) & 1);
        mbar_arrive_remote(map_to_cta(b_peer(sb), 0));
      }
    }
  } else {
    const int t = threadIdx.x - 64;
    const int n4 = p.a_box_bytes / 16;
    for (int ia = 0; ia < NA; ++ia) {
      const int sa = ia % p.SA;
      mbar_wait(a_full(sa), (ia / p.SA) & 1);
      float4* hi = reinterpret_cast<float4*>(smem + (size_t)sa * a_stage);
      float4* lo = reinterpret_cast<float4*>(smem + (size_t)sa * a_stage + p.a_slot_bytes);
      for (int i = t; i < n4; i += 128) {
        const float4 v = hi[i];
        const float4 h = rn_tf32_4(v);
        hi[i] = h;
        lo[i] = rn_tf32_4(v - h);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_remote(map_to_cta(a_ready(sa), 0));      // the leader's barrier counts both CTAs' transform threads
    }
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    mbar_wait_cluster(tmem_full_bar, 0);
    tc_fence_after();
    const int ly = r / p.RW, lx = r - ly * p.RW;
    for (int tile = 0; tile < p.MT; ++tile) {
      const int y = y0 + tile * p.BH + ly;
      const bool valid = live && ly < p.BH && lx < p.W && y < p.H;
      float* orow = out + (((size_t)img * p.H + y) * p.W + lx) * p.ldc;
      const float* b9 = nullptr;
      if (p.bias9) {
        const int cls = (y < p.dil ? 0 : (y >= p.H - p.dil ? 2 : 1)) * 3 + (lx < p.dil ? 0 : (lx >= p.W - p.dil ? 2 : 1));
        b9 = p.bias9 + ((size_t)img * 9 + cls) * p.N;
      }
      const uint32_t tbase = tmem_acc + (uint32_t)(tile * p.ncol_acc) + ((uint32_t)(quarter * 32) << 16);
      for (int c = 0; c < p.BN; c += 16) {
        uint32_t v[16];
        __syncwarp();
        tc_ld16(tbase + (uint32_t)c, v);
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = n0 + c + q * 4;
            if (n < p.N) {
              float4 o = f4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                            __uint_as_float(v[q * 4 + 3]));
              if (bias) o = o + ld4(bias + n);
              if (b9) o = o + ld4(b9 + n);
              if (p.accumulate) o = o + ld4(orow + n);
              st4(orow + n, o);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  // neither CTA may leave (or free its TMEM) while the pair's MMAs / multicast arrivals can still touch it
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static bool encode(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// outermost "task slot" dimension of every tensor map: nz slots, zs floats apart (one slot: any legal stride)
static inline cuuint64_t slot_stride_bytes(cuuint64_t natural) {
  if (MLIIS_NZ > 1) return (cuuint64_t)MLIIS_ZS * 4;
  return (natural + 15) / 16 * 16;
}

// Shared-memory budget of the one-CTA-per-SM tensor-core kernels (rings of operand stages).  MLIIS_TC_SMEM_KB < 216 leaves
// room for CTAs of the HBM-bound kernels of ANOTHER task group on the same SM (experiment knob, read once).
static int tc_smem_budget() {
  static int kb = -1;
  if (kb < 0) { const char* e = getenv("MLIIS_TC_SMEM_KB"); kb = e ? atoi(e) : 216; if (kb < 96 || kb > 216) kb = 216; }
  return kb * 1024;
}
int tc_pick_bn(int N) {
  int tiles = (N + 255) / 256;
  int bn = (N + tiles - 1) / tiles;
  // several N tiles: every tile must end on a 32-column boundary (the epilogue stores 32-column TMA boxes; a single
  // tile may overhang N because the TMA unit clips at the tensor edge)
  return tiles > 1 ? (bn + 31) / 32 * 32 : (bn + 15) / 16 * 16;
}

bool tc_supported(int conv, int W, int C, int N) {
  if (C % 4 || N % 4) return false;
  if (conv && (W > 128 || W < 1)) return false;
  return encode_fn() != nullptr;
}

// second-generation 3x3 path; returns false when the shape does not fit (caller uses tc_conv_kernel)
static bool tc_conv3(const float* A, int lda, const float* Wt, const float* bias, float* out, int ldc, int B, int H, int W,
                     int C, int dil, int N, int accumulate, int split, cudaStream_t s, const float* bias9) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MLIIS_TC_CONV3"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  TcC3Params p{};
  static int boff = -1, dbg3 = -1;
  if (boff < 0) { const char* e = getenv("MLIIS_TC_BASEOFF"); boff = e ? atoi(e) : 0; }
  if (dbg3 < 0) { const char* e = getenv("MLIIS_TC_DEBUG"); dbg3 = e ? atoi(e) : 0; }
  p.boff = boff;
  p.debug = dbg3;
  p.bias9 = bias9;
  p.H = H; p.W = W; p.C = C; p.dil = dil; p.N = N; p.ldc = ldc; p.accumulate = accumulate;
  p.split = split == 3 ? 3 : 1;
  p.RW = W + 2 * dil;
  if (p.RW > 128) return false;
  p.BH = 128 / p.RW;
  if (p.BH > H) p.BH = H;
  p.MT = (dbg3 & 4) ? 1 : 2;
  p.tiles_per_image = (H + p.BH - 1) / p.BH;
  if (p.tiles_per_image < 2) p.MT = 1;
  p.groups_per_image = (p.tiles_per_image + p.MT - 1) / p.MT;
  const int box_rows = p.MT * p.BH * p.RW;
  if (box_rows > 256 || p.MT * p.BH > 256) return false;
  p.BN = tc_pick_bn(N);
  // one N tile too wide for the rings (the dgrad of conv2d_2: N = 224 leaves room for ONE 56 KB weight slot): fall back
  // to 128-column tiles (grid.y = ceil(N / 128); the TMA unit zero-fills weight rows and clips output columns past N)
  // instead of giving the layer to the generic one-box-per-tap kernel (113 us there)
  if (p.split == 3 && p.BN > 128 && p.MT == 2) {
    const int slot_rows_w = (p.MT - 1) * p.BH * p.RW + 2 * dil + 128;
    const int a_bytes_w = ((slot_rows_w > box_rows ? slot_rows_w : box_rows) * 128 + 1023) / 1024 * 1024;
    if ((tc_smem_budget() - 2 * 2 * a_bytes_w) / (2 * p.BN * 128) < 2) p.BN = 128;
  }
  // CTA pairs (cta_group::2, M = 256): 3xTF32 only (the kernel these layers are MMA-issue-bound in), even BN halves
  static int pair_on = -1;
  if (pair_on < 0) { const char* e = getenv("MLIIS_TC_PAIR"); pair_on = e ? atoi(e) : 0; }
  const bool pair = pair_on && p.split == 3 && p.BN % 16 == 0 && !(dbg3 & 3);
  p.wide = !pair && p.split == 3 && 2 * p.BN <= 256 && 2 * p.BN * p.MT <= 512 && !(dbg3 & 8);
  p.ncol_acc = p.wide ? 2 * p.BN : p.BN;
  p.ncols_alloc = 32;
  while (p.ncols_alloc < p.ncol_acc * p.MT) p.ncols_alloc <<= 1;
  if (p.ncols_alloc > 512) return false;
  p.a_box_bytes = box_rows * 128;
  // the last tile's MMAs read up to (MT-1)*BH*RW + 2*dil + 128 rows: keep them inside the slot
  int slot_rows = (p.MT - 1) * p.BH * p.RW + 2 * dil + 128;
  if (slot_rows < box_rows) slot_rows = box_rows;
  p.a_slot_bytes = (slot_rows * 128 + 1023) / 1024 * 1024;
  p.b_plane_bytes = (pair ? p.BN / 2 : p.BN) * 128;
  const int planes = p.split == 3 ? 2 : 1;
  const int budget = tc_smem_budget();
  p.SA = p.split == 3 ? 2 : 3;
  p.SB = (budget - p.SA * planes * p.a_slot_bytes) / (planes * p.b_plane_bytes);
  if (p.SB > 8) p.SB = 8;
  if (p.SB < 2) return false;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)MLIIS_NZ};
    cuuint64_t str[4] = {(cuuint64_t)lda * 4, (cuuint64_t)W * lda * 4, (cuuint64_t)H * W * lda * 4,
                         slot_stride_bytes((cuuint64_t)B * H * W * lda * 4)};
    cuuint32_t box[5] = {32, (cuuint32_t)p.RW, (cuuint32_t)(p.MT * p.BH), 1, 1};
    if (!encode(&tmA, A, 5, dims, str, box)) return false;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, 9, (cuuint64_t)N, (cuuint64_t)planes, (cuuint64_t)MLIIS_NZ};
    cuuint64_t str[4] = {(cuuint64_t)C * 4, (cuuint64_t)9 * C * 4, (cuuint64_t)N * 9 * C * 4,
                         slot_stride_bytes((cuuint64_t)planes * N * 9 * C * 4)};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)(pair ? p.BN / 2 : p.BN), 1, 1};
    if (!encode(&tmB, Wt, 5, dims, str, box)) return false;
  }
  CUtensorMap tmC;
  if (!pair) {
    // output [slot, B, H*W, N]: one box = the BH*W pixels of a tile x 32 columns (rows past the image are clipped)
    cuuint64_t cd[4] = {(cuuint64_t)N, (cuuint64_t)H * W, (cuuint64_t)B, (cuuint64_t)MLIIS_NZ};
    cuuint64_t cs[3] = {(cuuint64_t)ldc * 4, (cuuint64_t)H * W * ldc * 4, slot_stride_bytes((cuuint64_t)B * H * W * ldc * 4)};
    cuuint32_t cb[4] = {32, (cuuint32_t)(p.BH * W), 1, 1};
    if (p.BH * W > 128 || !encode(&tmC, out, 4, cd, cs, cb)) return false;
  }
  const size_t smem = (size_t)p.SA * planes * p.a_slot_bytes + (size_t)p.SB * planes * p.b_plane_bytes +
                      (3 * p.SA + 3 * p.SB + 2) * 8 + 32 + (bias9 ? 9 * p.BN * 4 : 0) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(tc_conv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tc_conv3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  const int n_ctas = B * p.groups_per_image;
  if (pair) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((n_ctas + 1) / 2 * 2, (N + p.BN - 1) / p.BN, MLIIS_NZ);
    cfg.blockDim = dim3(kC3Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const long long zs = MLIIS_ZS;
    MLIIS_COUNT(), cudaLaunchKernelEx(&cfg, tc_conv3_pair_kernel, tmA, tmB, bias, out, p, zs, n_ctas);
    return true;
  }
  dim3 grid(n_ctas, (N + p.BN - 1) / p.BN, MLIIS_NZ);
  MLIIS_COUNT(), tc_conv3_kernel<<<grid, kC3Threads, smem, s>>>(tmA, tmB, tmC, bias, p, MLIIS_ZS);
  return true;
}

// persistent pointwise path; returns false when the shape does not qualify (caller uses tc_conv_kernel)
static bool tc_pw(const float* A, int lda, const float* Wt, const float* bias, float* out, int ldc, int M, int C, int N,
                  int accumulate, int split, cudaStream_t s, const float* pa, const float* pb, const float* gate, int HW) {
  static int enabled = -1, dbg = -1;
  if (enabled < 0) { const char* e = getenv("MLIIS_TC_PW"); enabled = e ? atoi(e) : 1; }
  if (dbg < 0) { const char* e = getenv("MLIIS_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  if (!enabled || (dbg & 3)) return false;
  TcPwParams p{};
  p.debug = dbg;
  p.M = M; p.C = C; p.N = N; p.accumulate = accumulate; p.split = split == 3 ? 3 : 1;
  p.pa = pa; p.pb = pb; p.gate = gate; p.HW = HW > 0 ? HW : 1;
  p.BN = tc_pick_bn(N);
  if (p.BN < N) return false;                      // one N tile only
  p.KB = (C + 31) / 32;
  if (p.KB > 8) return false;
  p.n_tiles = (M + 127) / 128;
  int sms = 148;
  const int per_slot = sms / MLIIS_NZ > 0 ? sms / MLIIS_NZ : 1;
  if (p.n_tiles < 2 * per_slot) return false;      // fewer than two tiles per CTA: nothing to pipeline
  const int planes = p.split == 3 ? 2 : 1;
  p.wide = p.split == 3 && 2 * p.BN <= 256;
  p.ncol_acc = p.wide ? 2 * p.BN : p.BN;
  p.ncols_alloc = 32;
  while (p.ncols_alloc < 2 * p.ncol_acc) p.ncols_alloc <<= 1;
  if (p.ncols_alloc > 512) return false;
  p.b_res_bytes = p.KB * planes * p.BN * 128;
  const int a_stage = planes * kABytes;
  const int Kp = (C + 31) / 32 * 32;
  const int tail = 8 * (3 * 6 + 5) + 32 + 8 * Kp + 1024;
  p.SA = (tc_smem_budget() + 4 * 1024 - p.b_res_bytes - 4 * 16384 - tail) / a_stage;
  if (p.SA > 6) p.SA = 6;
  if (p.SA < 2) return false;
  CUtensorMap tmA, tmB, tmC;
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)M, (cuuint64_t)MLIIS_NZ};
    cuuint64_t str[2] = {(cuuint64_t)lda * 4, slot_stride_bytes((cuuint64_t)M * lda * 4)};
    cuuint32_t box[3] = {32, 128, 1};
    if (!encode(&tmA, A, 3, dims, str, box)) return false;
    cuuint64_t cd[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)MLIIS_NZ};
    cuuint64_t cs[2] = {(cuuint64_t)ldc * 4, slot_stride_bytes((cuuint64_t)M * ldc * 4)};
    if (!encode(&tmC, out, 3, cd, cs, box)) return false;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, 1, (cuuint64_t)N, (cuuint64_t)planes, (cuuint64_t)MLIIS_NZ};
    cuuint64_t str[4] = {(cuuint64_t)C * 4, (cuuint64_t)C * 4, (cuuint64_t)N * C * 4,
                         slot_stride_bytes((cuuint64_t)planes * N * C * 4)};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)p.BN, 1, 1};
    if (!encode(&tmB, Wt, 5, dims, str, box)) return false;
  }
  const size_t smem = (size_t)p.b_res_bytes + (size_t)p.SA * a_stage + 4 * 16384 + 8 * (3 * p.SA + 5) + 32 + 8 * Kp + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(tc_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  const int gx = p.n_tiles < per_slot ? p.n_tiles : per_slot;
  MLIIS_COUNT(), tc_pw_kernel<<<dim3(gx, 1, MLIIS_NZ), kPwThreads, smem, s>>>(tmA, tmB, tmC, bias, p, MLIIS_ZS);
  return true;
}

// Wt layout expected by the kernel: [N][taps][C]  (K-major rows of B)
bool tc_conv(const float* A, int lda, const float* Wt, const float* bias, float* out, int ldc, int conv, int M, int B,
             int H, int W, int C, int taps, int dil, int N, int accumulate, int split, cudaStream_t s,
             const float* pa, const float* pb, const float* gate, int HW, const float* bias9) {
  TcParams p{};
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("MLIIS_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  p.split = split == 3 ? 3 : 1;
  if (pa && conv) return false;
  p.pa = pa; p.pb = pb; p.gate = gate; p.HW = HW > 0 ? HW : 1;
  p.conv = conv; p.M = M; p.H = H; p.W = W; p.C = C; p.taps = taps; p.dil = dil; p.N = N;
  p.accumulate = accumulate;
  if (conv && taps == 9 && tc_conv3(A, lda, Wt, bias, out, ldc, B, H, W, C, dil, N, accumulate, p.split, s, bias9))
    return true;
  if (bias9) return false;        // only tc_conv3_kernel adds the border-class bias
  if (!conv && taps == 1 && tc_pw(A, lda, Wt, bias, out, ldc, M, C, N, accumulate, p.split, s, pa, pb, gate, HW)) return true;
  p.BN = tc_pick_bn(N);
  CUtensorMap tmA, tmB, tmC;
  int grid_x;
  if (conv) {
    p.BH = 128 / W;
    if (p.BH > H) p.BH = H;
    p.tiles_per_image = (H + p.BH - 1) / p.BH;
    p.a_box_bytes = p.BH * W * 128;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)MLIIS_NZ};
    cuuint64_t str[4] = {(cuuint64_t)lda * 4, (cuuint64_t)W * lda * 4, (cuuint64_t)H * W * lda * 4,
                         slot_stride_bytes((cuuint64_t)B * H * W * lda * 4)};
    cuuint32_t box[5] = {32, (cuuint32_t)W, (cuuint32_t)p.BH, 1, 1};
    if (!encode(&tmA, A, 5, dims, str, box)) return false;
    grid_x = B * p.tiles_per_image;
    // output [slot, B, H*W, N]: one box = the BH*W pixels of the tile x 32 columns (rows past the image are clipped)
    cuuint64_t cd[4] = {(cuuint64_t)N, (cuuint64_t)H * W, (cuuint64_t)B, (cuuint64_t)MLIIS_NZ};
    cuuint64_t cs[3] = {(cuuint64_t)ldc * 4, (cuuint64_t)H * W * ldc * 4, slot_stride_bytes((cuuint64_t)B * H * W * ldc * 4)};
    cuuint32_t cb[4] = {32, (cuuint32_t)(p.BH * W), 1, 1};
    if (!encode(&tmC, out, 4, cd, cs, cb)) return false;
  } else {
    p.a_box_bytes = kABytes;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)M, (cuuint64_t)MLIIS_NZ};
    cuuint64_t str[2] = {(cuuint64_t)lda * 4, slot_stride_bytes((cuuint64_t)M * lda * 4)};
    cuuint32_t box[3] = {32, 128, 1};
    if (!encode(&tmA, A, 3, dims, str, box)) return false;
    grid_x = (M + 127) / 128;
    cuuint64_t cd[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)MLIIS_NZ};
    cuuint64_t cs[2] = {(cuuint64_t)ldc * 4, slot_stride_bytes((cuuint64_t)M * ldc * 4)};
    if (!encode(&tmC, out, 3, cd, cs, box)) return false;
  }
  {
    // operand planes: [hi][N][taps][C] and, for split == 3, [lo][N][taps][C] right behind it
    const int planes = p.split == 3 ? 2 : 1;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)taps, (cuuint64_t)N, (cuuint64_t)planes, (cuuint64_t)MLIIS_NZ};
    cuuint64_t str[4] = {(cuuint64_t)C * 4, (cuuint64_t)taps * C * 4, (cuuint64_t)N * taps * C * 4,
                         slot_stride_bytes((cuuint64_t)planes * N * taps * C * 4)};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)p.BN, 1, 1};
    if (!encode(&tmB, Wt, 5, dims, str, box)) return false;
  }
  const int stage_bytes = (p.split == 3 ? 2 : 1) * (kABytes + p.BN * 128);
  const int KB = taps * ((C + 31) / 32);
  const int Kp = (C + 31) / 32 * 32;
  const int tail_bytes = 8 * (3 * 6 + 2) + 16 + (pa ? 16 * Kp : 0) + 1024;
  // small stages: size the ring so that two CTAs share an SM (their TMA round trips and fixed start-up costs overlap);
  // big stages (BN >= 128 with both planes): one CTA per SM with a deeper ring
  // Short-K layers (the MBConv pointwise convolutions: KB <= 8 k-blocks) are bound by per-CTA latency, not by the depth
  // of the ring: two stages and two CTAs per SM beat five stages and one (project 144->24 at 56x56, 6 slots per launch:
  // 96 us with one resident CTA).  Long-K layers keep the >= 3 stage rule.
  const int budget2 = (227 * 1024) / 2 - tail_bytes;
  int stages = budget2 / stage_bytes;
  if (stages < (KB <= 8 ? 2 : 3)) stages = (200 * 1024 - tail_bytes) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages > KB) stages = KB;        // small-K problems: less shared memory -> several CTAs per SM
  if (stages < 1) return false;
  p.stages = stages;
  p.region_bytes = stages * stage_bytes;
  if (p.region_bytes < 65536) p.region_bytes = 65536;     // the epilogue reuses the ring as four 16 KB staging tiles
  const size_t smem = (size_t)p.region_bytes + ((8 * (3 * stages + 2) + 15) & ~15) + (pa ? 16 * Kp : 0) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  dim3 grid(grid_x, (N + p.BN - 1) / p.BN, MLIIS_NZ);
  MLIIS_COUNT(), tc_conv_kernel<<<grid, kTcThreads, smem, s>>>(tmA, tmB, tmC, bias, p, MLIIS_ZS);
  return true;
}

// =================================================================================================
// wgrad on tensor cores:  dW[tap][c][n] = sum_pixels A[pixel + offset(tap), c] * G[pixel, n]
//
// The reduction (MMA K) dimension is the pixel index, so both operands are "MN-major": channels are contiguous
// in memory.  One TMA box {32 channels, 32 pixels} (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) lands as a
// [32 pixel rows][128 B] slab = the canonical MN-major SWIZZLE_128B_BASE32B atom stack, the only shared-memory layout
// tcgen05 accepts for MN-major TF32 operands (4 pixel rows x 128 B per atom, SBO = 512 B between atoms along K,
// LBO = 4096 B between 32-channel groups along M/N).  One tcgen05.mma (K = 8) consumes two atoms per channel group.  grid = (128-channel tiles of A, taps, pixel-range splits); each CTA writes its fp32 partial
// [128 x N] and the partials are reduced in a fixed order by reduce_partials (deterministic, no atomics).
// =================================================================================================
struct TcWgParams {
  int conv, M;                 // plain: A [M, C], G [M, N]
  int H, W, BX, BY, tiles_x, tiles_y;   // conv: pixel box BX x BY (= 32 pixels)
  int C, N, BN, NG;            // NG = 32-channel groups of G
  int taps, dil;
  int splits, tiles_total;
  int stages, split;
  // share: conv mode, one CTA makes the 3 horizontal taps of a filter row from ONE halo'd A box {32 ch, BX+2*dil, BY}
  // (tap dx = the same slab read through a descriptor start shifted by dx*dil pixel rows); grid.y = 3 filter rows
  int share, a_group_bytes, a_tx_bytes;
  int nga_alloc;               // 32-channel groups of A that a stage holds (4, or ceil(C/32) when C < 128)
  int debug;                   // MLIIS_TC_DEBUG bits (bottleneck experiments only): 1 skip transform, 2 skip MMA
  const float* pa; const float* pb; const float* gate;   // A prologue (plain mode), as in tc_conv_kernel
  int HW;
};

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  // MN-major TF32 operands only exist in the SWIZZLE_128B_BASE32B layout (layout type 1): rows of 128 B (32 channels)
  // whose four 32-byte chunks are XOR-swizzled with (pixel row & 3); atom = 4 pixel rows (512 B).
  //  LBO (bits 16..29) = 4096 B >> 4 : stride between 32-channel groups along M/N
  //  SBO (bits 32..45) =  512 B >> 4 : stride between 4-row atoms along K
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (256ull << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(lbo_bytes >> 4) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}

constexpr int kWgGroupBytes = 32 * 128;   // one TMA box: 32 pixel rows x 32 channels fp32
// wgrad splits BOTH operands in the kernel (activations and gradients), so it runs 8 transform/epilogue warps
constexpr int kWgXformThreads = 512;
constexpr int kWgThreads = 64 + kWgXformThreads;

__global__ void __launch_bounds__(kWgThreads)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG,
                float* __restrict__ partial, TcWgParams p, long long zs) {
  extern __shared__ uint8_t smem_raw[];
  if (p.debug & 16) return;   // experiment: cost of everything except the tensor-core kernels
  const int slot = blockIdx.z / p.splits;      // grid.z = (task slot, pixel-range split)
  { const size_t zo = (size_t)slot * zs; partial += zo; p.pa = zp(p.pa, zo); p.pb = zp(p.pb, zo); p.gate = zp(p.gate, zo); }
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const bool x3 = p.split == 3;
  const int a_group = p.share ? p.a_group_bytes : kWgGroupBytes;
  // a stage holds only the channel groups that exist (C = 16..96: 1-3 of 4).  The M = 128 MMAs still read four groups
  // (LBO apart): the rows past nga_alloc alias the following planes, give garbage accumulator rows c >= C and are never
  // stored.  Smaller stages = more loads in flight per SM (these layers are bound by the load round trip, not by the
  // MMAs or the transform: skip experiments of profiles/r02zn) and two CTAs per SM.
  const int a_bytes = p.nga_alloc * a_group, g_bytes = p.NG * kWgGroupBytes;
  const int a_lo = a_bytes;                                  // offset of the A lo plane (x3)
  const int g_off = x3 ? 2 * a_bytes : a_bytes;
  const int g_lo = g_off + g_bytes;
  const int stage_bytes = g_off + (x3 ? 2 : 1) * g_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  const uint32_t bar0 = base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (p.stages + s); };
  auto ready_bar = [&](int s) { return bar0 + 8u * (2 * p.stages + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (3 * p.stages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // wide: the G lo plane starts right behind the hi plane's last 32-channel group, so [G_hi | G_lo] is one MN-major
  // operand of 2*NG groups: one N = 64*NG MMA forms a_hi*g_hi | a_hi*g_lo (epilogue adds the two column ranges)
  const bool wide = x3 && p.NG <= 4 && !p.share;       // share: three tap accumulators of BN columns each
  const int wide_off = p.NG * 32;
  const int ntap = p.share ? 3 : 1;
  uint32_t ncols = 32;
  while ((int)ncols < (p.share ? 3 * p.BN : (wide ? 2 * wide_off : p.BN))) ncols <<= 1;
  const int c0 = blockIdx.x * 128, tap = blockIdx.y, split = blockIdx.z - slot * p.splits;
  // 32-channel groups of the A tile that hold real channels: the others are never loaded or transformed - their shared
  // memory is zeroed once below and the MMAs read zero rows (C = 16 .. 96 layers used to move three or two all-zero boxes
  // per stage through TMA and the transform)
  const int nga = min(4, (p.C - c0 + 31) / 32);
  const int per = (p.tiles_total + p.splits - 1) / p.splits;
  const int t_beg = split * per, t_end = min(p.tiles_total, t_beg + per);
  const int KB = max(t_end - t_beg, 0);
  const int dy = p.conv ? ((p.share ? tap : tap / 3) - 1) * p.dil : 0;
  const int dx = p.conv ? (p.share ? -p.dil : (tap % 3 - 1) * p.dil) : 0;       // share: left edge of the halo box

  // TMA producer step kb; the first min(stages, KB) steps are issued before the block-wide sync (see tc_conv_kernel)
  auto produce = [&](int kb) {
    const int s = kb % p.stages;
    mbar_wait(empty_bar(s), ((kb / p.stages) & 1) ^ 1);
    const uint32_t sa = base + (uint32_t)s * stage_bytes, sg = sa + g_off;
    mbar_expect_tx(full_bar(s), (uint32_t)((p.share ? (p.a_tx_bytes >> 2) : kWgGroupBytes) * nga + g_bytes));
    const int t = t_beg + kb;
    if (p.conv) {
      const int per_img = p.tiles_x * p.tiles_y;
      const int img = t / per_img, rem = t - img * per_img, ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * p.BX, y0 = ty * p.BY;
      for (int g = 0; g < nga; ++g)
        tma_load_5d(sa + g * a_group, &tmA, full_bar(s), c0 + 32 * g, x0 + dx, y0 + dy, img, slot);
      for (int g = 0; g < p.NG; ++g) tma_load_5d(sg + g * kWgGroupBytes, &tmG, full_bar(s), 32 * g, x0, y0, img, slot);
    } else {
      const int m0 = t * 32;
      for (int g = 0; g < nga; ++g) tma_load_3d(sa + g * kWgGroupBytes, &tmA, full_bar(s), c0 + 32 * g, m0, slot);
      for (int g = 0; g < p.NG; ++g) tma_load_3d(sg + g * kWgGroupBytes, &tmG, full_bar(s), 32 * g, m0, slot);
    }
  };
  const int n_pre = min(p.stages, KB);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(ready_bar(s), kWgXformThreads / 32); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // barrier words (generic proxy) -> TMA (async proxy)
    for (int kb = 0; kb < n_pre; ++kb) produce(kb);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (nga < p.nga_alloc && threadIdx.x >= 64) {
    const int z0 = nga * a_group / 16, z1 = a_bytes / 16;          // float4 range of the unused groups in a plane
    for (int s = 0; s < p.stages; ++s) {
      float4* st = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
      for (int i = z0 + (int)threadIdx.x - 64; i < z1; i += kWgXformThreads) {
        st[i] = f4s(0.f);
        if (x3) st[a_lo / 16 + i] = f4s(0.f);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros -> tensor-core operand reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = n_pre; kb < KB; ++kb) produce(kb);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // TF32, fp32 accumulate, BOTH operands MN-major (bits 15 and 16)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_w = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(2 * wide_off >> 3) << 17) | ((128u >> 4) << 24);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % p.stages;
        mbar_wait(ready_bar(s), (kb / p.stages) & 1);
        tc_fence_after();
        const uint32_t sa = base + (uint32_t)s * stage_bytes;
        const uint64_t dg = make_mnmajor_sw128_desc(sa + g_off), dgl = make_mnmajor_sw128_desc(sa + g_lo);
        if (p.share) {
          // pixel rows of the halo slab: row(y, x) = y * (BX + 2*dil) + x ; tap t starts t*dil rows further right
          const int rw = p.BX + 2 * p.dil, kpr = p.BX / 8;      // K-steps (8 pixels) per image row of the box
          for (int t = 0; t < 3 && !(p.debug & 2); ++t) {
            const uint32_t acc = tmem_acc + (uint32_t)(t * p.BN);
            for (int k = 0; k < 4; ++k) {
              const uint32_t row = (uint32_t)((k / kpr) * rw + t * p.dil + (k % kpr) * 8);
              const uint64_t da = make_mnmajor_sw128_desc_lbo(sa + row * 128u, (uint32_t)a_group);
              const uint64_t dal = make_mnmajor_sw128_desc_lbo(sa + a_lo + row * 128u, (uint32_t)a_group);
              const uint64_t adv = (uint64_t)(64 * k);
              tc_mma_tf32(acc, da, dg + adv, idesc, (kb | k) ? 1u : 0u);
              if (x3) {
                tc_mma_tf32(acc, dal, dg + adv, idesc, 1u);
                tc_mma_tf32(acc, da, dgl + adv, idesc, 1u);
              }
            }
          }
        } else {
        const uint64_t da = make_mnmajor_sw128_desc(sa), dal = make_mnmajor_sw128_desc(sa + a_lo);
        for (int k = 0; k < 4 && !(p.debug & 2); ++k) {            // 4 atoms of 8 pixel rows: +1024 B each
          const uint64_t adv = (uint64_t)(64 * k);
          if (wide) {
            tc_mma_tf32(tmem_acc, da + adv, dg + adv, idesc_w, (kb | k) ? 1u : 0u);
            tc_mma_tf32(tmem_acc, dal + adv, dg + adv, idesc, 1u);
          } else {
            tc_mma_tf32(tmem_acc, da + adv, dg + adv, idesc, (kb | k) ? 1u : 0u);
            if (x3) {
              tc_mma_tf32(tmem_acc, dal + adv, dg + adv, idesc, 1u);
              tc_mma_tf32(tmem_acc, da + adv, dgl + adv, idesc, 1u);
            }
          }
        }
        }
        tc_commit(empty_bar(s));
      }
      tc_commit(tmem_full_bar);
    }
  } else {
    const int t = threadIdx.x - 64;
    // A tile (plain mode): float4 i = t + 512*j, j = 0,1 -> 32-channel group (t>>8) + 2*j, pixel row (t&255)>>3 and
    // 16-byte chunk t&7 are the same for both: the prologue coefficients of a thread never change (c0 is per CTA)
    const int arow = (t & 255) >> 3, af = t & 7;
    // SWIZZLE_128B_ATOM_32B: logical 32-byte chunk = physical chunk XOR (row & 3)
    const int kk = ((((af >> 1) ^ (arow & 3)) << 1) | (af & 1)) << 2;
    float4 pa4[2], pb4[2];
    int kch[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      kch[j] = c0 + 32 * ((t >> 8) + 2 * j) + kk;
      const bool in = p.pa != nullptr && kch[j] < p.C;
      pa4[j] = in ? ld4(p.pa + kch[j]) : f4s(0.f);      // zero coefficients keep the TMA zero fill: swish(0) = 0
      pb4[j] = in ? ld4(p.pb + kch[j]) : f4s(0.f);
    }
    const int a_valid4 = nga * a_group / 16;           // float4 of a plane that hold loaded data
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % p.stages;
      float4 gt[2];
      if (p.gate) {
        const int m = (t_beg + kb) * 32 + arow;
        const int im = m < p.M ? m / p.HW : 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) gt[j] = kch[j] < p.C ? ld4(p.gate + (size_t)im * p.C + kch[j]) : f4s(0.f);
      }
      mbar_wait(full_bar(s), (kb / p.stages) & 1);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      float4* a_hi = reinterpret_cast<float4*>(st);
      float4* a_lop = reinterpret_cast<float4*>(st + a_lo);
#pragma unroll
      for (int j = 0; j < 1024 / kWgXformThreads && !(p.debug & 1); ++j) {            // A: 4 groups x 256 float4
        const int i = t + kWgXformThreads * j;
        if (i >= a_valid4) continue;                 // plain mode: group (t >> 8) + 2 * j holds no real channel
        float4 v = a_hi[i];
        if (p.pa) {
          v = swish_fast4(affine4(v, pa4[j], pb4[j]));
          if (p.gate) v = v * gt[j];
        }
        const float4 h = rn_tf32_4(v);
        a_hi[i] = h;
        if (x3) a_lop[i] = rn_tf32_4(v - h);
      }
      if (p.share) {       // halo slab (no prologue in conv mode): 4 groups x a_group bytes, plain hi/lo split
        for (int i = t + 1024; i < a_valid4 && !(p.debug & 1); i += kWgXformThreads) {
          const float4 v = a_hi[i];
          const float4 h = rn_tf32_4(v);
          a_hi[i] = h;
          if (x3) a_lop[i] = rn_tf32_4(v - h);
        }
      }
      float4* g_hi = reinterpret_cast<float4*>(st + g_off);
      float4* g_lop = reinterpret_cast<float4*>(st + g_lo);
      const int ng4 = p.NG * 256;
      for (int i = t; i < ng4 && !(p.debug & 1); i += kWgXformThreads) {
        const float4 v = g_hi[i];
        const float4 h = rn_tf32_4(v);
        g_hi[i] = h;
        if (x3) g_lop[i] = rn_tf32_4(v - h);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_warp(ready_bar(s));
    }
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int c = c0 + r;
    if (KB > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    for (int tt = 0; tt < ntap; ++tt) {
      const int tap_out = p.share ? tap * 3 + tt : tap;
      float* orow = partial + (((size_t)split * p.taps + tap_out) * p.C + c) * p.N;
      const uint32_t tbase = tmem_acc + (uint32_t)(tt * p.BN) + ((uint32_t)(quarter * 32) << 16);
      for (int cc = ((warp - 2) >> 2) * 16; cc < p.BN; cc += 16 * (kWgXformThreads / 128)) {   // warps share a lane quarter
        uint32_t v[16];
        __syncwarp();
        if (KB > 0) {
          tc_ld16(tbase + (uint32_t)cc, v);
          if (wide) {
            uint32_t w[16];
            tc_ld16(tbase + (uint32_t)(wide_off + cc), w);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(w[q]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0u;
        }
        if (c < p.C) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = cc + q * 4;
            if (n < p.N)
              st4(orow + n, f4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                               __uint_as_float(v[q * 4 + 3])));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(ncols) : "memory");
  }
}

static int wg_splits(int ctiles, int taps, int tiles_total, int ctas_per_sm = 1) {
  // CTAs per launch the pixel range is split for.  One CTA per SM: every split writes a full [taps*C, N] fp32 partial
  // that reduce_partials reads back, so 296 (two waves at one resident CTA per SM) doubled that traffic for nothing -
  // measured on the whole job: 296 -> 104.5, 222 -> 107.2, 148 -> 109.2 tasks/s (MLIIS_WG_TARGET overrides).
  // Task-batched launches (nz slots per launch) split every slot's range nz times less: the target is CTAs per LAUNCH.
  // Measured at 16 slots per launch: 148 splits per slot (2368 CTAs of 2-21 K-blocks each) 122.0 tasks/s, 37 -> 131.7,
  // 19 -> 134.5, 10 -> 135.5.  The summation tree of dW therefore depends on the group size (partition_nz()).
  static int target = -1;
  if (target < 0) { const char* e = getenv("MLIIS_WG_TARGET"); target = e ? atoi(e) : 148; }
  const int wave = 148 * ctas_per_sm;
  int base = ctiles * taps * partition_nz();
  int S = (target == 148 ? wave : target) / base;
  if (S < 1) S = 1;
  const int s_max = tiles_total / 2 < 148 ? tiles_total / 2 : 148;   // at least 2 pixel tiles (64 pixels) per CTA
  if (target == 148 && base * S != wave) {
    // one resident CTA per SM: prefer the split count whose CTAs fill whole waves of 148 (16-slot launch of the decoder
    // wgrad: 96 CTAs of the whole pixel range = 0.65 of one wave, 1607 us; 3 splits = 288 CTAs = 0.97 of two waves)
    auto eff = [&](int s) { const int n = base * s; return (double)n / (double)(((n + wave - 1) / wave) * wave); };
    int best = S;
    for (int c = S + 1; c <= 2 * S + 1 && c <= s_max; ++c)
      if (eff(c) > eff(best) + 0.1) best = c;
    S = best;
  }
  if (S > s_max) S = s_max;
  if (S < 1) S = 1;
  return S;
}

bool tc_wgrad_supported(int conv, int W, int C, int N) {
  if (C % 4 || N % 4 || N > 256) return false;
  if (conv && (W < 1)) return false;
  return encode_fn() != nullptr;
}
size_t tc_wgrad_scratch(int conv, int M, int B, int H, int W, int C, int N, int taps) {
  int tiles;
  if (conv) { int BX = W > 16 ? 32 : 16, BY = 32 / BX; tiles = B * ((W + BX - 1) / BX) * ((H + BY - 1) / BY); }
  else tiles = (M + 31) / 32;
  const int gt = taps == 9 ? 3 : taps;                                // 3x3: the tap-shared grid has 3 filter rows
  const int S1 = wg_splits((C + 127) / 128, gt, tiles, 1), S2 = wg_splits((C + 127) / 128, gt, tiles, 2);
  return (size_t)(S1 > S2 ? S1 : S2) * taps * C * N;                  // small-channel layers run two CTAs per SM
}

// dW[taps*C, N] = sum A^T G ; scratch holds the per-split partials (tc_wgrad_scratch floats)
bool tc_wgrad(const float* A, int lda, const float* G, int ldg, float* dW, float* scratch, int conv, int M, int B, int H,
              int W, int C, int taps, int dil, int N, int split, cudaStream_t s, const float* pa, const float* pb,
              const float* gate, int HW, int dw_tap_stride) {
  TcWgParams p{};
  p.conv = conv; p.M = M; p.H = H; p.W = W; p.C = C; p.N = N; p.taps = taps; p.dil = dil;
  static int dbgw = -1;
  if (dbgw < 0) { const char* e = getenv("MLIIS_TC_DEBUG"); dbgw = e ? atoi(e) : 0; }
  p.debug = dbgw;
  p.split = split == 3 ? 3 : 1;
  if (pa && conv) return false;
  p.pa = pa; p.pb = pb; p.gate = gate; p.HW = HW > 0 ? HW : 1;
  p.BN = (N + 15) / 16 * 16;
  p.NG = (p.BN + 31) / 32;
  CUtensorMap tmA, tmG;
  if (conv) {
    p.BX = W > 16 ? 32 : 16;
    p.BY = 32 / p.BX;
    p.tiles_x = (W + p.BX - 1) / p.BX;
    p.tiles_y = (H + p.BY - 1) / p.BY;
    p.tiles_total = B * p.tiles_x * p.tiles_y;
    static int share_on = -1;
    if (share_on < 0) { const char* e = getenv("MLIIS_TC_WGRAD_SHARE"); share_on = e ? atoi(e) : 1; }
    // tap sharing pays when every CTA still walks a long pixel range (>= 8 K-blocks of 32 pixels); the small 14x14
    // layers keep one tap per CTA (measured: 21 -> 31 us with sharing, three epilogues and a short pipeline per CTA)
    {
      const int ct = (C + 127) / 128, s3 = 296 / (ct * 3) > 0 ? 296 / (ct * 3) : 1;
      p.share = (share_on && taps == 9 && 3 * p.BN <= 512 && p.tiles_total / s3 >= 8) ? 1 : 0;
    }
    cuuint32_t box[5] = {32, (cuuint32_t)p.BX, (cuuint32_t)p.BY, 1, 1};
    cuuint32_t boxA[5] = {32, (cuuint32_t)(p.BX + (p.share ? 2 * dil : 0)), (cuuint32_t)p.BY, 1, 1};
    if (p.share) {
      const int rows = (p.BX + 2 * dil) * p.BY;
      p.a_tx_bytes = 4 * rows * 128;
      p.a_group_bytes = (rows * 128 + 511) / 512 * 512;        // slabs start on swizzle-atom (512 B) boundaries
    }
    cuuint64_t dA[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)MLIIS_NZ};
    cuuint64_t sA[4] = {(cuuint64_t)lda * 4, (cuuint64_t)W * lda * 4, (cuuint64_t)H * W * lda * 4,
                        slot_stride_bytes((cuuint64_t)B * H * W * lda * 4)};
    if (!encode(&tmA, A, 5, dA, sA, boxA, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return false;
    cuuint64_t dG[5] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)MLIIS_NZ};
    cuuint64_t sG[4] = {(cuuint64_t)ldg * 4, (cuuint64_t)W * ldg * 4, (cuuint64_t)H * W * ldg * 4,
                        slot_stride_bytes((cuuint64_t)B * H * W * ldg * 4)};
    if (!encode(&tmG, G, 5, dG, sG, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return false;
  } else {
    p.tiles_total = (M + 31) / 32;
    cuuint32_t box[3] = {32, 32, 1};
    cuuint64_t dA[3] = {(cuuint64_t)C, (cuuint64_t)M, (cuuint64_t)MLIIS_NZ};
    cuuint64_t sA[2] = {(cuuint64_t)lda * 4, slot_stride_bytes((cuuint64_t)M * lda * 4)};
    if (!encode(&tmA, A, 3, dA, sA, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return false;
    cuuint64_t dG[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)MLIIS_NZ};
    cuuint64_t sG[2] = {(cuuint64_t)ldg * 4, slot_stride_bytes((cuuint64_t)M * ldg * 4)};
    if (!encode(&tmG, G, 3, dG, sG, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return false;
  }
  const int ctiles = (C + 127) / 128;
  const int grid_taps = p.share ? 3 : taps;
  const int planes = p.split == 3 ? 2 : 1;
  p.nga_alloc = (!p.share && ctiles == 1) ? ((C + 31) / 32 < 4 ? (C + 31) / 32 : 4) : 4;
  const int stage_bytes = planes * ((p.share ? 4 * p.a_group_bytes : p.nga_alloc * kWgGroupBytes) + p.NG * kWgGroupBytes);
  // compact stages: the MMAs of the last stage read up to 16 KB + one plane past its start (garbage rows, see the kernel)
  int pad = 0;
  if (p.nga_alloc < 4) {
    const int reach = (planes == 2 ? p.nga_alloc * kWgGroupBytes : 0) + 4 * kWgGroupBytes;   // of the lo (hi) plane's MMA
    pad = reach > stage_bytes ? reach - stage_bytes + 1024 : 0;
  }
  // two CTAs per SM when three stages of each fit: two independent load -> transform -> MMA chains hide each other's
  // round trips (C = 32, N = 16 at 112x112, 16 slots: 374 -> 292 us)
  const int ctas_per_sm = (!p.share && 3 * stage_bytes + pad + 2048 <= 112 * 1024) ? 2 : 1;
  p.splits = wg_splits(ctiles, grid_taps, p.tiles_total, ctas_per_sm);
  p.stages = ((ctas_per_sm == 2 ? 110 * 1024 : tc_smem_budget()) - pad) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  const int per = (p.tiles_total + p.splits - 1) / p.splits;
  if (p.stages > per) p.stages = per;
  if (p.stages < 1) return false;
  const size_t smem = (size_t)p.stages * stage_bytes + (3 * p.stages + 2) * 8 + 1024 + pad;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  dim3 grid(ctiles, grid_taps, p.splits * MLIIS_NZ);
  MLIIS_COUNT(), tc_wgrad_kernel<<<grid, kWgThreads, smem, s>>>(tmA, tmG, scratch, p, MLIIS_ZS);
  if (dw_tap_stride > 0 && dw_tap_stride != C * N)
    reduce_partials_strided(scratch, p.splits, C * N, taps, dW, dw_tap_stride, s);
  else
    reduce_partials(scratch, p.splits, taps * C * N, dW, s);
  return true;
}

// =================================================================================================
// Tensor-pipe peak of the mode the convolutions run in: kind::tf32, M = 128, N = 256, K = 8, cta_group::1, operands
// resident in shared memory, no loads, one CTA per SM issuing back to back.  The roofline denominator that belongs to
// these kernels (MEASURED_PEAKS.json only has the cuBLAS bf16 number; TF32 is nominally half of it).
// =================================================================================================
__global__ void __launch_bounds__(128) tc_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar = base + 48 * 1024;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 48 * 1024 + 16);
  for (int i = threadIdx.x; i < 48 * 1024 / 16; i += 128) reinterpret_cast<float4*>(smem)[i] = f4s(0.f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = make_kmajor_sw128_desc(base), db = make_kmajor_sw128_desc(base + 16 * 1024);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem_acc, da + 2 * k, db + 2 * k, idesc, (it | k) ? 1u : 0u);
    tc_commit(bar);
    mbar_wait(bar, 0);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(256u) : "memory");
  }
}
// returns the algorithmic TFLOP/s of `iters` x 4 MMAs per CTA on every SM (CUDA events around the launch, after a warm-up)
double tc_peak_tf32(int iters, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 48 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(tc_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, s);
    MLIIS_COUNT(), tc_peak_kernel<<<sms, 128, smem, s>>>(iters);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = (double)sms * iters * 4.0 * 2.0 * 128 * 256 * 8;
  return flops / (best * 1e-3) / 1e12;
}

// MMA issue-rate microbenchmark: what one SM's tensor pipe sustains for the instruction shapes the convolutions issue
// (operands resident in shared memory, no loads, no transform).  pattern 0: one N = n MMA per k-step; 1: the 3xTF32
// "wide" pair (N = 2n with A_hi, then N = n with A_lo, same accumulator range as tc_conv3_kernel); 2: pattern 1
// alternating between two pixel tiles / accumulators (MT = 2); 3: three N = n MMAs per k-step.  shift_rows moves the A
// descriptor start by whole 128-byte rows (the tap-shifted starts of tc_conv3_kernel).  Result: clocks per k-step.
__global__ void __launch_bounds__(128) tc_rate_kernel(int iters, int n, int pattern, int shift_rows, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int kA = 96 * 1024, kB = 64 * 1024;
  const uint32_t bar = base + kA + kB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kA + kB + 16);
  for (int i = threadIdx.x; i < (kA + kB) / 16; i += 128) reinterpret_cast<float4*>(smem)[i] = f4s(0.f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_w = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 2) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = base + (uint32_t)shift_rows * 128u, a_lo = a_hi + 48 * 1024;
    const uint64_t db = make_kmajor_sw128_desc(base + kA), dbl = make_kmajor_sw128_desc(base + kA + 32 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int tile = pattern == 2 ? (it & 1) : 0;
      const uint64_t da = make_kmajor_sw128_desc(a_hi + (uint32_t)tile * 16384u), dal = make_kmajor_sw128_desc(a_lo + (uint32_t)tile * 16384u);
      const uint32_t acc = tmem_acc + (uint32_t)(tile * 256);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(2 * k);
        if (pattern == 0) {
          tc_mma_tf32(acc, da + adv, db + adv, idesc, (it | k) ? 1u : 0u);
        } else if (pattern == 3) {
          tc_mma_tf32(acc, da + adv, db + adv, idesc, (it | k) ? 1u : 0u);
          tc_mma_tf32(acc, dal + adv, db + adv, idesc, 1u);
          tc_mma_tf32(acc, da + adv, dbl + adv, idesc, 1u);
        } else {
          tc_mma_tf32(acc, da + adv, db + adv, idesc_w, (it > 1 || k) ? 1u : 0u);
          tc_mma_tf32(acc, dal + adv, db + adv, idesc, 1u);
        }
      }
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (float)((double)(t1 - t0) / ((double)iters * 4.0));
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512u) : "memory");
  }
}
// clocks per k-step (K = 8) of the given MMA pattern on every SM at once; < 0 on a bad argument
double tc_mma_rate(int iters, int n, int pattern, int shift_rows, cudaStream_t s) {
  if (n < 8 || n > 256 || (n & 7) || pattern < 0 || pattern > 3 || shift_rows < 0 || shift_rows > 64) return -1.0;
  if ((pattern == 1 || pattern == 2) && 2 * n > 256) return -1.0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 160 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float* d = nullptr;
  if (cudaMalloc(&d, sizeof(float)) != cudaSuccess) return -1.0;
  float h = -1.f;
  for (int rep = 0; rep < 2; ++rep) {
    MLIIS_COUNT(), tc_rate_kernel<<<sms, 128, smem, s>>>(iters, n, pattern, shift_rows, d);
    cudaStreamSynchronize(s);
  }
  cudaMemcpy(&h, d, sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return (double)h;
}

// weights W[tap][ci][co] (HWIO) -> forward operand Wt[co][tap][ci]   or   dgrad operand Wt[ci][taps-1-tap][co],
// rounded to nearest TF32; split == 3 also writes the residual plane lo = rn(w - hi) behind the hi plane.
__device__ __forceinline__ void prep_one(const float* __restrict__ w, float* __restrict__ wt, int i, int n, int taps,
                                         int Ci, int Co, int dgrad, int split, int Cs) {
  float v;
  if (!dgrad) {
    const int co = i / (taps * Ci), rem = i - co * taps * Ci, tap = rem / Ci, ci = rem - tap * Ci;
    v = w[((size_t)tap * Cs + ci) * Co + co];
  } else {
    const int ci = i / (taps * Co), rem = i - ci * taps * Co, tap = rem / Co, co = rem - tap * Co;
    v = w[((size_t)(taps - 1 - tap) * Cs + ci) * Co + co];
  }
  const float h = rn_tf32(v);
  wt[i] = h;
  if (split == 3) wt[(size_t)n + i] = rn_tf32(v - h);
}
__global__ void tc_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int taps, int Ci, int Co,
                                       int dgrad, int split, int Cs, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; w += zo; wt += zo; }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = taps * Ci * Co;
  if (i < n) prep_one(w, wt, i, n, taps, Ci, Co, dgrad, split, Cs);
}
// every dense layer's forward and dgrad operand in ONE launch per step (blockIdx.y = job)
__global__ void tc_prep_all_kernel(const float* __restrict__ theta, float* __restrict__ wcache,
                                   const TcPrepJob* __restrict__ jobs, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; theta += zo; wcache += zo; }
  const TcPrepJob j = jobs[blockIdx.y];
  const int n = j.taps * j.Ci * j.Co;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    prep_one(theta + j.w_off, wcache + j.dst, i, n, j.taps, j.Ci, j.Co, j.dgrad, j.split, j.Cs > 0 ? j.Cs : j.Ci);
}
void tc_prep_all(const float* theta, float* wcache, const TcPrepJob* dev_jobs, int n_jobs, cudaStream_t s) {
  if (n_jobs <= 0) return;
  MLIIS_COUNT(), tc_prep_all_kernel<<<dim3(24, n_jobs, MLIIS_NZ), 256, 0, s>>>(theta, wcache, dev_jobs, MLIIS_ZS);
}
void tc_prep_weights(const float* w, float* wt, int taps, int Ci, int Co, int dgrad, int split, cudaStream_t s, int Cs) {
  const int n = taps * Ci * Co;
  MLIIS_COUNT(), tc_prep_weights_kernel<<<dim3(cdiv(n, 256), 1, MLIIS_NZ), 256, 0, s>>>(w, wt, taps, Ci, Co, dgrad, split,
                                                                                       Cs > 0 ? Cs : Ci, MLIIS_ZS);
}

}  // namespace mliis
