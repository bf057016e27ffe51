// Image-pooling branch of the residual skip decoder, FOLDED into its consumer (models/efficientlab.py:192-197,
// :220-224): `pool_image_features` tiles the per-image channel mean of `concat` over H x W and feeds it, as the last
// catC channels of the pyramid, to the 3x3 `conv2d_2`.  That tile is a per-image constant, so its contribution to
// conv2d_2 is a per-image vector per filter tap, summed over the taps that fall inside the image - with TF's zero
// SAME padding that depends only on the BORDER CLASS of the output pixel (3 row classes x 3 column classes):
//
//   forward   out[b,y,x,n] += bias9[b][cls(y,x)][n],   bias9[b][cls][n] = sum_{tap valid in cls} tap9[b][tap][n],
//             tap9[b][tap][n] = sum_c p[b,c] W[tap][2D+c][n]   (pool_taps_kernel; the class sums: tc_conv3_kernel)
//   wgrad     dW[tap][2D+c][n] = sum_b p[b,c] S[b][tap][n],   S[b][tap][n] = sum_{cls where tap valid} Q[b][cls][n],
//             Q[b][cls][n] = sum of the output gradient over the pixels of class cls
//   dgrad     dp[b,c] = sum_{tap,n} S[b][tap][n] W[tap][2D+c][n]       (then / HW onto every pixel of `concat`)
//
// so the 136 (224) pooled channels never enter the implicit GEMM: 38 % (50 %) of the K dimension of the layer that
// holds 57 % of the network's FLOPs disappears from forward, dgrad and wgrad.  Everything here is fp32 FFMA with a
// fixed summation order (deterministic).
#include "common.cuh"
#include "kernels.h"

namespace mliis {

// tap (ty,tx) reads input pixel (y + (ty-1)*dil, x + (tx-1)*dil).  Row class 0: y < dil (the ty = 0 taps read above the
// image), 2: y >= H - dil (ty = 2 reads below), 1: interior; same for columns.  cls = 3*ry + rx.
__device__ __forceinline__ bool tap_valid(int tap, int cls) {
  const int ty = tap / 3, tx = tap - ty * 3, ry = cls / 3, rx = cls - ry * 3;
  return !((ty == 0 && ry == 0) || (ty == 2 && ry == 2) || (tx == 0 && rx == 0) || (tx == 2 && rx == 2));
}

// Per-tap vectors of the folded branch: tap9[b][tap][n] = sum_c p[b,c] W[tap][c_first + c][n].  grid (9 taps, B);
// block (D, 8): threadIdx.y splits the pooled channels (<= Cp/8 coalesced weight loads per thread), fixed-order combine
// through shared memory.  The border-class sums bias9[cls] = sum of the taps valid in cls are formed by the consumer
// (tc_conv3_kernel prologue, in shared memory), so this launch is 72 small CTAs instead of 8 long ones.
__global__ void __launch_bounds__(1024) pool_taps_kernel(const float* __restrict__ pooled, int ldp,
                                                         const float* __restrict__ w, int Cs, int c_first, int Cp, int D,
                                                         float* __restrict__ tap9, long long zs) {
  extern __shared__ float smf[];       // partial[P][D] | p[Cp]
  { const size_t zo = (size_t)blockIdx.z * zs; pooled += zo; w += zo; tap9 += zo; }
  const int tap = blockIdx.x, b = blockIdx.y, n = threadIdx.x, part = threadIdx.y, P = blockDim.y;
  float* ps = smf + P * D;
  for (int c = part * D + n; c < Cp; c += P * D) ps[c] = pooled[(size_t)b * ldp + c];
  __syncthreads();
  const float* wr = w + ((size_t)tap * Cs + c_first) * D + n;
  float s = 0.f;
  for (int c = part; c < Cp; c += P) s = fmaf(ps[c], wr[(size_t)c * D], s);
  smf[part * D + n] = s;
  __syncthreads();
  if (part == 0) {
    for (int j = 1; j < P; ++j) s += smf[j * D + n];
    tap9[((size_t)b * 9 + tap) * D + n] = s;
  }
}

void pool_bias9(const float* pooled, int ldp, const float* w_hwio, int Cs, int c_first, int Cp, int D, int B,
                float* tap9, cudaStream_t s) {
  int P = 1024 / D;
  if (P > 8) P = 8;
  if (P < 1) P = 1;
  MLIIS_COUNT(), pool_taps_kernel<<<dim3(9, B, MLIIS_NZ), dim3(D, P), (P * D + Cp) * sizeof(float), s>>>(pooled, ldp, w_hwio, Cs, c_first,
                                                                                                    Cp, D, tap9, MLIIS_ZS);
}

// Q partials: grid (G row chunks, B); block (D/4, R).  partial[b][g][cls][D]
__global__ void region_sums_kernel(const float* __restrict__ g, int ldg, int H, int W, int dil, int D, int rows_per_chunk,
                                   float* __restrict__ partial, long long zs) {
  extern __shared__ float4 sm[];
  { const size_t zo = (size_t)blockIdx.z * zs; g += zo; partial += zo; }
  const int cq = threadIdx.x, C4 = blockDim.x, R = blockDim.y, ty = threadIdx.y, b = blockIdx.y, G = gridDim.x;
  const int HW = H * W;
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  float4 acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = f4s(0.f);
  for (int r = r0 + ty; r < r1; r += R) {
    const int y = r / W, x = r - y * W;
    const int cls = (y < dil ? 0 : (y >= H - dil ? 2 : 1)) * 3 + (x < dil ? 0 : (x >= W - dil ? 2 : 1));
    const float4 v = ld4(g + ((size_t)b * HW + r) * ldg + cq * 4);
#pragma unroll
    for (int k = 0; k < 9; ++k)
      if (k == cls) acc[k] = acc[k] + v;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) sm[(k * R + ty) * C4 + cq] = acc[k];
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float4 s4 = acc[k];
      for (int j = 1; j < R; ++j) s4 = s4 + sm[(k * R + j) * C4 + cq];
      st4(partial + (((size_t)b * G + blockIdx.x) * 9 + k) * D + cq * 4, s4);
    }
  }
}

// S[b][tap][n] = sum over the classes in which the tap is valid of Q[b][cls][n];  grid B, block (D, 4): lane y sums the
// chunks y, y + 4, ... of the nine class partials (nine independent loads in flight), the lanes are combined in order
__global__ void __launch_bounds__(1024) region_sums_finalize_kernel(const float* __restrict__ partial, int G, int D,
                                                                    float* __restrict__ S, long long zs) {
  extern __shared__ double sdq[];      // [L][D][9]
  { const size_t zo = (size_t)blockIdx.z * zs; partial += zo; S += zo; }
  const int b = blockIdx.x, n = threadIdx.x, ly = threadIdx.y, L = blockDim.y;
  double q[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) q[k] = 0.0;
  for (int gch = ly; gch < G; gch += L) {
    float v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = partial[(((size_t)b * G + gch) * 9 + k) * D + n];
#pragma unroll
    for (int k = 0; k < 9; ++k) q[k] += (double)v[k];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) sdq[((size_t)ly * D + n) * 9 + k] = q[k];
  __syncthreads();
  if (ly != 0) return;
  for (int j = 1; j < L; ++j)
#pragma unroll
    for (int k = 0; k < 9; ++k) q[k] += sdq[((size_t)j * D + n) * 9 + k];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k)
      if (tap_valid(tap, k)) s += q[k];
    S[((size_t)b * 9 + tap) * D + n] = (float)s;
  }
}

int region_sums_chunks(int HW, int D) {
  int R = 256 / (D / 4);
  if (R < 1) R = 1;
  if (R > 16) R = 16;
  int G = cdiv(HW, R * 8);
  if (G > 32) G = 32;
  if (G < 1) G = 1;
  return G;
}

void region_sums(const float* g, int ldg, int B, int H, int W, int dil, int D, float* partial, float* S, cudaStream_t s) {
  int R = 256 / (D / 4);
  if (R < 1) R = 1;
  if (R > 16) R = 16;
  const int G = region_sums_chunks(H * W, D);
  dim3 blk(D / 4, R);
  MLIIS_COUNT(), region_sums_kernel<<<dim3(G, B, MLIIS_NZ), blk, 9 * blk.x * blk.y * sizeof(float4), s>>>(g, ldg, H, W, dil, D,
                                                                                                       cdiv(H * W, G), partial, MLIIS_ZS);
  int L = 1024 / D;
  if (L > 4) L = 4;                      // L * D * 72 bytes of shared memory: stays under the 48 KB default
  if (L < 1) L = 1;
  MLIIS_COUNT(), region_sums_finalize_kernel<<<dim3(B, 1, MLIIS_NZ), dim3(D, L), (size_t)L * D * 9 * sizeof(double), s>>>(partial, G, D, S,
                                                                                                                   MLIIS_ZS);
}

// dW[tap][c_first + c][n] = sum_b pooled[b][c] * S[b][tap][n];  grid (ceil(Cp/8), 9); block (D, 8)
__global__ void pool_wgrad_kernel(const float* __restrict__ pooled, int ldp, const float* __restrict__ S, int B, int Cs,
                                  int c_first, int Cp, int D, float* __restrict__ dw, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; pooled += zo; S += zo; dw += zo; }
  const int n = threadIdx.x, c = blockIdx.x * blockDim.y + threadIdx.y, tap = blockIdx.y;
  if (c >= Cp) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s = fmaf(pooled[(size_t)b * ldp + c], S[((size_t)b * 9 + tap) * D + n], s);
  dw[((size_t)tap * Cs + c_first + c) * D + n] = s;
}
void pool_wgrad(const float* pooled, int ldp, const float* S, int B, int Cs, int c_first, int Cp, int D, float* dw,
                cudaStream_t s) {
  MLIIS_COUNT(), pool_wgrad_kernel<<<dim3(cdiv(Cp, 8), 9, MLIIS_NZ), dim3(D, 8), 0, s>>>(pooled, ldp, S, B, Cs, c_first, Cp, D, dw,
                                                                                        MLIIS_ZS);
}

// dpooled[b][c] = scale * sum_{tap,n} S[b][tap][n] * W[tap][c_first + c][n];  grid (ceil(Cp/8), B); one warp per c
__global__ void __launch_bounds__(256) pool_dgrad_kernel(const float* __restrict__ S, const float* __restrict__ w, int Cs,
                                                         int c_first, int Cp, int D, float scale,
                                                         float* __restrict__ dpooled, int ldo, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; S += zo; w += zo; dpooled += zo; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp, b = blockIdx.y;
  if (c >= Cp) return;
  float s = 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    const float* wr = w + ((size_t)tap * Cs + c_first + c) * D;
    const float* sr = S + ((size_t)b * 9 + tap) * D;
    for (int n = lane; n < D; n += 32) s = fmaf(sr[n], wr[n], s);
  }
  s = warp_sum(s);
  if (lane == 0) dpooled[(size_t)b * ldo + c] = s * scale;
}
void pool_dgrad(const float* S, const float* w_hwio, int Cs, int c_first, int Cp, int D, int B, float scale,
                float* dpooled, int ldo, cudaStream_t s) {
  MLIIS_COUNT(), pool_dgrad_kernel<<<dim3(cdiv(Cp, 8), B, MLIIS_NZ), 256, 0, s>>>(S, w_hwio, Cs, c_first, Cp, D, scale, dpooled, ldo,
                                                                                 MLIIS_ZS);
}

}  // namespace mliis
