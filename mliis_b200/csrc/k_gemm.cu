// Dense contractions, fp32 FFMA path (MLIIS_GEMM_FP32): 1x1 convs, implicit-GEMM 3x3 (dilated) convs,
// their dgrad (same kernel on flipped/transposed weights) and wgrad (TN kernel with split-M partials
// reduced in a fixed order -> deterministic).  This path is the exact-order numeric reference for the
// tcgen05 kernels in k_tc.cu and is what parity tests run first.
//
// A-operand prologue fusion: the MBConv project conv consumes swish(BN(dw))*gate; that tensor is
// recomputed while the A tile is loaded (efficientnet_model.py:271-280) and never written to HBM.
#include "common.cuh"
#include "kernels.h"

namespace mliis {

struct ARow {   // per-thread row state for the A loader
  bool valid;
  int img, y, x;
  size_t base;  // plain mode: m*ld
};

template <int PRO, int CONV>
__device__ __forceinline__ float4 load_a(const GemmA& A, const ARow& r, int k, int K) {
  // k is a multiple of 4 and k < K
  const float* p;
  if (CONV) {
    const int tap = k / A.C, c = k - tap * A.C;
    const int ty = tap / 3, tx = tap - ty * 3;
    const int yy = r.y + (ty - 1) * A.dil, xx = r.x + (tx - 1) * A.dil;
    if (!r.valid || yy < 0 || yy >= A.H || xx < 0 || xx >= A.W) return f4s(0.f);
    p = A.ptr + (((size_t)r.img * A.H + yy) * A.W + xx) * A.ld + c;
  } else {
    if (!r.valid) return f4s(0.f);
    p = A.ptr + r.base + k;
  }
  float4 v = ld4(p);
  if (PRO) {
    v = swish4(affine4(v, ld4(A.pa + k), ld4(A.pb + k)));
    if (A.gate) v = v * ld4(A.gate + (size_t)r.img * K + k);
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// NN: C[M,N] = A[M,K] * W[K,N]      tile 128x64x8, 256 threads, 8x4 per thread
// ---------------------------------------------------------------------------------------------
template <int PRO, int CONV>
__global__ void __launch_bounds__(256) gemm_nn_kernel(GemmA A, const float* __restrict__ Wt,
                                                       const float* __restrict__ bias, float* __restrict__ Cout,
                                                       int ldc, int M, int K, int N, int HW, int accumulate) {
  constexpr int BM = 128, BN = 64, BK = 8, LDA = BM + 4;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int arow = tid >> 1, akq = (tid & 1) * 4;
  ARow r;
  {
    const int m = m0 + arow;
    r.valid = m < M;
    r.img = r.valid ? m / HW : 0;
    const int rem = m - r.img * HW;
    r.y = CONV ? rem / A.W : 0;
    r.x = CONV ? rem - r.y * A.W : 0;
    r.base = (size_t)m * A.ld;
  }
  const int bkr = tid >> 4, bnq = (tid & 15) * 4;   // B loader (threads 0..127)
  const bool bload = tid < 128;
  const bool bvalid = bload && (n0 + bnq < N);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int KT = K / BK;
  float4 ra = load_a<PRO, CONV>(A, r, akq, K);
  float4 rb = bvalid ? ld4(Wt + (size_t)bkr * N + n0 + bnq) : f4s(0.f);
  As[0][akq + 0][arow] = ra.x; As[0][akq + 1][arow] = ra.y; As[0][akq + 2][arow] = ra.z; As[0][akq + 3][arow] = ra.w;
  if (bload) st4(&Bs[0][bkr][bnq], rb);
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < KT; ++kt) {
    if (kt + 1 < KT) {
      const int k0 = (kt + 1) * BK;
      ra = load_a<PRO, CONV>(A, r, k0 + akq, K);
      rb = bvalid ? ld4(Wt + (size_t)(k0 + bkr) * N + n0 + bnq) : f4s(0.f);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = ld4(&As[cur][k][ty * 8]), a1 = ld4(&As[cur][k][ty * 8 + 4]);
      const float4 b0 = ld4(&Bs[cur][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < KT) {
      const int nx = cur ^ 1;
      As[nx][akq + 0][arow] = ra.x; As[nx][akq + 1][arow] = ra.y; As[nx][akq + 2][arow] = ra.z; As[nx][akq + 3][arow] = ra.w;
      if (bload) st4(&Bs[nx][bkr][bnq], rb);
    }
    __syncthreads();
    cur ^= 1;
  }
  const int n = n0 + tx * 4;
  if (n < N) {
    float4 bz = bias ? ld4(bias + n) : f4s(0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ty * 8 + i;
      if (m < M) {
        float* o = Cout + (size_t)m * ldc + n;
        float4 v = f4(acc[i][0] + bz.x, acc[i][1] + bz.y, acc[i][2] + bz.z, acc[i][3] + bz.w);
        if (accumulate) v = v + ld4(o);
        st4(o, v);
      }
    }
  }
}

void gemm_nn(const GemmA& A, const float* Wt, const float* bias, float* Cout, int ldc, int M, int K, int N, int HW,
             int accumulate, cudaStream_t s) {
  dim3 grid(cdiv(M, 128), cdiv(N, 64));
  const bool pro = A.pa != nullptr;
  if (A.conv) {
    MLIIS_COUNT(), gemm_nn_kernel<0, 1><<<grid, 256, 0, s>>>(A, Wt, bias, Cout, ldc, M, K, N, HW, accumulate);
  } else if (pro) {
    MLIIS_COUNT(), gemm_nn_kernel<1, 0><<<grid, 256, 0, s>>>(A, Wt, bias, Cout, ldc, M, K, N, HW, accumulate);
  } else {
    MLIIS_COUNT(), gemm_nn_kernel<0, 0><<<grid, 256, 0, s>>>(A, Wt, bias, Cout, ldc, M, K, N, HW, accumulate);
  }
}

// ---------------------------------------------------------------------------------------------
// TN (wgrad): dW[k,n] = sum_m A[m,k] * G[m,n]    tile 128(k) x 128(n), reduction step 8 rows, 8x8/thread
// grid = (k tiles, n tiles, taps*splits); partial[split][Ktot][N]; bias partial[split][N] after them.
// ---------------------------------------------------------------------------------------------
static inline int tn_splits(int M, int K, int N, int conv) {
  const int kdim = conv ? K / 9 : K;
  const int base = cdiv(kdim, 128) * cdiv(N, 128) * (conv ? 9 : 1);
  int S = 296 / base;
  if (S < 1) S = 1;
  const int maxS = cdiv(M, 256);
  if (S > maxS) S = maxS;
  if (S > 148) S = 148;
  if (S < 1) S = 1;
  return S;
}
size_t gemm_tn_scratch(int M, int K, int N, int conv) {
  const int S = tn_splits(M, K, N, conv);
  return (size_t)S * ((size_t)K * N + N);
}

template <int PRO, int CONV>
__global__ void __launch_bounds__(256) gemm_tn_kernel(GemmA A, const float* __restrict__ G, int ldg,
                                                       float* __restrict__ partial, float* __restrict__ bpartial,
                                                       int M, int Ktot, int N, int HW, int rows_per_split,
                                                       int splits) {
  constexpr int BT = 128, BR = 8;
  __shared__ __align__(16) float As[2][BR][BT];
  __shared__ __align__(16) float Bs[2][BR][BT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int Kdim = CONV ? A.C : Ktot;
  const int tap = CONV ? blockIdx.z / splits : 0;
  const int split = CONV ? blockIdx.z - tap * splits : blockIdx.z;
  const int k0 = blockIdx.x * BT, n0 = blockIdx.y * BT;
  const int mbeg = split * rows_per_split, mend = min(M, mbeg + rows_per_split);
  const int lr = tid >> 5, lq = (tid & 31) * 4;   // loader: row lr of the 8-row slab, float4 column lq
  const bool akvalid = k0 + lq < Kdim, bnvalid = n0 + lq < N;
  const int dty = CONV ? (tap / 3 - 1) * A.dil : 0, dtx = CONV ? (tap % 3 - 1) * A.dil : 0;
  const bool want_bias = bpartial != nullptr && blockIdx.x == 0 && tap == 0;

  auto loadA = [&](int m) -> float4 {
    if (m >= mend || !akvalid) return f4s(0.f);
    const int img = m / HW;
    const float* p;
    if (CONV) {
      const int rem = m - img * HW, y = rem / A.W, x = rem - y * A.W;
      const int yy = y + dty, xx = x + dtx;
      if (yy < 0 || yy >= A.H || xx < 0 || xx >= A.W) return f4s(0.f);
      p = A.ptr + (((size_t)img * A.H + yy) * A.W + xx) * A.ld + k0 + lq;
    } else {
      p = A.ptr + (size_t)m * A.ld + k0 + lq;
    }
    float4 v = ld4(p);
    if (PRO) {
      v = swish4(affine4(v, ld4(A.pa + k0 + lq), ld4(A.pb + k0 + lq)));
      if (A.gate) v = v * ld4(A.gate + (size_t)img * Kdim + k0 + lq);
    }
    return v;
  };
  auto loadB = [&](int m) -> float4 {
    if (m >= mend || !bnvalid) return f4s(0.f);
    return ld4(G + (size_t)m * ldg + n0 + lq);
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;

  float4 ra = loadA(mbeg + lr), rb = loadB(mbeg + lr);
  st4(&As[0][lr][lq], ra);
  st4(&Bs[0][lr][lq], rb);
  __syncthreads();
  int cur = 0;
  for (int m = mbeg; m < mend; m += BR) {
    const bool more = m + BR < mend;
    if (more) { ra = loadA(m + BR + lr); rb = loadB(m + BR + lr); }
#pragma unroll
    for (int rr = 0; rr < BR; ++rr) {
      const float4 a0 = ld4(&As[cur][rr][ty * 4]), a1 = ld4(&As[cur][rr][64 + ty * 4]);
      const float4 b0 = ld4(&Bs[cur][rr][tx * 4]), b1 = ld4(&Bs[cur][rr][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (want_bias && ty == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bsum[j] += bv[j];
      }
    }
    if (more) {
      st4(&As[cur ^ 1][lr][lq], ra);
      st4(&Bs[cur ^ 1][lr][lq], rb);
    }
    __syncthreads();
    cur ^= 1;
  }
  float* po = partial + (size_t)split * Ktot * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kk = k0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (kk >= Kdim) continue;
    const size_t krow = (size_t)(tap * Kdim + kk) * N;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n < N) st4(po + krow + n, f4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]));
    }
  }
  if (want_bias && ty == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n < N) st4(bpartial + (size_t)split * N + n, f4(bsum[h * 4 + 0], bsum[h * 4 + 1], bsum[h * 4 + 2], bsum[h * 4 + 3]));
    }
  }
}

void gemm_tn(const GemmA& A, const float* G, int ldg, float* dW, float* dbias, float* scratch, int M, int K, int N,
             int HW, cudaStream_t s) {
  const int S = tn_splits(M, K, N, A.conv);
  int rps = cdiv(M, S);
  rps = cdiv(rps, 8) * 8;
  const int kdim = A.conv ? K / 9 : K;
  dim3 grid(cdiv(kdim, 128), cdiv(N, 128), S * (A.conv ? 9 : 1));
  float* partial = scratch;
  float* bpartial = dbias ? scratch + (size_t)S * K * N : nullptr;
  const bool pro = A.pa != nullptr;
  if (A.conv)
    MLIIS_COUNT(), gemm_tn_kernel<0, 1><<<grid, 256, 0, s>>>(A, G, ldg, partial, bpartial, M, K, N, HW, rps, S);
  else if (pro)
    MLIIS_COUNT(), gemm_tn_kernel<1, 0><<<grid, 256, 0, s>>>(A, G, ldg, partial, bpartial, M, K, N, HW, rps, S);
  else
    MLIIS_COUNT(), gemm_tn_kernel<0, 0><<<grid, 256, 0, s>>>(A, G, ldg, partial, bpartial, M, K, N, HW, rps, S);
  reduce_partials(partial, S, K * N, dW, s);
  if (dbias) reduce_partials(bpartial, S, N, dbias, s);
}

// ---------------------------------------------------------------------------------------------
// small weight reshuffles for dgrad
// ---------------------------------------------------------------------------------------------
__global__ void transpose_w_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int N, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; w += zo; wt += zo; }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  int n = i / K, k = i - n * K;   // wt[n][k]
  wt[i] = w[(size_t)k * N + n];
}
void transpose_w(const float* w, float* wt, int K, int N, cudaStream_t s) {
  MLIIS_COUNT(), transpose_w_kernel<<<dim3(cdiv(K * N, 256), 1, MLIIS_NZ), 256, 0, s>>>(w, wt, K, N, MLIIS_ZS);
}
// wt[tap][n][c] = w[8-tap][c][n]
__global__ void flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int C, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C * N) return;
  int tap = i / (C * N), rem = i - tap * C * N, n = rem / C, c = rem - n * C;
  wt[i] = w[((size_t)(8 - tap) * C + c) * N + n];
}
void flip_transpose_w3x3(const float* w, float* wt, int C, int N, cudaStream_t s) {
  MLIIS_COUNT(), flip_transpose_kernel<<<cdiv(9 * C * N, 256), 256, 0, s>>>(w, wt, C, N);
}

}  // namespace mliis
