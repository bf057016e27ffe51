// Row-channel kernels: everything that walks an NHWC tensor as [rows, C] once or twice and is bound
// by HBM bandwidth - train-mode BatchNorm statistics / normalise / backward, swish, squeeze-excite
// pooling and gating gradients, residual + drop-connect, concat plumbing.
//
// Thread mapping: blockDim = (C/4, R).  threadIdx.x owns one float4 of channels, so a warp reads
// consecutive 16-byte words of consecutive rows (fully coalesced for any C % 4 == 0, ld == C) and the
// per-channel coefficients live in registers for the whole kernel.  Reductions over rows are done
// thread-serially, then across threadIdx.y through shared memory in a fixed order, then across blocks
// by a tiny finalize kernel in double precision -> deterministic, no float atomics.
//
// Reference semantics: models/efficientnet/utils.py:87-134 (non-fused BN), efficientnet_model.py:238-290
// (SE, block tail), models/efficientlab.py:185-197 (decoder conv->swish->BN, pooled features).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace mliis {

constexpr int kRcCluster = 8;     // CTAs per cluster in the whole-tensor reductions (portable maximum)

static inline int rc_R(int C) {
  int c4 = C / 4;
  int R = 256 / c4;
  if (R < 1) R = 1;
  if (R > 64) R = 64;
  return R;
}
static inline dim3 rc_block(int C) { return dim3(C / 4, rc_R(C)); }

// Row chunks = CTAs (and partial rows) per slot.  Task-batched launches (nz slots per launch) keep about the same number
// of CTAs per LAUNCH - 1184 = 8 resident 256-thread CTAs per SM - so every slot gets nz times fewer, longer chunks and
// its finalize kernel reads nz times fewer partials (partition_nz(), kernels.h).
int rc_num_chunks(int M, int C) {
  int R = rc_R(C);
  int G = cdiv(M, R * 8);
  int cap = 1184 / partition_nz();
  if (cap > 296) cap = 296;
  if (cap < 37) cap = 37;
  if (G > cap) G = cap;
  if (G < 1) G = 1;
  return G;
}
int rc_num_img_chunks(int HW, int C) {
  int R = rc_R(C);
  int G = cdiv(HW, R * 8);
  int cap = 128 / partition_nz();
  if (cap > 32) cap = 32;
  if (cap < 8) cap = 8;
  if (G > cap) G = cap;
  if (G < 1) G = 1;
  return G;
}

// block-level reduction of two float4 accumulators across threadIdx.y (fixed order)
__device__ __forceinline__ void block_reduce2(float4& s0, float4& s1, float4* sm) {
  const int C4 = blockDim.x, R = blockDim.y, cq = threadIdx.x, ty = threadIdx.y;
  sm[ty * C4 + cq] = s0;
  sm[(R + ty) * C4 + cq] = s1;
  __syncthreads();
  if (ty == 0) {
    for (int j = 1; j < R; ++j) {
      s0 = s0 + sm[j * C4 + cq];
      s1 = s1 + sm[(R + j) * C4 + cq];
    }
  }
}
__device__ __forceinline__ void block_reduce1(float4& s0, float4* sm) {
  const int C4 = blockDim.x, R = blockDim.y, cq = threadIdx.x, ty = threadIdx.y;
  sm[ty * C4 + cq] = s0;
  __syncthreads();
  if (ty == 0)
    for (int j = 1; j < R; ++j) s0 = s0 + sm[j * C4 + cq];
}

// Whole-tensor reduction of two per-channel sums WITHOUT a second launch (round 2 experiment, opt-in: see bn_clustered();
// the separate finalize kernels are 9-13 us of pure latency each, ~80 per step).  The row chunks are launched as clusters of 8 CTAs:
//   1. every CTA leaves its (s0, s1) float4 per channel quad in its own shared memory; cluster barrier;
//   2. rank 0 adds the eight partials through distributed shared memory (rank order, double) and writes ONE partial row
//      per cluster to global memory; cluster barrier (the peers' shared memory may go away now);
//   3. rank 0 takes a ticket; the CTA that draws the last one sums the n_clusters rows - thread (cq, ty) takes rows
//      ty, ty + R, ... with four rows in flight, the R lanes are combined in ty order through shared memory, in double.
// Fixed orders everywhere: deterministic.  The ticket resets itself for the next launch on the stream.
// Returns true in the finalizing CTA; there the threads with threadIdx.y == 0 hold the totals of their quad in t0 / t1.
// `sm` must hold blockDim.x * blockDim.y * 64 bytes.
__device__ __forceinline__ bool cluster_total2(float4 s0, float4 s1, float4* sm, float* __restrict__ partials, int C,
                                               unsigned* ticket, double (&t0)[4], double (&t1)[4]) {
  __shared__ int s_last;
  cg::cluster_group cluster = cg::this_cluster();
  const int C4 = blockDim.x, R = blockDim.y, cq = threadIdx.x, ty = threadIdx.y;
  const unsigned rank = cluster.block_rank();
  const int cluster_id = blockIdx.x / kRcCluster, n_clusters = gridDim.x / kRcCluster;
  __syncthreads();                                    // block_reduce2 has finished reading sm
  if (ty == 0) { sm[cq] = s0; sm[C4 + cq] = s1; }
  cluster.sync();
  if (rank == 0 && ty == 0) {
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < kRcCluster; ++r) {
      const float4* remote = cluster.map_shared_rank(sm, r);
      const float4 u = remote[cq], v = remote[C4 + cq];
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
      b[0] += v.x; b[1] += v.y; b[2] += v.z; b[3] += v.w;
    }
    st4(partials + ((size_t)cluster_id * 2 + 0) * C + cq * 4, f4((float)a[0], (float)a[1], (float)a[2], (float)a[3]));
    st4(partials + ((size_t)cluster_id * 2 + 1) * C + cq * 4, f4((float)b[0], (float)b[1], (float)b[2], (float)b[3]));
  }
  cluster.sync();
  if (rank != 0) return false;
  __threadfence();                                    // this CTA's partial row is visible device-wide before the ticket
  __syncthreads();
  if (cq == 0 && ty == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == (unsigned)n_clusters - 1u);
    if (s_last) atomicExch(ticket, 0u);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  for (int g = ty; g < n_clusters; g += 4 * R) {      // four rows (8 float4) in flight per thread; +0.0 tails are exact
    float4 u[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gi = g + k * R;
      u[k] = gi < n_clusters ? __ldcg(reinterpret_cast<const float4*>(partials + ((size_t)gi * 2 + 0) * C + cq * 4)) : f4s(0.f);
      v[k] = gi < n_clusters ? __ldcg(reinterpret_cast<const float4*>(partials + ((size_t)gi * 2 + 1) * C + cq * 4)) : f4s(0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a[0] += u[k].x; a[1] += u[k].y; a[2] += u[k].z; a[3] += u[k].w;
      b[0] += v[k].x; b[1] += v[k].y; b[2] += v[k].z; b[3] += v[k].w;
    }
  }
  double* sd = reinterpret_cast<double*>(sm);
  double* mine = sd + ((size_t)ty * C4 + cq) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) { mine[i] = a[i]; mine[4 + i] = b[i]; }
  __syncthreads();
  if (ty != 0) return true;
#pragma unroll
  for (int i = 0; i < 4; ++i) { t0[i] = a[i]; t1[i] = b[i]; }
  for (int j = 1; j < R; ++j) {
    const double* o = sd + ((size_t)j * C4 + cq) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) { t0[i] += o[i]; t1[i] += o[4 + i]; }
  }
  return true;
}

// launch helper: grid.x = row chunks rounded up to whole clusters (the padding CTAs see no rows and add zeros)
template <typename Kern, typename... Args>
static void launch_clustered(Kern kern, int G, int nz, dim3 blk, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((G + kRcCluster - 1) / kRcCluster * kRcCluster), 1, (unsigned)nz);
  cfg.blockDim = blk;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kRcCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, args...);
}

// The same finalize-in-the-reduction WITHOUT clusters (task-batched launches, where every slot has few, long row chunks):
// every CTA writes its partial row, takes a ticket, and the CTA that draws the slot's last ticket sums the gridDim.x
// rows (thread (cq, ty): rows ty, ty + R, .. with eight rows in flight; lanes combined in ty order, in double: a fixed
// order, whichever CTA comes last).  Returns true in that CTA; threads with threadIdx.y == 0 then hold the totals.
__device__ __forceinline__ bool ticket_total2(float4 s0, float4 s1, float4* sm, float* __restrict__ partials, int C,
                                              unsigned* ticket, double (&t0)[4], double (&t1)[4]) {
  __shared__ int s_last;
  const int C4 = blockDim.x, R = blockDim.y, cq = threadIdx.x, ty = threadIdx.y;
  const int n_rows = gridDim.x;
  if (ty == 0) {
    st4(partials + ((size_t)blockIdx.x * 2 + 0) * C + cq * 4, s0);
    st4(partials + ((size_t)blockIdx.x * 2 + 1) * C + cq * 4, s1);
  }
  __threadfence();                                    // this CTA's partial row is visible device-wide before the ticket
  __syncthreads();                                    // (and block_reduce2 has finished reading sm)
  if (cq == 0 && ty == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == (unsigned)n_rows - 1u);
    if (s_last) atomicExch(ticket, 0u);               // self-resetting: the next launch on the stream starts from zero
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  for (int g = ty; g < n_rows; g += 8 * R) {          // eight rows (16 float4) in flight per thread; +0.0 tails are exact
    float4 u[8], v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int gi = g + k * R;
      u[k] = gi < n_rows ? __ldcg(reinterpret_cast<const float4*>(partials + ((size_t)gi * 2 + 0) * C + cq * 4)) : f4s(0.f);
      v[k] = gi < n_rows ? __ldcg(reinterpret_cast<const float4*>(partials + ((size_t)gi * 2 + 1) * C + cq * 4)) : f4s(0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a[0] += u[k].x; a[1] += u[k].y; a[2] += u[k].z; a[3] += u[k].w;
      b[0] += v[k].x; b[1] += v[k].y; b[2] += v[k].z; b[3] += v[k].w;
    }
  }
  double* sd = reinterpret_cast<double*>(sm);
  double* mine = sd + ((size_t)ty * C4 + cq) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) { mine[i] = a[i]; mine[4 + i] = b[i]; }
  __syncthreads();
  if (ty != 0) return true;
#pragma unroll
  for (int i = 0; i < 4; ++i) { t0[i] = a[i]; t1[i] = b[i]; }
  for (int j = 1; j < R; ++j) {
    const double* o = sd + ((size_t)j * C4 + cq) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) { t0[i] += o[i]; t1[i] += o[4 + i]; }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// BN statistics
// ------------------------------------------------------------------------------------------------
struct BnFin {      // what the finalizing CTA needs (all per-slot pointers)
  const float* gamma; const float* beta; float* mm; float* mv;
  float* mean_o; float* rstd_o; float* a_o; float* b_o;
  int M, ema, bessel;
};
// one channel of bn_finalize (double statistics, float outputs, EMA of the moving statistics)
__device__ __forceinline__ void bn_finalize_channel(const BnFin& f, int c, double s, double ss) {
  const double mean = s / f.M;
  double var = ss / f.M - mean * mean;   // biased (tf.nn.moments)
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)kBnEps));
  const float a = f.gamma[c] * rstd;
  f.mean_o[c] = (float)mean;
  f.rstd_o[c] = rstd;
  f.a_o[c] = a;
  f.b_o[c] = f.beta[c] - (float)mean * a;
  if (f.ema) {
    // moving -= (moving - batch) * (1 - momentum)   [TF-ext assign_moving_average, no zero-debias]
    const float bv = (float)(f.bessel ? var * ((double)f.M / (double)(f.M - 1)) : var);
    const float m0 = f.mm[c], v0 = f.mv[c];
    f.mm[c] = m0 - (m0 - (float)mean) * (1.f - kBnMomentum);
    f.mv[c] = v0 - (v0 - bv) * (1.f - kBnMomentum);
  }
}

// MODE 0: partial row per CTA, finalized by a second launch; 1: clusters + ticket (measured slower); 2: ticket
template <bool PRE_SWISH, int MODE>
__global__ void bn_stats_kernel(const float* __restrict__ x, int ld, int M, int C, int rows_per_chunk,
                                float* __restrict__ partials, unsigned* ticket, BnFin fin, long long zs) {
  extern __shared__ float4 sm[];
  {
    const size_t zo = (size_t)blockIdx.z * zs;
    x += zo; partials += zo; ticket += zo;
    fin.gamma += zo; fin.beta += zo; fin.mm += zo; fin.mv += zo; fin.mean_o += zo; fin.rstd_o += zo; fin.a_o += zo; fin.b_o += zo;
  }
  const int cq = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_chunk;
  const int r1 = min(M, r0 + rows_per_chunk);
  float4 s = f4s(0.f), ss = f4s(0.f);
  constexpr int U = 8;   // rows in flight per thread (loads issued before any is consumed; same summation order)
  for (int rb = r0 + threadIdx.y; rb < r1; rb += U * blockDim.y) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * blockDim.y;
      if (r < r1) v[u] = ld4(x + (size_t)r * ld + cq * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (rb + u * blockDim.y < r1) {
        const float4 w = PRE_SWISH ? swish4(v[u]) : v[u];
        s = s + w;
        fma4(ss, w, w);
      }
    }
  }
  block_reduce2(s, ss, sm);
  if (MODE == 0) {      // one partial row per CTA, finalized by a second launch (round-1 structure)
    if (threadIdx.y == 0) {
      st4(partials + ((size_t)blockIdx.x * 2 + 0) * C + cq * 4, s);
      st4(partials + ((size_t)blockIdx.x * 2 + 1) * C + cq * 4, ss);
    }
    return;
  }
  double t0[4], t1[4];
  const bool last = MODE == 1 ? cluster_total2(s, ss, sm, partials, C, ticket, t0, t1)
                              : ticket_total2(s, ss, sm, partials, C, ticket, t0, t1);
  if (last && threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) bn_finalize_channel(fin, cq * 4 + i, t0[i], t1[i]);
  }
}

// blockDim = (32 channels, 16 partial lanes): each thread sums every 16th partial in double, then the 16
// lanes are combined in a fixed order through shared memory (deterministic).
__global__ void __launch_bounds__(512) bn_finalize_kernel(const float* __restrict__ partials, int G, int C, int M,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ mm, float* __restrict__ mv, int ema, int bessel,
                                   float* __restrict__ mean_o, float* __restrict__ rstd_o, float* __restrict__ a_o,
                                   float* __restrict__ b_o, long long zs) {
  __shared__ double red[2][16][33];
  {
    const size_t zo = (size_t)blockIdx.z * zs;
    partials += zo; gamma += zo; beta += zo; mm += zo; mv += zo; mean_o += zo; rstd_o += zo; a_o += zo; b_o += zo;
  }
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0, ss = 0.0;
  if (c < C) {
    s = strided_sum_d(partials + c, G, 2 * (size_t)C, threadIdx.y, 16);
    ss = strided_sum_d(partials + C + c, G, 2 * (size_t)C, threadIdx.y, 16);
  }
  red[0][threadIdx.y][threadIdx.x] = s;
  red[1][threadIdx.y][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.y != 0 || c >= C) return;
  for (int j = 1; j < 16; ++j) { s += red[0][j][threadIdx.x]; ss += red[1][j][threadIdx.x]; }
  double mean = s / M;
  double var = ss / M - mean * mean;   // biased (tf.nn.moments)
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + (double)kBnEps));
  float a = gamma[c] * rstd;
  mean_o[c] = (float)mean;
  rstd_o[c] = rstd;
  a_o[c] = a;
  b_o[c] = beta[c] - (float)mean * a;
  if (ema) {
    // moving -= (moving - batch) * (1 - momentum)   [TF-ext assign_moving_average, no zero-debias]
    float bv = (float)(bessel ? var * ((double)M / (double)(M - 1)) : var);
    float m0 = mm[c], v0 = mv[c];
    mm[c] = m0 - (m0 - (float)mean) * (1.f - kBnMomentum);
    mv[c] = v0 - (v0 - bv) * (1.f - kBnMomentum);
  }
}


// MEASURED (round 2, B200): the single-launch clustered reduction is SLOWER than reduce + finalize as two launches - whole
// job 105.4 vs 110.4 tasks/s, bn_stats + finalize at 14x14x672 30.2 vs 23.4 us - although it saves 390 launches per task
// (2502 -> 2112): the 8-CTA clusters constrain scheduling next to the other slots' kernels, three cluster barriers sit
// in every CTA's path and the finalize runs on one CTA while its grid drains.  It stays available (MLIIS_BN_CLUSTER=1) as
// a measured negative result; the default is the two-launch structure.
// Finalize inside the reduction through a per-slot ticket (no clusters), for task-batched launches where a slot has few
// row chunks (37-148).  MEASURED (round 2, B200, 48 slots in groups of 24): correct (kernel, group, meta-step parity tests)
// and 16 % fewer launches (42240 vs 50040 per 480 tasks) - and SLOWER, like the clustered variant: 144.5 vs 145.3
// tasks/s, FOMAML 20.4 vs 21.5, Reptile 3.48 vs 3.55 meta-steps/s.  The last CTA's serial sum holds back every kernel
// that depends on the statistics, while the separate finalize launch spreads the same sum over 100-500 CTAs.  Opt-in
// (MLIIS_BN_TICKET=1) as a measured negative result; the default stays reduce + finalize as two launches.
static bool bn_ticket() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MLIIS_BN_TICKET"); on = e ? atoi(e) : 0; }
  return on != 0 && partition_nz() >= 2;
}
static bool bn_clustered() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MLIIS_BN_CLUSTER"); on = e ? atoi(e) : 0; }
  return on != 0;
}

// train-mode batch statistics + finalize (mean, rstd, the affine coefficients a/b, EMA of the moving statistics) in ONE
// launch.  partials: rc_num_chunks(M, C) / 8 rows of [2][C]; ticket: one zero-initialised word per slot (self-resetting).
void bn_stats_finalize(const float* x, int ld, int M, int C, bool pre_swish, float* partials, unsigned* ticket,
                       const float* gamma, const float* beta, float* mm, float* mv, int ema, int bessel, float* mean,
                       float* rstd, float* a, float* b, cudaStream_t s) {
  const int G = rc_num_chunks(M, C);
  const int rpc = cdiv(M, G);
  const dim3 blk = rc_block(C);
  const size_t smem = (size_t)blk.x * blk.y * 64;
  BnFin fin{gamma, beta, mm, mv, mean, rstd, a, b, M, ema, bessel};
  const long long zs = MLIIS_ZS;
  if (bn_ticket()) {      // task-batched launch: finalize in the CTA that finishes the slot's reduction (no second launch)
    if (pre_swish) MLIIS_COUNT(), bn_stats_kernel<true, 2><<<dim3(G, 1, MLIIS_NZ), blk, smem, s>>>(x, ld, M, C, rpc, partials, ticket, fin, zs);
    else MLIIS_COUNT(), bn_stats_kernel<false, 2><<<dim3(G, 1, MLIIS_NZ), blk, smem, s>>>(x, ld, M, C, rpc, partials, ticket, fin, zs);
    return;
  }
  if (!bn_clustered()) {
    if (pre_swish) MLIIS_COUNT(), bn_stats_kernel<true, 0><<<dim3(G, 1, MLIIS_NZ), blk, smem, s>>>(x, ld, M, C, rpc, partials, ticket, fin, zs);
    else MLIIS_COUNT(), bn_stats_kernel<false, 0><<<dim3(G, 1, MLIIS_NZ), blk, smem, s>>>(x, ld, M, C, rpc, partials, ticket, fin, zs);
    MLIIS_COUNT(), bn_finalize_kernel<<<dim3(cdiv(C, 32), 1, MLIIS_NZ), dim3(32, 16), 0, s>>>(partials, G, C, M, gamma, beta, mm, mv,
                                                                                            ema, bessel, mean, rstd, a, b, zs);
    return;
  }
  if (pre_swish)
    MLIIS_COUNT(), launch_clustered(bn_stats_kernel<true, 1>, G, MLIIS_NZ, blk, smem, s, x, ld, M, C, rpc, partials, ticket, fin, zs);
  else
    MLIIS_COUNT(), launch_clustered(bn_stats_kernel<false, 1>, G, MLIIS_NZ, blk, smem, s, x, ld, M, C, rpc, partials, ticket, fin, zs);
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ theta, const int32_t* __restrict__ gi,
                                      const int32_t* __restrict__ bi, const float* __restrict__ mm,
                                      const float* __restrict__ mv, int n, float* __restrict__ a,
                                      float* __restrict__ b, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; theta += zo; mm += zo; mv += zo; a += zo; b += zo; }
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float inv = rsqrtf(mv[c] + kBnEps) * theta[gi[c]];
  a[c] = inv;
  b[c] = theta[bi[c]] - mm[c] * inv;
}
void bn_eval_coeffs(const float* theta, const int32_t* gi, const int32_t* bi, const float* mm, const float* mv,
                    int n, float* a, float* b, cudaStream_t s) {
  MLIIS_COUNT(), bn_eval_coeffs_kernel<<<dim3(cdiv(n, 256), 1, MLIIS_NZ), 256, 0, s>>>(theta, gi, bi, mm, mv, n, a, b, MLIIS_ZS);
}

// ------------------------------------------------------------------------------------------------
// elementwise appliers
// ------------------------------------------------------------------------------------------------
__global__ void dec_bn_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ a,
                                    const float* __restrict__ b, const float* __restrict__ res, int ldres,
                                    float* __restrict__ y, int ldy, int M, int rows_per_block, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; x += zo; a += zo; b += zo; res = zp(res, zo); y += zo; }
  const int cq = threadIdx.x;
  const float4 av = ld4(a + cq * 4), bv = ld4(b + cq * 4);
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    float4 v = affine4(swish4(ld4(x + (size_t)r * ldx + cq * 4)), av, bv);
    if (res) v = v + ld4(res + (size_t)r * ldres + cq * 4);
    st4(y + (size_t)r * ldy + cq * 4, v);
  }
}
void dec_bn_apply(const float* x, int ldx, const float* a, const float* b, const float* res, int ldres, float* y,
                  int ldy, int M, int C, cudaStream_t s) {
  dim3 blk = rc_block(C);
  int rpb = blk.y * 4;
  MLIIS_COUNT(), dec_bn_apply_kernel<<<dim3(cdiv(M, rpb), 1, MLIIS_NZ), blk, 0, s>>>(x, ldx, a, b, res, ldres, y, ldy, M, rpb, MLIIS_ZS);
}

__global__ void block_out_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ a,
                                 const float* __restrict__ b, const float* __restrict__ dcs,
                                 const float* __restrict__ res, int ldres, float* __restrict__ y, int ldy, int M,
                                 int HW, int rows_per_block, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; x += zo; a += zo; b += zo; dcs = zp(dcs, zo); res = zp(res, zo); y += zo; }
  const int cq = threadIdx.x;
  const float4 av = ld4(a + cq * 4), bv = ld4(b + cq * 4);
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    float4 v = affine4(ld4(x + (size_t)r * ldx + cq * 4), av, bv);
    if (dcs) v = v * dcs[r / HW];
    if (res) v = v + ld4(res + (size_t)r * ldres + cq * 4);
    st4(y + (size_t)r * ldy + cq * 4, v);
  }
}
void block_out(const float* x, int ldx, const float* a, const float* b, const float* dcs, const float* res,
               int ldres, float* y, int ldy, int M, int C, int HW, cudaStream_t s) {
  dim3 blk = rc_block(C);
  int rpb = blk.y * 4;
  MLIIS_COUNT(), block_out_kernel<<<dim3(cdiv(M, rpb), 1, MLIIS_NZ), blk, 0, s>>>(x, ldx, a, b, dcs, res, ldres, y, ldy, M, HW, rpb,
                                                                                 MLIIS_ZS);
}

// ------------------------------------------------------------------------------------------------
// squeeze-excite
// ------------------------------------------------------------------------------------------------
// MODE 0: sum swish(a*x+b)        (SE squeeze)
// MODE 1: sum g * swish(a*x+b)    (dgate)
// MODE 2: sum x                   (plain per-image column sum)
template <int MODE>
__global__ void img_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ g, int ldg,
                                  const float* __restrict__ a, const float* __restrict__ b, int HW, int C,
                                  int rows_per_chunk, float* __restrict__ partial, long long zs) {
  extern __shared__ float4 sm[];
  { const size_t zo = (size_t)blockIdx.z * zs; x += zo; g = zp(g, zo); a = zp(a, zo); b = zp(b, zo); partial += zo; }
  const int cq = threadIdx.x, img = blockIdx.y, G = gridDim.x;
  float4 av = f4s(1.f), bv = f4s(0.f);
  if (MODE != 2) { av = ld4(a + cq * 4); bv = ld4(b + cq * 4); }
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  float4 s = f4s(0.f);
  constexpr int U = MODE == 1 ? 4 : 8;
  for (int rb = r0 + threadIdx.y; rb < r1; rb += U * blockDim.y) {
    float4 v[U], gv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * blockDim.y;
      if (r < r1) {
        const size_t row = (size_t)img * HW + r;
        v[u] = ld4(x + row * ldx + cq * 4);
        if (MODE == 1) gv[u] = ld4(g + row * ldg + cq * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (rb + u * blockDim.y < r1) {
        float4 w = v[u];
        if (MODE != 2) w = swish4(affine4(w, av, bv));
        if (MODE == 1) w = w * gv[u];
        s = s + w;
      }
    }
  }
  block_reduce1(s, sm);
  if (threadIdx.y == 0) st4(partial + ((size_t)img * G + blockIdx.x) * C + cq * 4, s);
}

void se_pool(const float* x, int ldx, const float* a, const float* b, int B, int HW, int C, float* partial,
             cudaStream_t s) {
  int G = rc_num_img_chunks(HW, C);
  dim3 blk = rc_block(C);
  MLIIS_COUNT(), img_reduce_kernel<0><<<dim3(G, B, MLIIS_NZ), blk, blk.x * blk.y * sizeof(float4), s>>>(x, ldx, nullptr, 0, a, b, HW, C,
                                                                               cdiv(HW, G), partial, MLIIS_ZS);
}
void se_bwd_reduce(const float* x, int ldx, const float* gup, int ldg, const float* a, const float* b, int B, int HW,
                   int C, float* partial, cudaStream_t s) {
  int G = rc_num_img_chunks(HW, C);
  dim3 blk = rc_block(C);
  MLIIS_COUNT(), img_reduce_kernel<1><<<dim3(G, B, MLIIS_NZ), blk, blk.x * blk.y * sizeof(float4), s>>>(x, ldx, gup, ldg, a, b, HW, C,
                                                                               cdiv(HW, G), partial, MLIIS_ZS);
}

__global__ void img_colsum_finalize_kernel(const float* __restrict__ partial, int G, int C, float scale,
                                           float* __restrict__ out, int ldo, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; partial += zo; out += zo; }
  int c = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
  if (c >= C) return;
  const double s = strided_sum_d(partial + (size_t)img * G * C + c, G, (size_t)C);
  out[(size_t)img * ldo + c] = (float)(s * (double)scale);
}
void img_colsum(const float* x, int ldx, int B, int HW, int C, float scale, float* partial, float* out, int ldo,
                cudaStream_t s) {
  int G = rc_num_img_chunks(HW, C);
  dim3 blk = rc_block(C);
  MLIIS_COUNT(), img_reduce_kernel<2><<<dim3(G, B, MLIIS_NZ), blk, blk.x * blk.y * sizeof(float4), s>>>(x, ldx, nullptr, 0, nullptr, nullptr,
                                                                               HW, C, cdiv(HW, G), partial, MLIIS_ZS);
  MLIIS_COUNT(), img_colsum_finalize_kernel<<<dim3(cdiv(C, 128), B, MLIIS_NZ), 128, 0, s>>>(partial, G, C, scale, out, ldo, MLIIS_ZS);
}

// s += sum_i v[i0 + i*vstep] * w[(i0 + i*vstep) * wstride]  for i0 + i*vstep < n, in that order; the global loads of a batch of 8
// are issued before the first fma (these one-CTA-per-image kernels are pure latency: one round trip per 8 terms)
__device__ __forceinline__ float dot_strided(const float* v, const float* __restrict__ w, int i0, int vstep, int n,
                                             size_t wstride, float s) {
  for (int i = i0; i < n; i += 8 * vstep) {
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = i + u * vstep;
      wv[u] = k < n ? w[(size_t)k * wstride] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = i + u * vstep;
      if (k < n) s = fmaf(v[k], wv[u], s);
    }
  }
  return s;
}

// One block per image.  pool -> reduce FC (+bias, swish) -> expand FC (+bias) -> sigmoid.
__global__ void se_fc_fwd_kernel(const float* __restrict__ partial, int G, int HW, int C, int Cr,
                                 const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, const float* __restrict__ b2,
                                 float* __restrict__ pool_o, float* __restrict__ hidpre_o,
                                 float* __restrict__ gate_o, long long zs) {
  extern __shared__ float smf[];
  {
    const size_t zo = (size_t)blockIdx.z * zs;
    partial += zo; w1 += zo; b1 += zo; w2 += zo; b2 += zo; pool_o += zo; hidpre_o += zo; gate_o += zo;
  }
  float* pool = smf;          // [C]
  float* hid = smf + C;       // [Cr]
  const int img = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float inv = 1.f / (float)HW;
  for (int c = tid; c < C; c += nt) {
    float s = strided_sum_f(partial + (size_t)img * G * C + c, G, (size_t)C);
    s *= inv;
    pool[c] = s;
    pool_o[(size_t)img * C + c] = s;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int r = warp; r < Cr; r += nw) {
    float s = dot_strided(pool, w1 + r, lane, 32, C, (size_t)Cr, 0.f);
    s = warp_sum(s);
    if (lane == 0) {
      s += b1[r];
      hidpre_o[(size_t)img * Cr + r] = s;
      hid[r] = swish_f(s);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += nt) {
    const float s = dot_strided(hid, w2 + c, 0, 1, Cr, (size_t)C, b2[c]);
    gate_o[(size_t)img * C + c] = sigmoid_f(s);
  }
}
void se_fc_fwd(const float* partial, int G, int B, int HW, int C, int Cr, const float* w1, const float* b1,
               const float* w2, const float* b2, float* pool, float* hidpre, float* gate, cudaStream_t s) {
  MLIIS_COUNT(), se_fc_fwd_kernel<<<dim3(B, 1, MLIIS_NZ), 256, (C + Cr) * sizeof(float), s>>>(partial, G, HW, C, Cr, w1, b1, w2, b2, pool,
                                                                                hidpre, gate, MLIIS_ZS);
}

// SE FC backward, phase A (one block per image): d(gate pre-act), d(hidden pre-act), d(pool).
__global__ void __launch_bounds__(256) se_fc_bwd_img_kernel(const float* __restrict__ partial, int G, int HW, int C,
                                                             int Cr, const float* __restrict__ w1,
                                                             const float* __restrict__ w2,
                                                             const float* __restrict__ hidpre,
                                                             const float* __restrict__ gate, float* __restrict__ dgp_o,
                                                             float* __restrict__ dhp_o, float* __restrict__ dpool,
                                                             long long zs) {
  extern __shared__ float smf[];
  {
    const size_t zo = (size_t)blockIdx.z * zs;
    partial += zo; w1 += zo; w2 += zo; hidpre += zo; gate += zo; dgp_o += zo; dhp_o += zo; dpool += zo;
  }
  float* dgp = smf;        // [C]
  float* dhp = smf + C;    // [Cr]
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  for (int c = tid; c < C; c += nt) {
    const float s = strided_sum_f(partial + (size_t)b * G * C + c, G, (size_t)C);
    const float gt = gate[(size_t)b * C + c];
    const float v = s * gt * (1.f - gt);
    dgp[c] = v;
    dgp_o[(size_t)b * C + c] = v;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int r = warp; r < Cr; r += nw) {
    float s = dot_strided(dgp, w2 + (size_t)r * C, lane, 32, C, 1, 0.f);
    s = warp_sum(s);
    if (lane == 0) {
      const float v = s * swish_grad_f(hidpre[(size_t)b * Cr + r]);
      dhp[r] = v;
      dhp_o[(size_t)b * Cr + r] = v;
    }
  }
  __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int c = tid; c < C; c += nt) {
    const float s = dot_strided(dhp, w1 + (size_t)c * Cr, 0, 1, Cr, 1, 0.f);
    dpool[(size_t)b * C + c] = s * inv;
  }
}
// phase B (grid over channel chunks): weight / bias gradients, summed over the batch in a fixed order.
__global__ void __launch_bounds__(128) se_fc_bwd_w_kernel(int B, int C, int Cr, const float* __restrict__ pool,
                                                           const float* __restrict__ hidpre,
                                                           const float* __restrict__ dgp, const float* __restrict__ dhp,
                                                           float* __restrict__ dw1, float* __restrict__ db1,
                                                           float* __restrict__ dw2, float* __restrict__ db2,
                                                           long long zs) {
  {
    const size_t zo = (size_t)blockIdx.z * zs;
    pool += zo; hidpre += zo; dgp += zo; dhp += zo; dw1 += zo; db1 += zo; dw2 += zo; db2 += zo;
  }
  // the per-image hidden activations / gradients are shared by every channel thread: staged once in shared memory
  // (the first version re-evaluated swish(hidpre) and re-read both vectors from global memory for every (c, r, b))
  extern __shared__ float smw[];
  float* sh = smw;               // [B][Cr] swish(hidden pre-activation)
  float* sd = smw + B * Cr;      // [B][Cr] d(hidden pre-activation)
  for (int i = threadIdx.x; i < B * Cr; i += blockDim.x) { sh[i] = swish_f(hidpre[i]); sd[i] = dhp[i]; }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float sb = 0.f;
    for (int b = 0; b < B; ++b) sb += dgp[(size_t)b * C + c];
    db2[c] = sb;
    for (int r = 0; r < Cr; ++r) {
      float s2 = 0.f, s1 = 0.f;
      for (int b = 0; b < B; ++b) {
        s2 = fmaf(sh[b * Cr + r], dgp[(size_t)b * C + c], s2);
        s1 = fmaf(pool[(size_t)b * C + c], sd[b * Cr + r], s1);
      }
      dw2[(size_t)r * C + c] = s2;
      dw1[(size_t)c * Cr + r] = s1;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < Cr) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dhp[(size_t)b * Cr + threadIdx.x];
    db1[threadIdx.x] = s;
  }
}
void se_fc_bwd(const float* partial, int G, int B, int HW, int C, int Cr, const float* w1, const float* w2,
               const float* pool, const float* hidpre, const float* gate, float* dw1, float* db1, float* dw2,
               float* db2, float* dpool, cudaStream_t s) {
  // scratch for dgp [B][C] and dhp [B][Cr] lives right after the dgate partials
  float* dgp = const_cast<float*>(partial) + (size_t)B * G * C;
  float* dhp = dgp + (size_t)B * C;
  MLIIS_COUNT(), se_fc_bwd_img_kernel<<<dim3(B, 1, MLIIS_NZ), 256, (C + Cr) * sizeof(float), s>>>(partial, G, HW, C, Cr, w1, w2, hidpre,
                                                                                    gate, dgp, dhp, dpool, MLIIS_ZS);
  MLIIS_COUNT(), se_fc_bwd_w_kernel<<<dim3(cdiv(C, 128), 1, MLIIS_NZ), 128, 2 * B * Cr * sizeof(float), s>>>(B, C, Cr, pool, hidpre, dgp, dhp, dw1, db1, dw2, db2,
                                                                                   MLIIS_ZS);
}

// ------------------------------------------------------------------------------------------------
// BN backward (train mode): dgamma = sum g*xhat, dbeta = sum g, dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat))
// ------------------------------------------------------------------------------------------------
template <int VAR>
__device__ __forceinline__ void bn_bwd_elem(const BnBwdArgs& p, size_t r, int cq, float4 x, float4 g, float4 mean,
                                            float4 rstd, float4 av, float4 bv, float4& geff, float4& xhat,
                                            float4& post) {
  post = f4s(1.f);
  if (VAR == BN_PLAIN) {
    if (p.dcs) g = g * p.dcs[r / p.HW];
    geff = g;
    xhat = (x - mean) * rstd;
  } else if (VAR == BN_SWISH) {
    geff = g * swish_grad4(affine4(x, av, bv));
    xhat = (x - mean) * rstd;
  } else if (VAR == BN_SWISH_SE) {
    size_t ic = (r / p.HW) * (size_t)p.C + cq * 4;
    float4 dz = g * ld4(p.gate + ic) + ld4(p.dpool + ic);
    geff = dz * swish_grad4(affine4(x, av, bv));
    xhat = (x - mean) * rstd;
  } else {  // BN_DEC: s = swish(x); y = BN(s)
    xhat = (swish4(x) - mean) * rstd;
    geff = g;
    post = swish_grad4(x);
  }
}

__device__ __forceinline__ void bn_bwd_shift(BnBwdArgs& p, long long zs) {
  const size_t zo = (size_t)blockIdx.z * zs;
  p.x += zo; p.g += zo; p.dx += zo; p.mean += zo; p.rstd += zo; p.a += zo; p.b += zo; p.gamma += zo;
  p.dcs = zp(p.dcs, zo); p.gate = zp(p.gate, zo); p.dpool = zp(p.dpool, zo);
  p.partials += zo; p.k += zo; p.dgamma += zo; p.dbeta += zo; p.ticket += zo;
}

template <int VAR, int MODE>
__global__ void bn_bwd_reduce_kernel(BnBwdArgs p, int rows_per_chunk, long long zs) {
  extern __shared__ float4 sm[];
  bn_bwd_shift(p, zs);
  const int cq = threadIdx.x;
  const float4 mean = ld4(p.mean + cq * 4), rstd = ld4(p.rstd + cq * 4);
  const float4 av = ld4(p.a + cq * 4), bv = ld4(p.b + cq * 4);
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(p.M, r0 + rows_per_chunk);
  float4 s0 = f4s(0.f), s1 = f4s(0.f);
  constexpr int U = 4;     // rows in flight per thread: the 2*U loads are issued before any is consumed
  for (int rb = r0 + threadIdx.y; rb < r1; rb += U * blockDim.y) {
    float4 x[U], g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * blockDim.y;
      if (r < r1) { x[u] = ld4(p.x + (size_t)r * p.ldx + cq * 4); g[u] = ld4(p.g + (size_t)r * p.ldg + cq * 4); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * blockDim.y;
      if (r < r1) {
        float4 ge, xh, po;
        bn_bwd_elem<VAR>(p, (size_t)r, cq, x[u], g[u], mean, rstd, av, bv, ge, xh, po);
        s0 = s0 + ge;
        fma4(s1, ge, xh);
      }
    }
  }
  block_reduce2(s0, s1, sm);
  if (MODE == 0) {
    if (threadIdx.y == 0) {
      st4(p.partials + ((size_t)blockIdx.x * 2 + 0) * p.C + cq * 4, s0);
      st4(p.partials + ((size_t)blockIdx.x * 2 + 1) * p.C + cq * 4, s1);
    }
    return;
  }
  double t0[4], t1[4];
  const bool last = MODE == 1 ? cluster_total2(s0, s1, sm, p.partials, p.C, p.ticket, t0, t1)
                              : ticket_total2(s0, s1, sm, p.partials, p.C, p.ticket, t0, t1);
  if (last && threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {       // dbeta = sum g, dgamma = sum g*xhat, and their means for the apply pass
      const int c = cq * 4 + i;
      p.dbeta[c] = (float)t0[i];
      p.dgamma[c] = (float)t1[i];
      p.k[c] = (float)(t0[i] / p.M);
      p.k[p.C + c] = (float)(t1[i] / p.M);
    }
  }
}

__global__ void __launch_bounds__(512) bn_bwd_finalize_kernel(const float* __restrict__ partials, int G, int C, int M,
                                       float* __restrict__ k, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, long long zs) {
  __shared__ double red[2][16][33];
  { const size_t zo = (size_t)blockIdx.z * zs; partials += zo; k += zo; dgamma += zo; dbeta += zo; }
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  if (c < C) {
    s0 = strided_sum_d(partials + c, G, 2 * (size_t)C, threadIdx.y, 16);
    s1 = strided_sum_d(partials + C + c, G, 2 * (size_t)C, threadIdx.y, 16);
  }
  red[0][threadIdx.y][threadIdx.x] = s0;
  red[1][threadIdx.y][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.y != 0 || c >= C) return;
  for (int j = 1; j < 16; ++j) { s0 += red[0][j][threadIdx.x]; s1 += red[1][j][threadIdx.x]; }
  dbeta[c] = (float)s0;
  dgamma[c] = (float)s1;
  k[c] = (float)(s0 / M);
  k[C + c] = (float)(s1 / M);
}

template <int VAR>
__global__ void bn_bwd_apply_kernel(BnBwdArgs p, int rows_per_block, long long zs) {
  bn_bwd_shift(p, zs);
  const int cq = threadIdx.x;
  const float4 mean = ld4(p.mean + cq * 4), rstd = ld4(p.rstd + cq * 4);
  const float4 av = ld4(p.a + cq * 4), bv = ld4(p.b + cq * 4);
  const float4 ga = ld4(p.gamma + cq * 4) * rstd;
  const float4 k1 = ld4(p.k + cq * 4), k2 = ld4(p.k + p.C + cq * 4);
  const int r0 = blockIdx.x * rows_per_block, r1 = min(p.M, r0 + rows_per_block);
  constexpr int U = 4;     // == rows_per_block / blockDim.y: all loads of the block are in flight before the stores
  for (int rb = r0 + threadIdx.y; rb < r1; rb += U * blockDim.y) {
    float4 x[U], g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * blockDim.y;
      if (r < r1) { x[u] = ld4(p.x + (size_t)r * p.ldx + cq * 4); g[u] = ld4(p.g + (size_t)r * p.ldg + cq * 4); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u * blockDim.y;
      if (r < r1) {
        float4 ge, xh, po;
        bn_bwd_elem<VAR>(p, (size_t)r, cq, x[u], g[u], mean, rstd, av, bv, ge, xh, po);
        float4 d = ga * (ge - k1 - xh * k2);
        if (VAR == BN_DEC) d = d * po;
        st4(p.dx + (size_t)r * p.lddx + cq * 4, d);
      }
    }
  }
}

template <int VAR>
static void bn_bwd_t(const BnBwdArgs& p, cudaStream_t s) {
  int G = rc_num_chunks(p.M, p.C);
  dim3 blk = rc_block(p.C);
  const int nz = MLIIS_NZ;
  const long long zs = MLIIS_ZS;
  if (bn_ticket()) {
    MLIIS_COUNT(), bn_bwd_reduce_kernel<VAR, 2><<<dim3(G, 1, nz), blk, (size_t)blk.x * blk.y * 64, s>>>(p, cdiv(p.M, G), zs);
  } else if (bn_clustered()) {
    MLIIS_COUNT(), launch_clustered(bn_bwd_reduce_kernel<VAR, 1>, G, nz, blk, (size_t)blk.x * blk.y * 64, s, p, cdiv(p.M, G), zs);
  } else {
    MLIIS_COUNT(), bn_bwd_reduce_kernel<VAR, 0><<<dim3(G, 1, nz), blk, (size_t)blk.x * blk.y * 64, s>>>(p, cdiv(p.M, G), zs);
    MLIIS_COUNT(), bn_bwd_finalize_kernel<<<dim3(cdiv(p.C, 32), 1, nz), dim3(32, 16), 0, s>>>(p.partials, G, p.C, p.M, p.k, p.dgamma,
                                                                                             p.dbeta, zs);
  }
  int rpb = blk.y * 4;
  MLIIS_COUNT(), bn_bwd_apply_kernel<VAR><<<dim3(cdiv(p.M, rpb), 1, nz), blk, 0, s>>>(p, rpb, zs);
}
void bn_bwd(int var, const BnBwdArgs& p, cudaStream_t s) {
  switch (var) {
    case BN_PLAIN: bn_bwd_t<BN_PLAIN>(p, s); break;
    case BN_SWISH: bn_bwd_t<BN_SWISH>(p, s); break;
    case BN_SWISH_SE: bn_bwd_t<BN_SWISH_SE>(p, s); break;
    default: bn_bwd_t<BN_DEC>(p, s); break;
  }
}

// ------------------------------------------------------------------------------------------------
// concat / broadcast plumbing
// ------------------------------------------------------------------------------------------------
__global__ void bcast_rows_kernel(const float* __restrict__ pimg, int ldp, float* __restrict__ y, int ldy, int M,
                                  int HW, int rows_per_block, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; pimg += zo; y += zo; }
  const int cq = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y)
    st4(y + (size_t)r * ldy + cq * 4, ld4(pimg + (size_t)(r / HW) * ldp + cq * 4));
}
void bcast_rows(const float* pimg, int ldp, float* y, int ldy, int B, int HW, int C, cudaStream_t s) {
  dim3 blk = rc_block(C);
  int rpb = blk.y * 4, M = B * HW;
  MLIIS_COUNT(), bcast_rows_kernel<<<dim3(cdiv(M, rpb), 1, MLIIS_NZ), blk, 0, s>>>(pimg, ldp, y, ldy, M, HW, rpb, MLIIS_ZS);
}

__global__ void add3_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ a, int lda,
                            const float* __restrict__ b, int ldb, const float* __restrict__ pimg, int ldp, int M,
                            int HW, int rows_per_block, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; dst += zo; a += zo; b = zp(b, zo); pimg = zp(pimg, zo); }
  const int cq = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    float4 v = ld4(a + (size_t)r * lda + cq * 4);
    if (b) v = v + ld4(b + (size_t)r * ldb + cq * 4);
    if (pimg) v = v + ld4(pimg + (size_t)(r / HW) * ldp + cq * 4);
    st4(dst + (size_t)r * ldd + cq * 4, v);
  }
}
void add3(float* dst, int ldd, const float* a, int lda, const float* b, int ldb, const float* pimg, int ldp, int M,
          int C, int HW, cudaStream_t s) {
  dim3 blk = rc_block(C);
  int rpb = blk.y * 4;
  MLIIS_COUNT(), add3_kernel<<<dim3(cdiv(M, rpb), 1, MLIIS_NZ), blk, 0, s>>>(dst, ldd, a, lda, b, ldb, pimg, ldp, M, HW, rpb, MLIIS_ZS);
}

}  // namespace mliis
