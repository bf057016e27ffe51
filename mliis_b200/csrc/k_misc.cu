// Bilinear resize (align_corners=True) fwd/bwd, logits head, fused softmax-CE / soft-dice loss and its
// gradient, prediction threshold + integer IoU counts, multi-tensor Adam/SGD, meta-update kernels.
//
// Reference: models/efficientlab.py:161-177 (dropout, head, resize, softmax), :291-327 (threshold, loss),
// :329-396 (soft IoU); reptile.py:526-549 (_iou); meta_learners/variables.py:9-45; args.py:151-154.
#include "common.cuh"
#include "kernels.h"

namespace mliis {

// =============================================================================================
// generic partial reducer (fixed order, double accumulation)
// =============================================================================================
// blockDim = (32 outputs, 8 partial lanes); coalesced over outputs, fixed-order combine over lanes.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int G, int n,
                                                              float* __restrict__ out, long long zs) {
  __shared__ double red[8][33];
  { const size_t zo = (size_t)blockIdx.z * zs; partials += zo; out += zo; }
  const int i = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (i < n) s = strided_sum_d(partials + i, G, (size_t)n, threadIdx.y, 8);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || i >= n) return;
  for (int j = 1; j < 8; ++j) s += red[j][threadIdx.x];
  out[i] = (float)s;
}
__global__ void __launch_bounds__(256) reduce_partials_strided_kernel(const float* __restrict__ partials, int G, int n,
                                                                      int n_inner, float* __restrict__ out,
                                                                      int64_t out_stride, long long zs) {
  __shared__ double red[8][33];
  { const size_t zo = (size_t)blockIdx.z * zs; partials += zo; out += zo; }
  const int i = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (i < n) s = strided_sum_d(partials + i, G, (size_t)n, threadIdx.y, 8);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || i >= n) return;
  for (int j = 1; j < 8; ++j) s += red[j][threadIdx.x];
  const int o = i / n_inner;
  out[(size_t)o * out_stride + (i - o * n_inner)] = (float)s;
}
void reduce_partials_strided(const float* partials, int G, int n_inner, int n_outer, float* out, int64_t out_stride,
                             cudaStream_t s) {
  const int n = n_inner * n_outer;
  MLIIS_COUNT(), reduce_partials_strided_kernel<<<dim3(cdiv(n, 32), 1, MLIIS_NZ), dim3(32, 8), 0, s>>>(partials, G, n, n_inner, out,
                                                                                                      out_stride, MLIIS_ZS);
}
// G <= 8 partials (task-batched launches split every slot's reduction a few times only): one float4 of outputs per
// thread, the partials added in index order in double - the same order as the lane version, where every lane then holds
// at most one partial.
__global__ void __launch_bounds__(256) reduce_partials_small_kernel(const float* __restrict__ partials, int G, int n4,
                                                                    float* __restrict__ out, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; partials += zo; out += zo; }
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 v[8];
#pragma unroll
  for (int g = 0; g < 8; ++g)
    if (g < G) v[g] = ld4(partials + ((size_t)g * n4 + i) * 4);
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
#pragma unroll
  for (int g = 0; g < 8; ++g)
    if (g < G) { a += (double)v[g].x; b += (double)v[g].y; c += (double)v[g].z; d += (double)v[g].w; }
  st4(out + (size_t)i * 4, f4((float)a, (float)b, (float)c, (float)d));
}
void reduce_partials(const float* partials, int G, int n, float* out, cudaStream_t s) {
  if (G <= 8 && n % 4 == 0 && (reinterpret_cast<uintptr_t>(partials) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      (MLIIS_ZS % 4) == 0) {
    MLIIS_COUNT(), reduce_partials_small_kernel<<<dim3(cdiv(n / 4, 256), 1, MLIIS_NZ), 256, 0, s>>>(partials, G, n / 4, out, MLIIS_ZS);
    return;
  }
  MLIIS_COUNT(), reduce_partials_kernel<<<dim3(cdiv(n, 32), 1, MLIIS_NZ), dim3(32, 8), 0, s>>>(partials, G, n, out, MLIIS_ZS);
}

// =============================================================================================
// bilinear, align_corners=True.  out = top + (bottom - top) * ylerp ; top = tl + (tr - tl) * xlerp [TF-ext]
// =============================================================================================
__global__ void bilinear_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int B,
                                    int Hi, int Wi, int Ho, int Wo, ResizeTab ty, ResizeTab tx, int rows_per_block,
                                    long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; x += zo; y += zo; }
  const int cq = threadIdx.x;
  const int M = B * Ho * Wo;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    const int b = r / (Ho * Wo), rem = r - b * (Ho * Wo), oy = rem / Wo, ox = rem - oy * Wo;
    const int y0 = ty.lo[oy], y1 = ty.hi[oy], x0 = tx.lo[ox], x1 = tx.hi[ox];
    const float yl = ty.lerp[oy], xl = tx.lerp[ox];
    const float* base = x + (size_t)b * Hi * Wi * ldx + cq * 4;
    const float4 tl = ld4(base + ((size_t)y0 * Wi + x0) * ldx), tr = ld4(base + ((size_t)y0 * Wi + x1) * ldx);
    const float4 bl = ld4(base + ((size_t)y1 * Wi + x0) * ldx), br = ld4(base + ((size_t)y1 * Wi + x1) * ldx);
    const float4 top = tl + (tr - tl) * xl, bot = bl + (br - bl) * xl;
    st4(y + (size_t)r * ldy + cq * 4, top + (bot - top) * yl);
  }
}
void bilinear_fwd(const float* x, int ldx, float* y, int ldy, int B, int Hi, int Wi, int Ho, int Wo, int C,
                  ResizeTab ty, ResizeTab tx, cudaStream_t s) {
  int c4 = C / 4, R = 256 / c4;
  if (R < 1) R = 1;
  if (R > 64) R = 64;
  dim3 blk(c4, R);
  int rpb = R * 4;
  MLIIS_COUNT(), bilinear_fwd_kernel<<<dim3(cdiv(B * Ho * Wo, rpb), 1, MLIIS_NZ), blk, 0, s>>>(x, ldx, y, ldy, B, Hi, Wi, Ho, Wo, ty, tx,
                                                                                              rpb, MLIIS_ZS);
}

__device__ __forceinline__ float gather_w(const ResizeTab& t, int o, int i) {
  const float l = t.lerp[o];
  return (t.lo[o] == i ? 1.f - l : 0.f) + (t.hi[o] == i ? l : 0.f);
}

// gather form of the adjoint: dx[iy,ix] = sum_{oy,ox} wy(oy,iy) * wx(ox,ix) * dy[oy,ox]
template <int VEC>
__global__ void bilinear_bwd_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx, int B,
                                    int Hi, int Wi, int Ho, int Wo, ResizeTab ty, ResizeTab tx, int rows_per_block,
                                    long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; dy += zo; dx += zo; }
  const int cq = threadIdx.x;
  const int M = B * Hi * Wi;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    const int b = r / (Hi * Wi), rem = r - b * (Hi * Wi), iy = rem / Wi, ix = rem - iy * Wi;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const int oy0 = ty.g_lo[iy], oy1 = ty.g_hi[iy], ox0 = tx.g_lo[ix], ox1 = tx.g_hi[ix];
    for (int oy = oy0; oy <= oy1; ++oy) {
      const float wy = gather_w(ty, oy, iy);
      if (wy == 0.f) continue;
      const float* rowp = dy + ((size_t)(b * Ho + oy) * Wo) * lddy + cq * VEC;
      for (int ox = ox0; ox <= ox1; ++ox) {
        const float w = wy * gather_w(tx, ox, ix);
        const float* p = rowp + (size_t)ox * lddy;
        if (VEC == 4) {
          const float4 g = ld4(p);
          acc[0] = fmaf(w, g.x, acc[0]); acc[1] = fmaf(w, g.y, acc[1]);
          acc[2 % VEC] = fmaf(w, g.z, acc[2 % VEC]); acc[3 % VEC] = fmaf(w, g.w, acc[3 % VEC]);
        } else {
          const float2 g = *reinterpret_cast<const float2*>(p);
          acc[0] = fmaf(w, g.x, acc[0]); acc[1] = fmaf(w, g.y, acc[1]);
        }
      }
    }
    float* o = dx + (size_t)r * lddx + cq * VEC;
    if (VEC == 4) st4(o, f4(acc[0], acc[1], acc[2 % VEC], acc[3 % VEC]));
    else *reinterpret_cast<float2*>(o) = make_float2(acc[0], acc[1]);
  }
}
void bilinear_bwd(const float* dy, int lddy, float* dx, int lddx, int B, int Hi, int Wi, int Ho, int Wo, int C,
                  ResizeTab ty, ResizeTab tx, cudaStream_t s) {
  if (C == 2) {
    dim3 blk(1, 128);
    int rpb = 128;
    MLIIS_COUNT(), bilinear_bwd_kernel<2><<<dim3(cdiv(B * Hi * Wi, rpb), 1, MLIIS_NZ), blk, 0, s>>>(dy, lddy, dx, lddx, B, Hi, Wi, Ho, Wo, ty,
                                                                                                   tx, rpb, MLIIS_ZS);
  } else {
    int c4 = C / 4, R = 256 / c4;
    if (R < 1) R = 1;
    if (R > 64) R = 64;
    dim3 blk(c4, R);
    int rpb = R;
    MLIIS_COUNT(), bilinear_bwd_kernel<4><<<dim3(cdiv(B * Hi * Wi, rpb), 1, MLIIS_NZ), blk, 0, s>>>(dy, lddy, dx, lddx, B, Hi, Wi, Ho, Wo, ty,
                                                                                                   tx, rpb, MLIIS_ZS);
  }
}

// =============================================================================================
// logits head: dropout -> 1x1 conv C->2 + bias       (efficientlab.py:161-167)
// =============================================================================================
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ x, int ldx,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ mask, float keep_scale,
                                                        float* __restrict__ z, int M, int C, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; x += zo; w += zo; bias += zo; mask = zp(mask, zo); z += zo; }
  const int l8 = threadIdx.x & 7;
  const int p = blockIdx.x * 32 + (threadIdx.x >> 3);
  float a0 = 0.f, a1 = 0.f;
  if (p < M) {
    for (int c4 = l8; c4 < C / 4; c4 += 8) {
      float4 v = ld4(x + (size_t)p * ldx + c4 * 4);
      if (mask) v = v * ld4(mask + (size_t)p * C + c4 * 4) * keep_scale;
      const float4 w0 = ld4(w + c4 * 8), w1 = ld4(w + c4 * 8 + 4);   // w[c][j], j in {0,1}
      a0 += v.x * w0.x + v.y * w0.z + v.z * w1.x + v.w * w1.z;
      a1 += v.x * w0.y + v.y * w0.w + v.z * w1.y + v.w * w1.w;
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (p < M && l8 == 0) *reinterpret_cast<float2*>(z + (size_t)p * 2) = make_float2(a0 + bias[0], a1 + bias[1]);
}
void head_fwd(const float* x, int ldx, const float* w, const float* bias, const float* drop_mask, float keep_scale,
              float* z, int M, int C, cudaStream_t s) {
  MLIIS_COUNT(), head_fwd_kernel<<<dim3(cdiv(M, 32), 1, MLIIS_NZ), 256, 0, s>>>(x, ldx, w, bias, drop_mask, keep_scale, z, M, C, MLIIS_ZS);
}

// dx = (dz . w^T) * mask*scale ; dW[c][j] = sum_p xd[p,c] dz[p,j] ; db[j] = sum_p dz[p,j]
__global__ void head_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                const float* __restrict__ mask, float keep_scale, const float* __restrict__ dz,
                                float* __restrict__ dx, int lddx, float* __restrict__ partials, int M, int C,
                                int rows_per_chunk, long long zs) {
  extern __shared__ float4 sm[];
  { const size_t zo = (size_t)blockIdx.z * zs; x += zo; w += zo; mask = zp(mask, zo); dz += zo; dx += zo; partials += zo; }
  const int cq = threadIdx.x, C4 = blockDim.x, R = blockDim.y, ty = threadIdx.y;
  const float4 w0 = ld4(w + cq * 8), w1 = ld4(w + cq * 8 + 4);
  const float4 wj0 = f4(w0.x, w0.z, w1.x, w1.z), wj1 = f4(w0.y, w0.w, w1.y, w1.w);
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float4 g0 = f4s(0.f), g1 = f4s(0.f);
  float sb0 = 0.f, sb1 = 0.f;
  for (int r = r0 + ty; r < r1; r += R) {
    const float2 d = *reinterpret_cast<const float2*>(dz + (size_t)r * 2);
    float4 v = ld4(x + (size_t)r * ldx + cq * 4);
    float4 g = wj0 * d.x + wj1 * d.y;
    if (mask) {
      const float4 mk = ld4(mask + (size_t)r * C + cq * 4) * keep_scale;
      v = v * mk;
      g = g * mk;
    }
    st4(dx + (size_t)r * lddx + cq * 4, g);
    g0 = g0 + v * d.x;
    g1 = g1 + v * d.y;
    sb0 += d.x;
    sb1 += d.y;
  }
  // reduce over ty
  sm[ty * C4 + cq] = g0;
  sm[(R + ty) * C4 + cq] = g1;
  float* sbm = reinterpret_cast<float*>(sm + 2 * R * C4);
  if (cq == 0) { sbm[ty * 2] = sb0; sbm[ty * 2 + 1] = sb1; }
  __syncthreads();
  if (ty == 0) {
    for (int j = 1; j < R; ++j) { g0 = g0 + sm[j * C4 + cq]; g1 = g1 + sm[(R + j) * C4 + cq]; }
    float* o = partials + (size_t)blockIdx.x * (C * 2 + 2);
    o[(cq * 4 + 0) * 2] = g0.x; o[(cq * 4 + 0) * 2 + 1] = g1.x;
    o[(cq * 4 + 1) * 2] = g0.y; o[(cq * 4 + 1) * 2 + 1] = g1.y;
    o[(cq * 4 + 2) * 2] = g0.z; o[(cq * 4 + 2) * 2 + 1] = g1.z;
    o[(cq * 4 + 3) * 2] = g0.w; o[(cq * 4 + 3) * 2 + 1] = g1.w;
    if (cq == 0) {
      for (int j = 1; j < R; ++j) { sb0 += sbm[j * 2]; sb1 += sbm[j * 2 + 1]; }
      o[C * 2] = sb0;
      o[C * 2 + 1] = sb1;
    }
  }
}
void head_bwd(const float* x, int ldx, const float* w, const float* drop_mask, float keep_scale, const float* dz,
              float* dx, int lddx, float* partials, float* dw, float* db, int M, int C, cudaStream_t s) {
  // dw and db are contiguous in theta? not necessarily: reduce them separately from one partial buffer
  int c4 = C / 4, R = 256 / c4;
  if (R < 1) R = 1;
  if (R > 64) R = 64;
  int G = cdiv(M, R * 8);
  if (G > 296) G = 296;
  dim3 blk(c4, R);
  size_t smem = 2 * (size_t)R * c4 * sizeof(float4) + 2 * R * sizeof(float);
  MLIIS_COUNT(), head_bwd_kernel<<<dim3(G, 1, MLIIS_NZ), blk, smem, s>>>(x, ldx, w, drop_mask, keep_scale, dz, dx, lddx, partials, M, C,
                                                                        cdiv(M, G), MLIIS_ZS);
  // partial rows are [C*2 | 2]; reduce into a staging area right after the partials, then scatter
  float* stage = partials + (size_t)G * (C * 2 + 2);
  reduce_partials(partials, G, C * 2 + 2, stage, s);
  for (int z = 0; z < MLIIS_NZ; ++z) {
    const size_t zo = (size_t)z * MLIIS_ZS;
    cudaMemcpyAsync(dw + zo, stage + zo, (size_t)C * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(db + zo, stage + zo + C * 2, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  }
}

// =============================================================================================
// loss: softmax CE (+ label smoothing) - ln(dice) ; fused with the final bilinear upsample
// =============================================================================================
struct Up2 { float z0, z1; };
__device__ __forceinline__ Up2 upsample_logits(const float* __restrict__ z_lo, int b, int h, int w, int Y, int X,
                                               const ResizeTab& ty, const ResizeTab& tx) {
  const int y0 = ty.lo[Y], y1 = ty.hi[Y], x0 = tx.lo[X], x1 = tx.hi[X];
  const float yl = ty.lerp[Y], xl = tx.lerp[X];
  const float2* base = reinterpret_cast<const float2*>(z_lo) + (size_t)b * h * w;
  const float2 tl = base[y0 * w + x0], tr = base[y0 * w + x1], bl = base[y1 * w + x0], br = base[y1 * w + x1];
  Up2 r;
  float top = tl.x + (tr.x - tl.x) * xl, bot = bl.x + (br.x - bl.x) * xl;
  r.z0 = top + (bot - top) * yl;
  top = tl.y + (tr.y - tl.y) * xl; bot = bl.y + (br.y - bl.y) * xl;
  r.z1 = top + (bot - top) * yl;
  return r;
}

constexpr int kLossChunks = 32;   // row chunks per image

__device__ __forceinline__ void loss_shift(LossArgs& a, long long zs) {
  const size_t zo = (size_t)blockIdx.z * zs;
  a.z_lo += zo; a.labels += zo; a.index = zp(a.index, zo); a.p1 += zo; a.partials += zo; a.coef += zo; a.dz_hi += zo;
  a.loss_out = zp(a.loss_out, zo); a.theta = zp(a.theta, zo);
}

__global__ void __launch_bounds__(256) loss_fwd_kernel(LossArgs a, long long zs) {
  __shared__ float red[8][4];
  loss_shift(a, zs);
  const int b = blockIdx.y, g = blockIdx.x;
  const int img = a.index ? a.index[b] : b;
  const int rows_per = (a.H + kLossChunks - 1) / kLossChunks;
  const int Y0 = g * rows_per, Y1 = min(a.H, Y0 + rows_per);
  const float ls = a.label_smoothing;
  float s_ce = 0.f, s_i = 0.f, s_p = 0.f, s_y = 0.f;
  const int npix = (Y1 - Y0) * a.W;
  constexpr int U = 4;      // pixels in flight per thread: all gathers / label loads of a batch are issued before the
  for (int i0 = threadIdx.x; i0 < npix; i0 += U * 256) {      // first store (the stores would otherwise fence them)
    Up2 z[U];
    float2 y[U];
    size_t po[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 256;
      if (i < npix) {
        const int Y = Y0 + i / a.W, X = i - (i / a.W) * a.W;
        z[u] = upsample_logits(a.z_lo, b, a.h, a.w, Y, X, a.ty, a.tx);
        y[u] = *reinterpret_cast<const float2*>(a.labels + (((size_t)img * a.H + Y) * a.W + X) * 2);
        po[u] = ((size_t)b * a.H + Y) * a.W + X;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * 256 < npix) {
        const float m = fmaxf(z[u].z0, z[u].z1);
        const float e0 = expf(z[u].z0 - m), e1 = expf(z[u].z1 - m);
        const float sum = e0 + e1;
        const float p1 = e1 / sum;
        const float lse = m + logf(sum);
        const float yc0 = y[u].x * (1.f - ls) + 0.5f * ls, yc1 = y[u].y * (1.f - ls) + 0.5f * ls;
        s_ce += yc0 * (lse - z[u].z0) + yc1 * (lse - z[u].z1);
        s_i += p1 * y[u].y;
        s_p += p1;
        s_y += y[u].y;
        a.p1[po[u]] = p1;
      }
    }
  }
  s_ce = warp_sum(s_ce); s_i = warp_sum(s_i); s_p = warp_sum(s_p); s_y = warp_sum(s_y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp][0] = s_ce; red[warp][1] = s_i; red[warp][2] = s_p; red[warp][3] = s_y; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    a.partials[((size_t)b * kLossChunks + g) * 4 + threadIdx.x] = s;
  }
}

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out /*[gridDim.x]*/,
                             long long zs) {
  __shared__ float red[8];
  { const size_t zo = (size_t)blockIdx.z * zs; x += zo; out += zo; }
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fmaf(x[i], x[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += red[wv];
    out[blockIdx.x] = t;
  }
}

void sumsq_partials(const float* x, int64_t n, float* out, int n_blocks, cudaStream_t s) {
  MLIIS_COUNT(), sumsq_kernel<<<dim3(n_blocks, 1, MLIIS_NZ), 256, 0, s>>>(x, n, out, MLIIS_ZS);
}

// one warp: per-image IoU, dice, loss value and the per-image gradient coefficients.  Lane g owns row chunk g of every
// image (kLossChunks == 32): one float4 load per image, the four sums by a fixed xor-shuffle tree in double (every lane
// ends up with the same totals; deterministic).  The first version was ONE thread walking 1024 dependent loads (19 us).
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void __launch_bounds__(32) loss_finalize_kernel(LossArgs a, const float* __restrict__ l2_partials,
                                                           int n_l2_partials, long long zs) {
  static_assert(kLossChunks == 32, "lane == row chunk");
  __shared__ double sI[256], sU[256];
  if (blockIdx.x != 0) return;
  loss_shift(a, zs);
  l2_partials += (size_t)blockIdx.z * zs;
  const int lane = threadIdx.x;
  const double eps = 1e-7;
  double ce = 0.0, iou = 0.0;
  for (int b0 = 0; b0 < a.B; b0 += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (b0 + u < a.B) v[u] = ld4(a.partials + ((size_t)(b0 + u) * kLossChunks + lane) * 4);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (b0 + u < a.B) {
        ce += warp_sum_d((double)v[u].x);
        const double I = warp_sum_d((double)v[u].y), P = warp_sum_d((double)v[u].z), Yv = warp_sum_d((double)v[u].w);
        const double U = P + Yv - I;
        iou += (I + eps) / (U + eps);
        if (lane == 0 && b0 + u < 256) { sI[b0 + u] = I; sU[b0 + u] = U; }
      }
    }
  }
  __syncwarp();
  iou /= a.B;
  double loss = ce / ((double)a.B * a.H * a.W);
  double dLdiou = 0.0;
  if (a.dice) {
    loss -= log(2.0 * iou / (iou + 1.0));
    dLdiou = -1.0 / (iou * (iou + 1.0));
  }
  for (int b = lane; b < a.B; b += 32) {
    const double I = sI[b], U = sU[b];
    a.coef[b * 2 + 0] = (float)(dLdiou / (a.B * (U + eps)));                       // d loss / d I_b
    a.coef[b * 2 + 1] = (float)(-dLdiou * (I + eps) / (a.B * (U + eps) * (U + eps)));  // d loss / d U_b
  }
  if (a.loss_out) {
    double l2 = 0.0;
    for (int i = lane; i < n_l2_partials; i += 32) l2 += (double)l2_partials[i];
    l2 = warp_sum_d(l2);
    if (lane == 0) *a.loss_out = (float)(loss + 0.5 * (double)a.l2_coef * l2);
  }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(LossArgs a, long long zs) {
  loss_shift(a, zs);
  const int b = blockIdx.y;
  const int img = a.index ? a.index[b] : b;
  const float cI = a.coef[b * 2 + 0], cU = a.coef[b * 2 + 1];
  const float inv = 1.f / ((float)a.B * (float)a.H * (float)a.W);
  const float ls = a.label_smoothing;
  const int npix = a.H * a.W;
  constexpr int U = 4;
  const int step = gridDim.x * 256;
  for (int i0 = blockIdx.x * 256 + threadIdx.x; i0 < npix; i0 += U * step) {
    float p1v[U];
    float2 yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * step;
      if (i < npix) {
        p1v[u] = a.p1[(size_t)b * npix + i];
        yv[u] = *reinterpret_cast<const float2*>(a.labels + ((size_t)img * npix + i) * 2);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * step;
      if (i < npix) {
        const float p1 = p1v[u];
        const float2 y = yv[u];
        const float yc0 = y.x * (1.f - ls) + 0.5f * ls, yc1 = y.y * (1.f - ls) + 0.5f * ls;
        // d loss / d p1 = cI * y1 + cU * (1 - y1)      (dI/dp1 = y1, dU/dp1 = 1 - y1)
        const float t = (cI * y.y + cU * (1.f - y.y)) * p1 * (1.f - p1);
        // TF SoftmaxCrossEntropyWithLogits backprop = softmax - labels  [TF-ext]
        const float d1 = (p1 - yc1) * inv + t;
        const float d0 = ((1.f - p1) - yc0) * inv - t;
        *reinterpret_cast<float2*>(a.dz_hi + ((size_t)b * npix + i) * 2) = make_float2(d0, d1);
      }
    }
  }
}

void loss_fwd_bwd(const LossArgs& a, cudaStream_t s) {
  const int nz = MLIIS_NZ;
  const long long zs = MLIIS_ZS;
  MLIIS_COUNT(), loss_fwd_kernel<<<dim3(kLossChunks, a.B, nz), 256, 0, s>>>(a, zs);
  float* l2p = a.partials + (size_t)a.B * kLossChunks * 4;
  int nl2 = 0;
  if (a.loss_out && a.l2_coef != 0.f && a.n_l2 > 0) {
    nl2 = 148;
    MLIIS_COUNT(), sumsq_kernel<<<dim3(nl2, 1, nz), 256, 0, s>>>(a.theta, a.n_l2, l2p, zs);
  }
  MLIIS_COUNT(), loss_finalize_kernel<<<dim3(1, 1, nz), 32, 0, s>>>(a, l2p, nl2, zs);
  MLIIS_COUNT(), loss_bwd_kernel<<<dim3(cdiv(a.H * a.W, 256 * 4), a.B, nz), 256, 0, s>>>(a, zs);
}

// =============================================================================================
// predictions: float(p > 0.5) on both channels + integer IoU counts on channel 1
// =============================================================================================
__global__ void __launch_bounds__(256) predict_kernel(const float* __restrict__ z_lo, const float* __restrict__ labels,
                                                       const int32_t* __restrict__ index, int h, int w, int H, int W,
                                                       ResizeTab ty, ResizeTab tx, float* __restrict__ pred,
                                                       float* __restrict__ logits, uint32_t* __restrict__ inter,
                                                       uint32_t* __restrict__ uni, long long zs) {
  __shared__ uint32_t red[8][2];
  {
    const size_t zo = (size_t)blockIdx.z * zs;
    z_lo += zo; labels = zp(labels, zo); index = zp(index, zo); pred = zp(pred, zo); logits = zp(logits, zo);
    inter = zp(inter, zo); uni = zp(uni, zo);
  }
  const int b = blockIdx.y;
  const int img = index ? index[b] : b;
  const int npix = H * W;
  uint32_t ci = 0, cu = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < npix; i += gridDim.x * 256) {
    const int Y = i / W, X = i - Y * W;
    const Up2 z = upsample_logits(z_lo, b, h, w, Y, X, ty, tx);
    const float m = fmaxf(z.z0, z.z1);
    const float e0 = expf(z.z0 - m), e1 = expf(z.z1 - m);
    const float sum = e0 + e1;
    const bool q0 = e0 / sum > 0.5f, q1 = e1 / sum > 0.5f;
    if (pred) *reinterpret_cast<float2*>(pred + ((size_t)b * npix + i) * 2) = make_float2(q0 ? 1.f : 0.f, q1 ? 1.f : 0.f);
    if (logits) *reinterpret_cast<float2*>(logits + ((size_t)b * npix + i) * 2) = make_float2(z.z0, z.z1);
    if (labels) {
      // np.round(label) on [0,1] data: 1 iff label > 0.5 (round-half-even sends 0.5 to 0)
      const bool l1 = labels[((size_t)img * npix + i) * 2 + 1] > 0.5f;
      ci += (q1 && l1) ? 1u : 0u;
      cu += (q1 || l1) ? 1u : 0u;
    }
  }
  if (inter) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ci += __shfl_xor_sync(0xffffffffu, ci, o);
      cu += __shfl_xor_sync(0xffffffffu, cu, o);
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = ci; red[threadIdx.x >> 5][1] = cu; }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t ti = 0, tu = 0;
      for (int wv = 0; wv < 8; ++wv) { ti += red[wv][0]; tu += red[wv][1]; }
      atomicAdd(inter + b, ti);
      atomicAdd(uni + b, tu);
    }
  }
}
void predict_mask_iou(const float* z_lo, const float* labels, const int32_t* index, int B, int h, int w, int H, int W,
                      ResizeTab ty, ResizeTab tx, float* pred_out, float* logits_out, uint32_t* inter, uint32_t* uni,
                      cudaStream_t s) {
  if (inter) {
    for (int z = 0; z < MLIIS_NZ; ++z) {
      cudaMemsetAsync(inter + (size_t)z * MLIIS_ZS, 0, B * sizeof(uint32_t), s);
      cudaMemsetAsync(uni + (size_t)z * MLIIS_ZS, 0, B * sizeof(uint32_t), s);
    }
  }
  MLIIS_COUNT(), predict_kernel<<<dim3(cdiv(H * W, 256 * 4), B, MLIIS_NZ), 256, 0, s>>>(z_lo, inter ? labels : nullptr, index, h, w, H, W,
                                                                                       ty, tx, pred_out, logits_out, inter, uni,
                                                                                       MLIIS_ZS);
}

// =============================================================================================
// optimizer + meta-update over the flat parameter buffer
// =============================================================================================
__global__ void scale_kernel(float* __restrict__ x, int64_t n, float s, long long zs) {
  x += (size_t)blockIdx.z * zs;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i * 4 + 3 < n) st4(x + i * 4, ld4(x + i * 4) * s);
  else for (int64_t j = i * 4; j < n; ++j) x[j] *= s;
}
void scale_buffer(float* x, int64_t n, float sc, cudaStream_t st) {
  MLIIS_COUNT(), scale_kernel<<<dim3((unsigned)cdiv64(cdiv64(n, 4), 256), 1, MLIIS_NZ), 256, 0, st>>>(x, n, sc, MLIIS_ZS);
}

// TF ApplyAdam with beta1 = 0 (m == g) / ApplyGradientDescent.  g' = g + l2*theta on the first n_l2 floats
// (gradient of 0.0005 * sum l2_loss(v), regularizers.py:4-10).  powers = {beta1_power, beta2_power}.
__device__ __forceinline__ void adam_elem(float& th, float& v, float g, bool l2, float l2_coef, float lr, float alpha, int sgd) {
  const float gi = g + (l2 ? l2_coef * th : 0.f);
  if (sgd) {
    th = th - lr * gi;
  } else {
    const float b2 = 0.999f, eps = 1e-8f;
    const float vi = b2 * v + (1.f - b2) * gi * gi;
    v = vi;
    th = th - alpha * gi / (sqrtf(vi) + eps);
  }
}
// One pass over the flat buffer: 20 bytes per parameter (Adam) / 12 (SGD).  A thread owns kAdamU float4 words spaced a
// block apart and issues all of its loads before the first use (HBM-bound: bytes in flight per SM are what matter).
constexpr int kAdamU = 4;
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ theta, float* __restrict__ v,
                                                   const float* __restrict__ g, int64_t n, int64_t n_l2,
                                                   const float* __restrict__ lr_dev, const float* __restrict__ powers,
                                                   float l2_coef, int sgd, long long zs, float lr_imm,
                                                   float b2p_imm) {
  { const size_t zo = (size_t)blockIdx.z * zs; theta += zo; v += zo; g += zo; lr_dev = zp(lr_dev, zo); powers = zp(powers, zo); }
  // lr_dev / powers null: immediate scalars (per-kernel entry point; beta1_power = 0)
  const float lr = lr_dev ? *lr_dev : lr_imm;
  const float alpha = sgd ? lr : lr * sqrtf(1.f - (powers ? powers[1] : b2p_imm)) / (1.f - (powers ? powers[0] : 0.f));
  const int64_t n4 = n >> 2;
  const int64_t base = (int64_t)blockIdx.x * (256 * kAdamU) + threadIdx.x;
  float4 th[kAdamU], gg[kAdamU], vv[kAdamU];
#pragma unroll
  for (int u = 0; u < kAdamU; ++u) {
    const int64_t i = base + u * 256;
    if (i < n4) {
      th[u] = ld4(theta + 4 * i);
      gg[u] = ld4(g + 4 * i);
      vv[u] = sgd ? f4s(0.f) : ld4(v + 4 * i);
    }
  }
#pragma unroll
  for (int u = 0; u < kAdamU; ++u) {
    const int64_t i = base + u * 256;
    if (i < n4) {
      const int64_t e = 4 * i;
      adam_elem(th[u].x, vv[u].x, gg[u].x, e + 0 < n_l2, l2_coef, lr, alpha, sgd);
      adam_elem(th[u].y, vv[u].y, gg[u].y, e + 1 < n_l2, l2_coef, lr, alpha, sgd);
      adam_elem(th[u].z, vv[u].z, gg[u].z, e + 2 < n_l2, l2_coef, lr, alpha, sgd);
      adam_elem(th[u].w, vv[u].w, gg[u].w, e + 3 < n_l2, l2_coef, lr, alpha, sgd);
      st4(theta + e, th[u]);
      if (!sgd) st4(v + e, vv[u]);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {      // tail of a buffer whose length is not a multiple of 4
    const int64_t e = 4 * n4 + threadIdx.x;
    float t = theta[e], w = sgd ? 0.f : v[e];
    adam_elem(t, w, g[e], e < n_l2, l2_coef, lr, alpha, sgd);
    theta[e] = t;
    if (!sgd) v[e] = w;
  }
}
__global__ void adam_finish_kernel(float* powers, long long zs) {
  powers += (size_t)blockIdx.z * zs;
  powers[0] *= 0.0f;      // beta1 = 0
  powers[1] *= 0.999f;
}
void adam_step(float* theta, float* v, const float* g, int64_t n, int64_t n_l2, const float* lr_dev, float* powers,
               float l2_coef, int sgd, cudaStream_t s, float lr_imm, float b2p_imm) {
  MLIIS_COUNT(), adam_kernel<<<dim3((unsigned)cdiv64(cdiv64(n, 4), 256 * kAdamU), 1, MLIIS_NZ), 256, 0, s>>>(
      theta, v, g, n, n_l2, lr_dev, powers, l2_coef, sgd, MLIIS_ZS, lr_imm, b2p_imm);
  if (!sgd && powers) MLIIS_COUNT(), adam_finish_kernel<<<dim3(1, 1, MLIIS_NZ), 1, 0, s>>>(powers, MLIIS_ZS);
}

__global__ void delta_acc_kernel(float* __restrict__ d, const float* __restrict__ a, const float* __restrict__ b,
                                 int64_t n, int first) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = a[i] - b[i];
  d[i] = first ? v : d[i] + v;
}
void delta_accumulate(float* dsum, const float* a, const float* b, int64_t n, int first, cudaStream_t s) {
  MLIIS_COUNT(), delta_acc_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, s>>>(dsum, a, b, n, first);
}
__global__ void meta_apply_kernel(float* __restrict__ th, const float* __restrict__ d, float scale, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) th[i] = fmaf(scale, d[i], th[i]);
}
void meta_apply(float* theta, const float* dsum, float scale, int64_t n, cudaStream_t s) {
  MLIIS_COUNT(), meta_apply_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, s>>>(theta, dsum, scale, n);
}

// Keras Dropout keep mask: keep where U >= rate [TF-ext]; U from a counter-based hash (the reference's
// stream is unseeded, so any uniform stream is a valid draw; tests inject masks instead).
// seed_dev: optional device scalar added to the host seed at run time (CUDA-graph replays draw fresh masks).
__global__ void dropout_mask_kernel(float* __restrict__ mask, int64_t n, float rate, uint64_t seed,
                                    const uint64_t* __restrict__ seed_dev, long long zs) {
  mask += (size_t)blockIdx.z * zs;
  if (seed_dev) seed_dev += ((size_t)blockIdx.z * zs) / 2;       // zs is in floats; the seed is a 64-bit scalar
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (seed_dev) seed += *seed_dev * 0xD1B54A32D192ED03ull;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.f / 16777216.f);
  mask[i] = u >= rate ? 1.f : 0.f;
}
void fill_dropout_mask(float* mask, int64_t n, float rate, uint64_t seed, const uint64_t* seed_dev, cudaStream_t s) {
  MLIIS_COUNT(), dropout_mask_kernel<<<dim3((unsigned)cdiv64(n, 256), 1, MLIIS_NZ), 256, 0, s>>>(mask, n, rate, seed, seed_dev, MLIIS_ZS);
}

}  // namespace mliis
