// Multi-class head loss for the supervised joint-training path (BASELINE config 5).
//
// Reference: EfficientLab(n_classes=1000, seperate_background_channel=True, binary_iou_loss=False)
//   logits [B,H,W,C] (C = 1001) = bilinear_ac(head(features) [B,h,w,C])          models/efficientlab.py:165-172
//   loss = mean_pixels softmax-CE(labels, logits) - ln(2*IoU/(IoU+1)) (+ L2)          :294-327
//   IoU  = mean_b (I_b + eps) / (sum p + sum y - I_b + eps) over ALL channels            :355-396
// The reference materialises [B,224,224,1001] float tensors several times (200 MB per image each) and feeds dense
// one-hot uint8 labels (50 MB per image).  Here the labels stay sparse — a class id per image and a binary mask — and
// nothing of size H*W*C ever exists: the loss kernels recompute the upsampled logits from the low-resolution head
// output on the fly (each hi-res pixel is a bilinear blend of 4 low-res rows of C logits).
//   target(b, Y, X) = mask[b,Y,X] > 0.5 ? cls[b] : 0         (channel 0 = background)
//   sum_c y = sum_c p = 1 per pixel  =>  IoU_b = (I_b + eps) / (2*H*W - I_b + eps),  I_b = sum_pixels p[target]
#include "common.cuh"
#include "kernels.h"

namespace mliis {

namespace {

constexpr int kMcRowsPerBlock = 4;     // hi-res rows per block in the forward / predict kernels

struct HiPix {          // bilinear source of one hi-res pixel
  int y0, y1, x0, x1;
  float yl, xl;
};
__device__ __forceinline__ HiPix hi_pix(const ResizeTab& ty, const ResizeTab& tx, int Y, int X) {
  HiPix h;
  h.y0 = ty.lo[Y]; h.y1 = ty.hi[Y]; h.x0 = tx.lo[X]; h.x1 = tx.hi[X];
  h.yl = ty.lerp[Y]; h.xl = tx.lerp[X];
  return h;
}
// same operation order as upsample_logits (k_misc.cu) and the oracle's resize_bilinear_ac
__device__ __forceinline__ float lerp4(float tl, float tr, float bl, float br, float xl, float yl) {
  const float top = tl + (tr - tl) * xl, bot = bl + (br - bl) * xl;
  return top + (bot - top) * yl;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct PixStats { float lse, zt, zsum, zmax; int amax; };

// one warp, one hi-res pixel: log-sum-exp over C channels of the upsampled logits (+ target logit, sum, argmax)
__device__ __forceinline__ PixStats pixel_stats(const float* __restrict__ zb, int w, int ldz, int C, const HiPix& h,
                                                int target, bool want_sum, bool want_argmax) {
  const int lane = threadIdx.x & 31;
  const float* r00 = zb + (size_t)(h.y0 * w + h.x0) * ldz;
  const float* r01 = zb + (size_t)(h.y0 * w + h.x1) * ldz;
  const float* r10 = zb + (size_t)(h.y1 * w + h.x0) * ldz;
  const float* r11 = zb + (size_t)(h.y1 * w + h.x1) * ldz;
  float m = -INFINITY, s = 0.f, zs = 0.f;
  int am = 0;
  for (int c = lane; c < C; c += 32) {
    const float z = lerp4(r00[c], r01[c], r10[c], r11[c], h.xl, h.yl);
    if (want_sum) zs += z;
    if (z > m) {                      // online softmax: rescale the running sum when the max moves
      s = s * __expf(m - z) + 1.f;
      m = z;
      am = c;
    } else {
      s += __expf(z - m);
    }
  }
  const float M = warp_max(m);
  s = warp_sum(s * __expf(m - M));
  PixStats ps;
  ps.lse = M + logf(s);
  ps.zmax = M;
  ps.zsum = want_sum ? warp_sum(zs) : 0.f;
  ps.zt = lerp4(r00[target], r01[target], r10[target], r11[target], h.xl, h.yl);
  ps.amax = 0;
  if (want_argmax) {                  // smallest channel index among the maxima (np.argmax convention)
    int cand = (m == M) ? am : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    ps.amax = cand;
  }
  return ps;
}

__device__ __forceinline__ int target_of(const McLossArgs& a, int img, int Y, int X) {
  const float m = a.mask[((size_t)img * a.H + Y) * a.W + X];
  return m > 0.5f ? a.cls[img] : 0;
}

// grid (ceil(H / kMcRowsPerBlock), B), 256 threads: each warp walks hi-res pixels of the block's rows
__global__ void __launch_bounds__(256) mc_loss_fwd_kernel(McLossArgs a) {
  __shared__ float red[8][2];
  const int b = blockIdx.y, img = a.index ? a.index[b] : b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Y0 = blockIdx.x * kMcRowsPerBlock, Y1 = min(a.H, Y0 + kMcRowsPerBlock);
  const float* zb = a.z_lo + (size_t)b * a.h * a.w * a.ldz;
  const float ls = a.label_smoothing;
  float s_ce = 0.f, s_pt = 0.f;
  const int npix = (Y1 - Y0) * a.W;
  for (int i = warp; i < npix; i += 8) {
    const int Y = Y0 + i / a.W, X = i - (i / a.W) * a.W;
    const HiPix h = hi_pix(a.ty, a.tx, Y, X);
    const int t = target_of(a, img, Y, X);
    const PixStats ps = pixel_stats(zb, a.w, a.ldz, a.C, h, t, ls > 0.f, false);
    if (lane == 0) {
      const float pt = __expf(ps.zt - ps.lse);
      // -sum_c y_c log p_c with y = onehot*(1-ls) + ls/C
      s_ce += ps.lse - (1.f - ls) * ps.zt - (ls > 0.f ? ls / (float)a.C * ps.zsum : 0.f);
      s_pt += pt;
      const size_t o = ((size_t)b * a.H + Y) * a.W + X;
      a.lse[o] = ps.lse;
      a.pt[o] = pt;
    }
  }
  if (lane == 0) { red[warp][0] = s_ce; red[warp][1] = s_pt; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    a.partials[((size_t)b * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = s;
  }
}

// single thread: per-image soft IoU over all channels, dice, loss value, d loss / d I_b
__global__ void mc_loss_finalize_kernel(McLossArgs a, int chunks, const float* __restrict__ l2_partials,
                                        int n_l2_partials) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double eps = 1e-7, D = 2.0 * (double)a.H * (double)a.W;
  double ce = 0.0, iou = 0.0;
  for (int b = 0; b < a.B; ++b) {
    double I = 0.0;
    for (int g = 0; g < chunks; ++g) {
      ce += a.partials[((size_t)b * chunks + g) * 2 + 0];
      I += a.partials[((size_t)b * chunks + g) * 2 + 1];
    }
    iou += (I + eps) / (D - I + eps);
    a.coef[b] = (float)I;
  }
  iou /= a.B;
  double loss = ce / ((double)a.B * a.H * a.W);
  double dLdiou = 0.0;
  if (a.dice) {
    loss -= log(2.0 * iou / (iou + 1.0));
    dLdiou = -1.0 / (iou * (iou + 1.0));
  }
  for (int b = 0; b < a.B; ++b) {
    const double I = a.coef[b], den = D - I + eps;
    a.coef[b] = (float)(dLdiou / a.B * (D + 2.0 * eps) / (den * den));     // d loss / d I_b
  }
  if (a.loss_out) {
    double l2 = 0.0;
    for (int i = 0; i < n_l2_partials; ++i) l2 += l2_partials[i];
    *a.loss_out = (float)(loss + 0.5 * (double)a.l2_coef * l2);
  }
}

// Backward straight to the LOW-resolution logits (adjoint of the bilinear upsample fused in):
//   dz_lo[b,y,x,c] = sum_{(Y,X) in footprint(y,x)} wy*wx * ( inv*(p_c - y_c) + coefI_b * p_t * ([c==t] - p_c) )
// with p_c = exp(z_c(Y,X) - lse(Y,X)) recomputed from the 3x3 low-res neighbourhood staged in shared memory.
// One block per low-res pixel; threads run over channels.
struct FootPix { float wgt, lse, ptc, yl, xl; int t, o00, o01, o10, o11; };
constexpr int kMaxFoot = 100;

__global__ void __launch_bounds__(256) mc_loss_bwd_kernel(McLossArgs a) {
  extern __shared__ float sm[];
  __shared__ FootPix foot[kMaxFoot];
  __shared__ int nfoot_s;
  float* nb = sm;                         // [3][3][C] low-res neighbourhood (clamped at the borders)
  const int C = a.C;
  const int r = blockIdx.x;
  const int b = r / (a.h * a.w), rem = r - b * (a.h * a.w), y = rem / a.w, x = rem - y * a.w;
  const int img = a.index ? a.index[b] : b;
  const float* zb = a.z_lo + (size_t)b * a.h * a.w * a.ldz;
  for (int j = 0; j < 9; ++j) {
    const int yy = min(max(y + j / 3 - 1, 0), a.h - 1), xx = min(max(x + j % 3 - 1, 0), a.w - 1);
    const float* src = zb + (size_t)(yy * a.w + xx) * a.ldz;
    for (int c = threadIdx.x; c < C; c += blockDim.x) nb[j * C + c] = src[c];
  }
  {
    // footprint candidates (Y, X) of this low-res pixel, one per thread, kept in a fixed order (deterministic sums);
    // candidates with zero weight stay in the list and are skipped uniformly
    const float coefI = a.coef[b];
    const int Ya = a.ty.g_lo[y], Yb = a.ty.g_hi[y], Xa = a.tx.g_lo[x], Xb = a.tx.g_hi[x];
    const int nX = Xb - Xa + 1, n = min((Yb - Ya + 1) * nX, kMaxFoot);
    if (threadIdx.x < n) {
      const int Y = Ya + threadIdx.x / nX, X = Xa + threadIdx.x % nX;
      const int y0 = a.ty.lo[Y], y1 = a.ty.hi[Y], x0 = a.tx.lo[X], x1 = a.tx.hi[X];
      const float yl = a.ty.lerp[Y], xl = a.tx.lerp[X];
      const float wy = (y0 == y ? 1.f - yl : 0.f) + (y1 == y ? yl : 0.f);
      const float wx = (x0 == x ? 1.f - xl : 0.f) + (x1 == x ? xl : 0.f);
      FootPix f;
      f.wgt = wy * wx; f.yl = yl; f.xl = xl;
      f.lse = 0.f; f.ptc = 0.f; f.t = 0; f.o00 = f.o01 = f.o10 = f.o11 = 0;
      if (f.wgt != 0.f) {
        const size_t o = ((size_t)b * a.H + Y) * a.W + X;
        f.lse = a.lse[o]; f.ptc = coefI * a.pt[o];
        f.t = target_of(a, img, Y, X);
        // neighbour (yy, xx) sits at ((yy - y + 1) * 3 + (xx - x + 1)) * C inside nb
        f.o00 = ((y0 - y + 1) * 3 + (x0 - x + 1)) * C; f.o01 = ((y0 - y + 1) * 3 + (x1 - x + 1)) * C;
        f.o10 = ((y1 - y + 1) * 3 + (x0 - x + 1)) * C; f.o11 = ((y1 - y + 1) * 3 + (x1 - x + 1)) * C;
      }
      foot[threadIdx.x] = f;
    }
    if (threadIdx.x == 0) nfoot_s = n;
  }
  __syncthreads();
  const int nfoot = nfoot_s;
  const float inv = 1.f / ((float)a.B * (float)a.H * (float)a.W);
  const float ls = a.label_smoothing, ybase = ls / (float)C;
  float* out = a.dz_lo + (size_t)r * a.lddz;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < nfoot; ++k) {
      const FootPix& f = foot[k];
      if (f.wgt == 0.f) continue;
      const float z = lerp4(nb[f.o00 + c], nb[f.o01 + c], nb[f.o10 + c], nb[f.o11 + c], f.xl, f.yl);
      const float p = __expf(z - f.lse);
      const float hit = (c == f.t) ? 1.f : 0.f;
      // TF SoftmaxCrossEntropyWithLogits backprop = softmax - labels [TF-ext]; dI/dz_c = p_t * ([c==t] - p_c)
      const float d = inv * (p - (hit * (1.f - ls) + ybase)) + f.ptc * (hit - p);
      acc = fmaf(f.wgt, d, acc);
    }
    out[c] = acc;
  }
  for (int c = C + threadIdx.x; c < a.lddz; c += blockDim.x) out[c] = 0.f;     // padded columns
}

// predictions: class map (argmax if its probability > 0.5, else -1) + integer IoU counts over all channels
__global__ void __launch_bounds__(256) mc_predict_kernel(McLossArgs a, int32_t* __restrict__ class_map,
                                                         uint32_t* __restrict__ inter, uint32_t* __restrict__ uni) {
  __shared__ unsigned int cnt[8][2];
  const int b = blockIdx.y, img = a.index ? a.index[b] : b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Y0 = blockIdx.x * kMcRowsPerBlock, Y1 = min(a.H, Y0 + kMcRowsPerBlock);
  const float* zb = a.z_lo + (size_t)b * a.h * a.w * a.ldz;
  unsigned int n_hit = 0, n_pred = 0;
  const int npix = (Y1 - Y0) * a.W;
  for (int i = warp; i < npix; i += 8) {
    const int Y = Y0 + i / a.W, X = i - (i / a.W) * a.W;
    const HiPix h = hi_pix(a.ty, a.tx, Y, X);
    const int t = a.mask ? target_of(a, img, Y, X) : 0;
    const PixStats ps = pixel_stats(zb, a.w, a.ldz, a.C, h, t, false, true);
    if (lane == 0) {
      const float pmax = __expf(ps.zmax - ps.lse);
      const int pred = pmax > 0.5f ? ps.amax : -1;          // float(p > 0.5) is one-hot or all-zero per pixel
      if (class_map) class_map[((size_t)b * a.H + Y) * a.W + X] = pred;
      n_pred += pred >= 0;
      n_hit += (a.mask && pred == t);
    }
  }
  if (lane == 0) { cnt[warp][0] = n_hit; cnt[warp][1] = n_pred; }
  __syncthreads();
  if (threadIdx.x == 0 && inter && uni) {
    unsigned int hsum = 0, psum = 0;
    for (int wv = 0; wv < 8; ++wv) { hsum += cnt[wv][0]; psum += cnt[wv][1]; }
    // |pred AND label| = hits ; |pred OR label| = (one label per pixel) + preds - hits     (integer, order-free)
    atomicAdd(inter + b, hsum);
    atomicAdd(uni + b, (unsigned int)npix + psum - hsum);
  }
}

__global__ void mc_pad_head_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ wp,
                                   float* __restrict__ bp, int K, int C, int Cp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * Cp) {
    const int k = i / Cp, c = i - k * Cp;
    wp[i] = c < C ? w[(size_t)k * C + c] : 0.f;
  }
  if (i < Cp) bp[i] = i < C ? bias[i] : 0.f;
}
__global__ void mc_unpad_grad_kernel(const float* __restrict__ gwp, const float* __restrict__ gbp,
                                     float* __restrict__ gw, float* __restrict__ gb, int K, int C, int Cp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * C) {
    const int k = i / C, c = i - k * C;
    gw[i] = gwp[(size_t)k * Cp + c];
  }
  if (i < C) gb[i] = gbp[i];
}
__global__ void mc_mul_mask_kernel(const float* __restrict__ x, const float* __restrict__ mask, float scale,
                                   float* __restrict__ y, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) st4(y + i * 4, ld4(x + i * 4) * ld4(mask + i * 4) * scale);
}

}  // namespace

int mc_loss_chunks(int H) { return cdiv(H, kMcRowsPerBlock); }

void mc_loss_fwd_bwd(const McLossArgs& a, cudaStream_t s) {
  const int chunks = mc_loss_chunks(a.H);
  MLIIS_COUNT(), mc_loss_fwd_kernel<<<dim3(chunks, a.B), 256, 0, s>>>(a);
  float* l2p = a.partials + (size_t)a.B * chunks * 2;
  int nl2 = 0;
  if (a.loss_out && a.l2_coef != 0.f && a.n_l2 > 0) {
    nl2 = 148;
    sumsq_partials(a.theta, a.n_l2, l2p, nl2, s);
  }
  MLIIS_COUNT(), mc_loss_finalize_kernel<<<1, 32, 0, s>>>(a, chunks, l2p, nl2);
  const size_t smem = (size_t)9 * a.C * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(mc_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  MLIIS_COUNT(), mc_loss_bwd_kernel<<<a.B * a.h * a.w, 256, smem, s>>>(a);
}

void mc_predict(const McLossArgs& a, int32_t* class_map, uint32_t* inter, uint32_t* uni, cudaStream_t s) {
  if (inter && uni) {
    cudaMemsetAsync(inter, 0, a.B * sizeof(uint32_t), s);
    cudaMemsetAsync(uni, 0, a.B * sizeof(uint32_t), s);
  }
  MLIIS_COUNT(), mc_predict_kernel<<<dim3(mc_loss_chunks(a.H), a.B), 256, 0, s>>>(a, class_map, inter, uni);
}

void mc_pad_head(const float* w, const float* bias, float* wp, float* bp, int K, int C, int Cp, cudaStream_t s) {
  MLIIS_COUNT(), mc_pad_head_kernel<<<cdiv(K * Cp, 256), 256, 0, s>>>(w, bias, wp, bp, K, C, Cp);
}
void mc_unpad_grad(const float* gwp, const float* gbp, float* gw, float* gb, int K, int C, int Cp, cudaStream_t s) {
  MLIIS_COUNT(), mc_unpad_grad_kernel<<<cdiv(K * C, 256), 256, 0, s>>>(gwp, gbp, gw, gb, K, C, Cp);
}
void mc_mul_mask(const float* x, const float* mask, float scale, float* y, int64_t n, cudaStream_t s) {
  MLIIS_COUNT(), mc_mul_mask_kernel<<<(unsigned)cdiv64(n / 4, 256), 256, 0, s>>>(x, mask, scale, y, n / 4);
}

}  // namespace mliis
