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

// Executed by a full warp: the contiguous run of hi-res indices whose lo (floor) source index is `c`.
__device__ __forceinline__ void cell_run(const ResizeTab& t, int c, int& first, int& count) {
  const int lane = threadIdx.x & 31;
  const int a0 = t.g_lo[c], a1 = t.g_hi[c];
  const bool mine = (a0 + lane <= a1) && t.lo[a0 + lane] == c;      // gather ranges are <= 10 wide for H/h ~ 4
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  first = a0 + __ffs(m) - 1;
  count = __popc(m);
}

// One block per low-res CELL (cy,cx) = the hi-res pixels whose top-left bilinear source is (cy,cx).  The four corner
// rows of logits (4 x C floats) are staged in shared memory once, so the low-res logits are read from global memory
// exactly once per step; each warp then walks the cell's hi-res pixels (log-sum-exp over C channels per pixel).
__global__ void __launch_bounds__(256) mc_loss_fwd_kernel(McLossArgs a) {
  extern __shared__ float corners[];          // [4][C]
  __shared__ float red[8][2];
  const int C = a.C;
  const int r = blockIdx.x;
  const int b = r / (a.h * a.w), rem = r - b * (a.h * a.w), cy = rem / a.w, cx = rem - cy * a.w;
  const int img = a.index ? a.index[b] : b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y1 = min(cy + 1, a.h - 1), x1 = min(cx + 1, a.w - 1);
  const float* zb = a.z_lo + (size_t)b * a.h * a.w * a.ldz;
  const float* src[4] = {zb + (size_t)(cy * a.w + cx) * a.ldz, zb + (size_t)(cy * a.w + x1) * a.ldz,
                         zb + (size_t)(y1 * a.w + cx) * a.ldz, zb + (size_t)(y1 * a.w + x1) * a.ldz};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    for (int c = threadIdx.x; c < C; c += 256) corners[j * C + c] = src[j][c];
  int Yf, nY, Xf, nX;
  cell_run(a.ty, cy, Yf, nY);
  cell_run(a.tx, cx, Xf, nX);
  __syncthreads();
  const float ls = a.label_smoothing;
  float s_ce = 0.f, s_pt = 0.f;
  for (int k = warp; k < nY * nX; k += 8) {
    const int Y = Yf + k / nX, X = Xf + k % nX;
    const float yl = a.ty.lerp[Y], xl = a.tx.lerp[X];
    const int t = target_of(a, img, Y, X);
    float m = -INFINITY, s = 0.f, zs = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float z = lerp4(corners[c], corners[C + c], corners[2 * C + c], corners[3 * C + c], xl, yl);
      zs += z;
      if (z > m) {                      // online softmax: rescale the running sum when the max moves
        s = s * __expf(m - z) + 1.f;
        m = z;
      } else {
        s += __expf(z - m);
      }
    }
    const float M = warp_max(m);
    s = warp_sum(s * __expf(m - M));
    if (ls > 0.f) zs = warp_sum(zs);
    if (lane == 0) {
      const float lse = M + logf(s);
      const float zt = lerp4(corners[t], corners[C + t], corners[2 * C + t], corners[3 * C + t], xl, yl);
      const float pt = __expf(zt - lse);
      // -sum_c y_c log p_c with y = onehot*(1-ls) + ls/C
      s_ce += lse - (1.f - ls) * zt - (ls > 0.f ? ls / (float)C * zs : 0.f);
      s_pt += pt;
      const size_t o = ((size_t)b * a.H + Y) * a.W + X;
      a.lse[o] = lse;
      a.pt[o] = pt;
    }
  }
  if (lane == 0) { red[warp][0] = s_ce; red[warp][1] = s_pt; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float sum = 0.f;
    for (int wv = 0; wv < 8; ++wv) sum += red[wv][threadIdx.x];
    a.partials[(size_t)r * 2 + threadIdx.x] = sum;       // [b][cell][2]
  }
}

// one block: thread b sums image b's cell partials (fixed order), thread 0 then forms IoU, dice, the loss value and
// d loss / d I_b
__global__ void __launch_bounds__(256) mc_loss_finalize_kernel(McLossArgs a, int chunks,
                                                               const float* __restrict__ l2_partials,
                                                               int n_l2_partials) {
  __shared__ double s_ce[256], s_I[256];
  const double eps = 1e-7, D = 2.0 * (double)a.H * (double)a.W;
  // warp w sums the cell partials of images w, w+8, ...: lanes stride over the cells, then a fixed-order butterfly
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < a.B; b += 8) {
    double ce = 0.0, I = 0.0;
    const float* p = a.partials + (size_t)b * chunks * 2;
    for (int g = lane; g < chunks; g += 32) { ce += p[2 * g]; I += p[2 * g + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ce += __shfl_xor_sync(0xffffffffu, ce, o);
      I += __shfl_xor_sync(0xffffffffu, I, o);
    }
    if (lane == 0) { s_ce[b] = ce; s_I[b] = I; }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double ce = 0.0, iou = 0.0;
  for (int b = 0; b < a.B; ++b) {
    ce += s_ce[b];
    iou += (s_I[b] + eps) / (D - s_I[b] + eps);
  }
  iou /= a.B;
  double loss = ce / ((double)a.B * a.H * a.W);
  double dLdiou = 0.0;
  if (a.dice) {
    loss -= log(2.0 * iou / (iou + 1.0));
    dLdiou = -1.0 / (iou * (iou + 1.0));
  }
  for (int b = 0; b < a.B; ++b) {
    const double den = D - s_I[b] + eps;
    a.coef[b] = (float)(dLdiou / a.B * (D + 2.0 * eps) / (den * den));     // d loss / d I_b
  }
  if (a.loss_out) {
    double l2 = 0.0;
    for (int i = 0; i < n_l2_partials; ++i) l2 += l2_partials[i];
    *a.loss_out = (float)(loss + 0.5 * (double)a.l2_coef * l2);
  }
}

// Backward straight to the LOW-resolution logits (adjoint of the bilinear upsample fused in), in two passes so that
// every (hi-res pixel, channel) pair is evaluated exactly once:
//   d(Y,X,c) = inv*(p_c - y_c) + coefI_b * p_t * ([c==t] - p_c),   p_c = exp(z_c(Y,X) - lse(Y,X))
//   pass 1 (one block per low-res CELL (cy,cx) = the hi-res pixels whose top-left source is (cy,cx); thread = channel):
//           the four corner logits of the channel live in registers, the cell's <= 36 hi-res pixels are walked once,
//           and the four corner sums  sum_k w_corner(k) * d(k,c)  go to cellgrad[cell][corner][c]
//   pass 2: dz_lo[y,x] = cell(y,x).c00 + cell(y,x-1).c01 + cell(y-1,x).c10 + cell(y-1,x-1).c11   (fixed order)
struct CellPix { float w00, w01, w10, w11; float lse, ptc, yl, xl; int t; };
constexpr int kMaxCell = 36;

__global__ void __launch_bounds__(256) mc_loss_bwd_cell_kernel(McLossArgs a, float* __restrict__ cellgrad) {
  __shared__ CellPix pix[kMaxCell];
  __shared__ int npix_s;
  const int C = a.C;
  const int r = blockIdx.x;
  const int b = r / (a.h * a.w), rem = r - b * (a.h * a.w), cy = rem / a.w, cx = rem - cy * a.w;
  const int img = a.index ? a.index[b] : b;
  {
    // the cell's hi-res pixels: (rows with lo == cy) x (columns with lo == cx), listed row-major (fixed order)
    int Yf, nY, Xf, nX;
    cell_run(a.ty, cy, Yf, nY);
    cell_run(a.tx, cx, Xf, nX);
    const int n = min(nY * nX, kMaxCell);
    if (threadIdx.x < n) {
      const int Y = Yf + threadIdx.x / nX, X = Xf + threadIdx.x % nX;
      const float yl = a.ty.lerp[Y], xl = a.tx.lerp[X];
      const size_t o = ((size_t)b * a.H + Y) * a.W + X;
      CellPix c;
      c.w00 = (1.f - yl) * (1.f - xl); c.w01 = (1.f - yl) * xl; c.w10 = yl * (1.f - xl); c.w11 = yl * xl;
      c.lse = a.lse[o]; c.ptc = a.coef[b] * a.pt[o]; c.yl = yl; c.xl = xl;
      c.t = target_of(a, img, Y, X);
      pix[threadIdx.x] = c;
    }
    if (threadIdx.x == 0) npix_s = n;
  }
  __syncthreads();
  const int npix = npix_s;
  const int y1 = min(cy + 1, a.h - 1), x1 = min(cx + 1, a.w - 1);
  const float* zb = a.z_lo + (size_t)b * a.h * a.w * a.ldz;
  const float* r00 = zb + (size_t)(cy * a.w + cx) * a.ldz;
  const float* r01 = zb + (size_t)(cy * a.w + x1) * a.ldz;
  const float* r10 = zb + (size_t)(y1 * a.w + cx) * a.ldz;
  const float* r11 = zb + (size_t)(y1 * a.w + x1) * a.ldz;
  const float inv = 1.f / ((float)a.B * (float)a.H * (float)a.W);
  const float ls = a.label_smoothing, ybase = ls / (float)C, yhit = 1.f - ls + ybase;
  float* out = cellgrad + (size_t)r * 4 * a.lddz;
  for (int c = threadIdx.x; c < a.lddz; c += blockDim.x) {
    float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
    if (c < C) {
      const float z00 = r00[c], z01 = r01[c], z10 = r10[c], z11 = r11[c];
      for (int k = 0; k < npix; ++k) {
        const CellPix& f = pix[k];
        const float z = lerp4(z00, z01, z10, z11, f.xl, f.yl);
        const float p = __expf(z - f.lse);
        const bool hit = c == f.t;
        // TF SoftmaxCrossEntropyWithLogits backprop = softmax - labels [TF-ext]; dI/dz_c = p_t * ([c==t] - p_c)
        const float d = inv * (p - (hit ? yhit : ybase)) + f.ptc * ((hit ? 1.f : 0.f) - p);
        g00 = fmaf(f.w00, d, g00); g01 = fmaf(f.w01, d, g01);
        g10 = fmaf(f.w10, d, g10); g11 = fmaf(f.w11, d, g11);
      }
    }
    out[c] = g00; out[a.lddz + c] = g01; out[2 * a.lddz + c] = g10; out[3 * a.lddz + c] = g11;
  }
}

// corner k of cell (cy,cx) is low-res pixel (cy + k/2, cx + k%2), clamped: at the last row / column the clamped
// corner coincides with the cell's own pixel and carries weight 0 there (lerp == 0), so it is simply skipped
__global__ void __launch_bounds__(256) mc_loss_bwd_gather_kernel(const float* __restrict__ cellgrad,
                                                                 float* __restrict__ dz_lo, int B, int h, int w,
                                                                 int ld4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * h * w * ld4;
  if (i >= total) return;
  const int c4 = (int)(i % ld4);
  const int64_t r = i / ld4;
  const int x = (int)(r % w), y = (int)((r / w) % h);
  const float4* cg = reinterpret_cast<const float4*>(cellgrad);
  auto cell = [&](int64_t rr, int corner) { return cg[((size_t)rr * 4 + corner) * ld4 + c4]; };
  float4 acc = cell(r, 0);
  if (x > 0) acc = acc + cell(r - 1, 1);
  if (y > 0) acc = acc + cell(r - w, 2);
  if (x > 0 && y > 0) acc = acc + cell(r - w - 1, 3);
  reinterpret_cast<float4*>(dz_lo)[(size_t)r * ld4 + c4] = acc;
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
  const int R = Y1 - Y0, nidx = ((a.W + 3) / 4) * 4 * R;
  for (int i = warp; i < nidx; i += 8) {
    const int cb = i / (4 * R), q = i - cb * (4 * R);
    const int wcols = min(4, a.W - cb * 4);
    const int Y = Y0 + q / wcols, X = cb * 4 + q % wcols;
    if (Y >= Y1) continue;
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
  const int chunks = a.h * a.w;                       // one partial pair per low-res cell
  const size_t smem_fwd = (size_t)4 * a.C * sizeof(float);
  MLIIS_COUNT(), mc_loss_fwd_kernel<<<a.B * chunks, 256, smem_fwd, s>>>(a);
  float* l2p = a.partials + (size_t)a.B * chunks * 2;
  int nl2 = 0;
  if (a.loss_out && a.l2_coef != 0.f && a.n_l2 > 0) {
    nl2 = 148;
    sumsq_partials(a.theta, a.n_l2, l2p, nl2, s);
  }
  MLIIS_COUNT(), mc_loss_finalize_kernel<<<1, 256, 0, s>>>(a, chunks, l2p, nl2);
  MLIIS_COUNT(), mc_loss_bwd_cell_kernel<<<a.B * a.h * a.w, 256, 0, s>>>(a, a.cellgrad);
  const int ld4 = a.lddz / 4;
  const int64_t total = (int64_t)a.B * a.h * a.w * ld4;
  MLIIS_COUNT(), mc_loss_bwd_gather_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, s>>>(a.cellgrad, a.dz_lo, a.B, a.h, a.w,
                                                                                    ld4);
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
