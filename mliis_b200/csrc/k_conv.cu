// Stem conv (normalise + 3x3 s2) and depthwise k x k convolutions, forward / dgrad / wgrad.
//
// Depthwise kernels are HBM-bound stencils.  Design: one CTA = one image x one output tile x 32
// channels.  The input tile (with halo) is staged ONCE in shared memory - the producing BatchNorm's
// normalise + swish is applied while staging (prologue fusion: the activated tensor never exists in
// HBM) - and every thread then produces a strip of 7 consecutive outputs for one float4 of channels,
// re-using each shared-memory operand across the strip (K + 6S loads for 7K FMAs per filter row).
// All spatial sizes of the canonical network (112, 56, 28, 14) are multiples of 7.
//
// Reference: efficientnet_model.py:190-196 (DepthwiseConv2D, SAME), :359-366 (stem), TF SAME padding
// is asymmetric for stride 2 (pad_lo = total/2) and is passed in as pad_t/pad_l.
#include "common.cuh"
#include "kernels.h"

namespace mliis {

// =============================================================================================
// stem
// =============================================================================================
__device__ __forceinline__ void load_norm_pixel(const float* px, float& r, float& g, float& b) {
  r = (px[0] - kMeanR) / kStdR;
  g = (px[1] - kMeanG) / kStdG;
  b = (px[2] - kMeanB) / kStdB;
}

__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ images,
                                                        const int32_t* __restrict__ index,
                                                        const float* __restrict__ w, float* __restrict__ y, int B,
                                                        int H, int W, int Ho, int Wo, int pad_t, int pad_l,
                                                        long long zs) {
  __shared__ __align__(16) float ws[27 * 32];
  { const size_t zo = (size_t)blockIdx.z * zs; images += zo; index = zp(index, zo); w += zo; y += zo; }
  for (int i = threadIdx.x; i < 27 * 32; i += 256) ws[i] = w[i];
  __syncthreads();
  const int p = blockIdx.x * 64 + (threadIdx.x >> 2);
  const int cg = threadIdx.x & 3;
  if (p >= B * Ho * Wo) return;
  const int b = p / (Ho * Wo), rem = p - b * (Ho * Wo), oy = rem / Wo, ox = rem - oy * Wo;
  const int img = index ? index[b] : b;
  const float* im = images + (size_t)img * H * W * 3;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - pad_t + ky;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - pad_l + kx;
      if (ix < 0 || ix >= W) continue;
      float v[3];
      load_norm_pixel(im + ((size_t)iy * W + ix) * 3, v[0], v[1], v[2]);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* wr = ws + ((ky * 3 + kx) * 3 + ci) * 32 + cg * 8;
        const float4 w0 = ld4(wr), w1 = ld4(wr + 4);
        acc[0] = fmaf(v[ci], w0.x, acc[0]); acc[1] = fmaf(v[ci], w0.y, acc[1]);
        acc[2] = fmaf(v[ci], w0.z, acc[2]); acc[3] = fmaf(v[ci], w0.w, acc[3]);
        acc[4] = fmaf(v[ci], w1.x, acc[4]); acc[5] = fmaf(v[ci], w1.y, acc[5]);
        acc[6] = fmaf(v[ci], w1.z, acc[6]); acc[7] = fmaf(v[ci], w1.w, acc[7]);
      }
    }
  }
  float* o = y + (size_t)p * 32 + cg * 8;
  st4(o, f4(acc[0], acc[1], acc[2], acc[3]));
  st4(o + 4, f4(acc[4], acc[5], acc[6], acc[7]));
}

void stem_fwd(const float* images, const int32_t* index, const float* w, float* y, int B, int H, int W, int Ho, int Wo,
              int pad_t, int pad_l, cudaStream_t s) {
  MLIIS_COUNT(), stem_fwd_kernel<<<dim3(cdiv(B * Ho * Wo, 64), 1, MLIIS_NZ), 256, 0, s>>>(images, index, w, y, B, H, W, Ho, Wo, pad_t,
                                                                                         pad_l, MLIIS_ZS);
}

constexpr int kStemWgPix = 512;   // output pixels per CTA
int stem_wgrad_blocks(int B, int Ho, int Wo) { return cdiv(B * Ho * Wo, kStemWgPix); }

// dW[k = tap*3+ci][co] partial per CTA.  288 threads = 3 k-thirds (9 of the 27 k each) x 3 pixel groups x 32 output
// channels: per staged pixel a thread loads one gradient and its 9 inputs (three 16-byte broadcast loads of a row padded to
// 12 floats per third) for 9 FMAs.  The first version (thread = 3 k x 1 channel over all 64 pixels: 4 shared loads per 3
// FMAs) was bound by the shared-memory pipe: 435 us for a 16-slot launch against 40 us of HBM time.
__global__ void __launch_bounds__(288) stem_wgrad_kernel(const float* __restrict__ images,
                                                          const int32_t* __restrict__ index,
                                                          const float* __restrict__ dy, float* __restrict__ partials,
                                                          int B, int H, int W, int Ho, int Wo, int pad_t, int pad_l,
                                                          long long zs) {
  __shared__ __align__(16) float xs[64][36];
  __shared__ float gs[64][32];
  __shared__ float red[3][27][32];
  __shared__ int4 pix[64];
  { const size_t zo = (size_t)blockIdx.z * zs; images += zo; index = zp(index, zo); dy += zo; partials += zo; }
  const int tid = threadIdx.x, co = tid & 31, wv = tid >> 5, kq = wv / 3, pg = wv - kq * 3;
  const int total = B * Ho * Wo;
  const int p_begin = blockIdx.x * kStemWgPix;
  float acc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = 0.f;
  for (int base = p_begin; base < min(total, p_begin + kStemWgPix); base += 64) {
    // pixel coordinates once per pixel (two divisions by run-time values), not once per staged element: the staging
    // loop below only divides by constants.  The per-element version spent most of the kernel's issue slots there.
    if (tid < 64) {
      const int p = base + tid;
      int4 d = make_int4(-1, 0, 0, 0);
      if (p < total) {
        const int b = p / (Ho * Wo), rem = p - b * (Ho * Wo), oy = rem / Wo, ox = rem - oy * Wo;
        d = make_int4(index ? index[b] : b, oy * 2 - pad_t, ox * 2 - pad_l, 0);
      }
      pix[tid] = d;
    }
    __syncthreads();
    for (int i = tid; i < 64 * 27; i += 288) {
      const int lp = i / 27, k = i - lp * 27;
      const int4 d = pix[lp];
      float v = 0.f;
      if (d.x >= 0) {
        const int tap = k / 3, ci = k - tap * 3, ky = tap / 3, kx = tap - ky * 3;
        const int iy = d.y + ky, ix = d.z + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const float raw = images[(((size_t)d.x * H + iy) * W + ix) * 3 + ci];
          const float mean = ci == 0 ? kMeanR : (ci == 1 ? kMeanG : kMeanB);
          const float sd = ci == 0 ? kStdR : (ci == 1 ? kStdG : kStdB);
          v = (raw - mean) / sd;
        }
      }
      xs[lp][(k / 9) * 12 + k % 9] = v;
    }
    for (int i = tid; i < 64 * 32; i += 288) {
      const int lp = i >> 5, c = i & 31;
      const int p = base + lp;
      gs[lp][c] = p < total ? dy[(size_t)p * 32 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int lp = pg; lp < 64; lp += 3) {
      const float g = gs[lp][co];
      const float4 x0 = ld4(&xs[lp][kq * 12]), x1 = ld4(&xs[lp][kq * 12 + 4]);
      const float x8 = xs[lp][kq * 12 + 8];
      acc[0] = fmaf(x0.x, g, acc[0]); acc[1] = fmaf(x0.y, g, acc[1]); acc[2] = fmaf(x0.z, g, acc[2]);
      acc[3] = fmaf(x0.w, g, acc[3]); acc[4] = fmaf(x1.x, g, acc[4]); acc[5] = fmaf(x1.y, g, acc[5]);
      acc[6] = fmaf(x1.z, g, acc[6]); acc[7] = fmaf(x1.w, g, acc[7]); acc[8] = fmaf(x8, g, acc[8]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) red[pg][kq * 9 + j][co] = acc[j];
  __syncthreads();
  float* o = partials + (size_t)blockIdx.x * 864;
  for (int i = tid; i < 864; i += 288) {
    const int k = i >> 5, c = i & 31;
    o[i] = (red[0][k][c] + red[1][k][c]) + red[2][k][c];      // fixed order over the pixel groups
  }
}

void stem_wgrad(const float* images, const int32_t* index, const float* dy, float* partials, float* dw, int B, int H,
                int W, int Ho, int Wo, int pad_t, int pad_l, cudaStream_t s) {
  int G = stem_wgrad_blocks(B, Ho, Wo);
  MLIIS_COUNT(), stem_wgrad_kernel<<<dim3(G, 1, MLIIS_NZ), 288, 0, s>>>(images, index, dy, partials, B, H, W, Ho, Wo, pad_t, pad_l,
                                                                       MLIIS_ZS);
  reduce_partials(partials, G, 864, dw, s);
}

// =============================================================================================
// depthwise
// =============================================================================================
constexpr int QC = 8;  // float4 channel groups per CTA (32 channels)

template <int K, int S>
struct DwGeom {
  static constexpr int TOX = 14;                        // output tile: 14 x 14 (stride 1), 14 x 7 (stride 2)
  static constexpr int TOY = (S == 1) ? 14 : 7;
  static constexpr int SPR = TOX / 7;                   // strips (7 outputs along x) per tile row
  static constexpr int NSTRIP = TOY * SPR;              // 28 | 14
  static constexpr int NT = 224;                        // every thread stages; NSTRIP * QC threads compute
  static constexpr int TIX = (TOX - 1) * S + K;         // staged input tile
  static constexpr int TIY = (TOY - 1) * S + K;
  static constexpr int NIN = 6 * S + K;                 // inputs per strip row
  static constexpr int NPIX = TIX * TIY;
  static constexpr size_t smem_bytes() { return (size_t)(NPIX + K * K) * QC * sizeof(float4); }
  static constexpr size_t wgrad_smem_bytes() { return (size_t)(NPIX + K * K * (NT / 32)) * QC * sizeof(float4); }
};

// stage swish(a*x+b) (or x when a == null) of the input tile into shared memory.  HBM-bound: a thread issues ALL of its
// global loads (U <= 12 float4) before the first one is consumed, so one memory round trip stages the whole tile.
template <int K, int S>
__device__ __forceinline__ void dw_stage_input(float4* tile, const float* __restrict__ x,
                                               const float* __restrict__ a, const float* __restrict__ b, int img,
                                               int H, int W, int C, int c0, int iy0, int ix0) {
  using G = DwGeom<K, S>;
  constexpr int TOT = G::NPIX * QC;
  constexpr int PER = (TOT + G::NT - 1) / G::NT;
  constexpr int U = PER <= 12 ? PER : (PER + 1) / 2;       // loads in flight per thread per round
  const int tid = threadIdx.x, q = tid % QC;     // NT is a multiple of QC: q is the same for every i of a thread
  const bool cvalid = c0 + q * 4 < C;
  float4 av = f4s(1.f), bv = f4s(0.f);
  if (a && cvalid) { av = ld4(a + c0 + q * 4); bv = ld4(b + c0 + q * 4); }
  const float* xb = x + (size_t)img * H * W * C + c0 + q * 4;
#pragma unroll 1
  for (int i0 = tid; i0 < TOT; i0 += U * G::NT) {
    float4 v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * G::NT;
      const int pix = i / QC, ly = pix / G::TIX, lx = pix - ly * G::TIX;
      const int gy = iy0 + ly, gx = ix0 + lx;
      ok[u] = i < TOT && cvalid && gy >= 0 && gy < H && gx >= 0 && gx < W;
      v[u] = ok[u] ? ld4(xb + ((size_t)gy * W + gx) * C) : f4s(0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * G::NT;
      if (i < TOT) tile[i] = (a && ok[u]) ? swish4(affine4(v[u], av, bv)) : v[u];
    }
  }
}

// FLIP: use w[K-1-ky][K-1-kx] (stride-1 dgrad == correlation with the flipped filter)
template <int K, int S, bool FLIP>
__global__ void __launch_bounds__(DwGeom<K, S>::NT, 4) dw_fwd_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ a,
                                                                   const float* __restrict__ b,
                                                                   const float* __restrict__ w, float* __restrict__ y,
                                                                   int H, int W, int C, int Ho, int Wo, int pad_t,
                                                                   int pad_l, int tiles_x, int nB, long long zs) {
  using G = DwGeom<K, S>;
  extern __shared__ float4 smem4[];
  float4* tile = smem4;
  float4* wsm = smem4 + G::NPIX * QC;
  const int tid = threadIdx.x, q = tid % QC, strip = tid / QC;
  const int ty0 = (blockIdx.x / tiles_x) * G::TOY, tx0 = (blockIdx.x % tiles_x) * G::TOX;
  const int slot = blockIdx.z / nB;
  const int c0 = blockIdx.y * (QC * 4), img = blockIdx.z - slot * nB;
  { const size_t zo = (size_t)slot * zs; x += zo; a = zp(a, zo); b = zp(b, zo); w += zo; y += zo; }
  for (int i = tid; i < K * K * QC; i += G::NT) {
    const int qq = i % QC, tap = i / QC;
    const int src = FLIP ? (K * K - 1 - tap) : tap;
    wsm[i] = (c0 + qq * 4 < C) ? ld4(w + (size_t)src * C + c0 + qq * 4) : f4s(0.f);
  }
  dw_stage_input<K, S>(tile, x, a, b, img, H, W, C, c0, ty0 * S - pad_t, tx0 * S - pad_l);
  __syncthreads();
  if (strip >= G::NSTRIP) return;
  const int oy = strip / G::SPR, ox0 = (strip % G::SPR) * 7;
  float4 acc[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc[j] = f4s(0.f);
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const float4* row = tile + ((size_t)(oy * S + ky) * G::TIX + ox0 * S) * QC + q;
    float4 wv[K];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) wv[kx] = wsm[(ky * K + kx) * QC + q];
#pragma unroll
    for (int j = 0; j < G::NIN; ++j) {     // sliding window: input j feeds output (j - kx) / S for every tap kx
      const float4 v = row[j * QC];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int t = j - kx;
        if (t >= 0 && t % S == 0 && t / S < 7) fma4(acc[t / S], v, wv[kx]);
      }
    }
  }
  const int gy = ty0 + oy;
  if (gy < Ho && c0 + q * 4 < C) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int gx = tx0 + ox0 + j;
      if (gx < Wo) st4(y + (((size_t)img * Ho + gy) * Wo + gx) * C + c0 + q * 4, acc[j]);
    }
  }
}

template <int K, int S, bool FLIP>
static void dw_fwd_launch(const float* x, const float* a, const float* b, const float* w, float* y, int B, int H, int W,
                          int C, int Ho, int Wo, int pad_t, int pad_l, cudaStream_t s) {
  using G = DwGeom<K, S>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(dw_fwd_kernel<K, S, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes());
    attr_done = true;
  }
  const int tiles_x = cdiv(Wo, G::TOX), tiles_y = cdiv(Ho, G::TOY);
  dim3 grid(tiles_x * tiles_y, cdiv(C, QC * 4), B * MLIIS_NZ);
  MLIIS_COUNT(), dw_fwd_kernel<K, S, FLIP><<<grid, G::NT, G::smem_bytes(), s>>>(x, a, b, w, y, H, W, C, Ho, Wo, pad_t, pad_l, tiles_x,
                                                                               B, MLIIS_ZS);
}

void dw_fwd(const float* x, const float* a, const float* b, const float* w, float* y, int B, int H, int W, int C, int k,
            int stride, int Ho, int Wo, int pad_t, int pad_l, cudaStream_t s) {
  if (k == 3 && stride == 1) dw_fwd_launch<3, 1, false>(x, a, b, w, y, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
  else if (k == 3 && stride == 2) dw_fwd_launch<3, 2, false>(x, a, b, w, y, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
  else if (k == 5 && stride == 1) dw_fwd_launch<5, 1, false>(x, a, b, w, y, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
  else dw_fwd_launch<5, 2, false>(x, a, b, w, y, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
}

// ---- stride-2 dgrad: dx[iy,ix] = sum_{ky,kx : parity ok} dy[(iy+pad_t-ky)/2, (ix+pad_l-kx)/2] * w[ky,kx]
template <int K>
struct DwBd2 {
  static constexpr int TO = 14;                       // dx tile edge
  static constexpr int TD = 7 + (K + 1) / 2 + 1;      // staged dy tile edge (10 | 11)
  static constexpr int NT = 224;
  static constexpr size_t smem_bytes() { return (size_t)(TD * TD + K * K) * QC * sizeof(float4); }
};

// One strip = the 7 outputs of one dx row that share the parity of (ix + pad_l): for them the contributing taps are the
// compile-time sets ky = RP, RP+2, .. and kx = XP, XP+2, .., consecutive outputs read consecutive dy columns, and the
// inner loops carry no parity tests.  (The first version walked 7 consecutive outputs per thread and tested the parity
// of every (tap, output) pair at run time: 63 predicated index computations per thread, 26 % of the DRAM peak.)
template <int K, int RP, int XP>
__device__ __forceinline__ void dwbd2_strip(const float4* __restrict__ tile, const float4* __restrict__ wsm, int q, int ly0,
                                            int lx0, float4 (&acc)[7]) {
  constexpr int TD = DwBd2<K>::TD;
  constexpr int NM = (K - RP + 1) / 2, NN = (K - XP + 1) / 2;      // taps of this parity class per column / per row
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const float4* row = tile + ((size_t)(ly0 - m) * TD + (lx0 - (NN - 1))) * QC + q;
    float4 wv[NN];
#pragma unroll
    for (int n = 0; n < NN; ++n) wv[n] = wsm[((RP + 2 * m) * K + XP + 2 * n) * QC + q];
#pragma unroll
    for (int i = 0; i < 7 + NN - 1; ++i) {       // dy column lx0 - (NN-1) + i feeds output j through tap n = j - i + NN - 1
      const float4 v = row[i * QC];
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        const int j = i + n - (NN - 1);
        if (j >= 0 && j < 7) fma4(acc[j], v, wv[n]);
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(224, 4) dw_bwd_data_s2_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                 float* __restrict__ dx, int H, int W, int C, int Ho,
                                                                 int Wo, int pad_t, int pad_l, int tiles_x, int nB,
                                                                 long long zs) {
  using G = DwBd2<K>;
  extern __shared__ float4 smem4[];
  float4* tile = smem4;
  float4* wsm = smem4 + G::TD * G::TD * QC;
  const int tid = threadIdx.x, q = tid % QC, strip = tid / QC;
  const int ty0 = (blockIdx.x / tiles_x) * G::TO, tx0 = (blockIdx.x % tiles_x) * G::TO;
  const int slot = blockIdx.z / nB;
  const int c0 = blockIdx.y * (QC * 4), img = blockIdx.z - slot * nB;
  { const size_t zo = (size_t)slot * zs; dy += zo; w += zo; dx += zo; }
  const bool cvalid = c0 + q * 4 < C;
  for (int i = tid; i < K * K * QC; i += G::NT) {
    const int qq = i % QC, tap = i / QC;
    wsm[i] = (c0 + qq * 4 < C) ? ld4(w + (size_t)tap * C + c0 + qq * 4) : f4s(0.f);
  }
  // first dy row/col that any dx of this tile can touch (floor division, may be negative)
  const int oy_lo = (ty0 + pad_t - (K - 1)) >> 1, ox_lo = (tx0 + pad_l - (K - 1)) >> 1;
  {
    constexpr int TOT = G::TD * G::TD * QC, U = (TOT + G::NT - 1) / G::NT;     // 4: every load in flight before the stores
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = tid + u * G::NT;
      const int pix = i / QC, ly = pix / G::TD, lx = pix - ly * G::TD;
      const int gy = oy_lo + ly, gx = ox_lo + lx;
      v[u] = f4s(0.f);
      if (i < TOT && cvalid && gy >= 0 && gy < Ho && gx >= 0 && gx < Wo)
        v[u] = ld4(dy + (((size_t)img * Ho + gy) * Wo + gx) * C + c0 + q * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (tid + u * G::NT < TOT) tile[tid + u * G::NT] = v[u];
  }
  __syncthreads();
  // 28 strips = 4 parity classes x 7 rows (the tile origin is even: the class of a local row / column is fixed).
  // strips of a class are consecutive, so five of the seven warps run one class and two run two.
  const int cls = strip / 7, k = strip - cls * 7, rp = cls >> 1, xp = cls & 1;
  const int ry = 2 * k + (rp ^ (pad_t & 1)), rx0 = xp ^ (pad_l & 1);       // local row, first local column of the strip
  const int iy = ty0 + ry, ix0 = tx0 + rx0;
  const int ly0 = ((iy + pad_t - rp) >> 1) - oy_lo, lx0 = ((ix0 + pad_l - xp) >> 1) - ox_lo;
  float4 acc[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc[j] = f4s(0.f);
  if (cls == 0) dwbd2_strip<K, 0, 0>(tile, wsm, q, ly0, lx0, acc);
  else if (cls == 1) dwbd2_strip<K, 0, 1>(tile, wsm, q, ly0, lx0, acc);
  else if (cls == 2) dwbd2_strip<K, 1, 0>(tile, wsm, q, ly0, lx0, acc);
  else dwbd2_strip<K, 1, 1>(tile, wsm, q, ly0, lx0, acc);
  if (iy < H && cvalid) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int gx = ix0 + 2 * j;
      if (gx < W) st4(dx + (((size_t)img * H + iy) * W + gx) * C + c0 + q * 4, acc[j]);
    }
  }
}

template <int K>
static void dw_bwd_data_s2_launch(const float* dy, const float* w, float* dx, int B, int H, int W, int C, int Ho, int Wo,
                                  int pad_t, int pad_l, cudaStream_t s) {
  using G = DwBd2<K>;
  const int tiles_x = cdiv(W, G::TO), tiles_y = cdiv(H, G::TO);
  dim3 grid(tiles_x * tiles_y, cdiv(C, QC * 4), B * MLIIS_NZ);
  MLIIS_COUNT(), dw_bwd_data_s2_kernel<K><<<grid, G::NT, G::smem_bytes(), s>>>(dy, w, dx, H, W, C, Ho, Wo, pad_t, pad_l, tiles_x, B,
                                                                              MLIIS_ZS);
}

void dw_bwd_data(const float* dy, const float* w, float* dx, int B, int H, int W, int C, int k, int stride, int Ho,
                 int Wo, int pad_t, int pad_l, cudaStream_t s) {
  if (stride == 1) {
    // dx = correlate(dy, flip(w)) with pad' = k-1-pad ; output size == input size
    if (k == 3) dw_fwd_launch<3, 1, true>(dy, nullptr, nullptr, w, dx, B, Ho, Wo, C, H, W, k - 1 - pad_t, k - 1 - pad_l, s);
    else dw_fwd_launch<5, 1, true>(dy, nullptr, nullptr, w, dx, B, Ho, Wo, C, H, W, k - 1 - pad_t, k - 1 - pad_l, s);
  } else {
    if (k == 3) dw_bwd_data_s2_launch<3>(dy, w, dx, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
    else dw_bwd_data_s2_launch<5>(dy, w, dx, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
  }
}

// ---- wgrad: dW[ky,kx,c] = sum_{b,oy,ox} act(x)[oy*S-pad+ky, ox*S-pad+kx, c] * dy[oy,ox,c]
int dw_wgrad_blocks(int B, int Ho, int Wo, int stride) {
  return B * cdiv(Ho, stride == 1 ? 14 : 7) * cdiv(Wo, 14);
}

template <int K, int S>
__global__ void __launch_bounds__(DwGeom<K, S>::NT, 4) dw_wgrad_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ a,
                                                                     const float* __restrict__ b,
                                                                     const float* __restrict__ dy,
                                                                     float* __restrict__ partials, int H, int W, int C,
                                                                     int Ho, int Wo, int pad_t, int pad_l,
                                                                     int tiles_x, int tiles, int nB, long long zs) {
  using G = DwGeom<K, S>;
  constexpr int NW = G::NT / 32;
  extern __shared__ float4 smem4[];
  float4* tile = smem4;
  const int tid = threadIdx.x, q = tid % QC, strip = tid / QC;
  const int ty0 = (blockIdx.x / tiles_x) * G::TOY, tx0 = (blockIdx.x % tiles_x) * G::TOX;
  const int slot = blockIdx.z / nB;
  const int c0 = blockIdx.y * (QC * 4), img = blockIdx.z - slot * nB;
  { const size_t zo = (size_t)slot * zs; x += zo; a = zp(a, zo); b = zp(b, zo); dy += zo; partials += zo; }
  const bool cvalid = c0 + q * 4 < C;
  dw_stage_input<K, S>(tile, x, a, b, img, H, W, C, c0, ty0 * S - pad_t, tx0 * S - pad_l);
  __syncthreads();
  float4* red = smem4 + G::NPIX * QC;  // [K*K][NW][QC]
  const int warp = tid >> 5, lane = tid & 31;
  const bool active = strip < G::NSTRIP;     // S == 2: the upper half of the block only stages and shuffles zeros
  const int oy = active ? strip / G::SPR : 0, ox0 = active ? (strip % G::SPR) * 7 : 0;
  float4 g[7];
  {
    const int gy = ty0 + oy;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int gx = tx0 + ox0 + j;
      g[j] = (active && cvalid && gy < Ho && gx < Wo) ? ld4(dy + (((size_t)img * Ho + gy) * Wo + gx) * C + c0 + q * 4)
                                                      : f4s(0.f);
    }
  }
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const float4* row = tile + ((size_t)(oy * S + ky) * G::TIX + ox0 * S) * QC + q;
    float4 wrow[K];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) wrow[kx] = f4s(0.f);
#pragma unroll
    for (int j = 0; j < G::NIN; ++j) {
      const float4 v = row[j * QC];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int t = j - kx;
        if (t >= 0 && t % S == 0 && t / S < 7) fma4(wrow[kx], v, g[t / S]);
      }
    }
    // reduce across the 4 strips of a warp (lane = strip_local*8 + q); across warps via smem below
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      float4 v = wrow[kx];
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
      }
      if (lane < QC) red[((size_t)(ky * K + kx) * NW + warp) * QC + lane] = v;
    }
  }
  __syncthreads();
  for (int i = tid; i < K * K * QC; i += G::NT) {
    const int t = i / QC, qq = i - t * QC;
    if (c0 + qq * 4 >= C) continue;
    float4 s4 = red[((size_t)t * NW) * QC + qq];
    for (int wv = 1; wv < NW; ++wv) s4 = s4 + red[((size_t)t * NW + wv) * QC + qq];
    const size_t blk = (size_t)img * tiles + blockIdx.x;
    st4(partials + (blk * (K * K) + t) * C + c0 + qq * 4, s4);
  }
}

template <int K, int S>
static void dw_wgrad_launch(const float* x, const float* a, const float* b, const float* dy, float* partials, float* dw,
                            int B, int H, int W, int C, int Ho, int Wo, int pad_t, int pad_l, cudaStream_t s) {
  using G = DwGeom<K, S>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(dw_wgrad_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::wgrad_smem_bytes());
    attr_done = true;
  }
  const int tiles_x = cdiv(Wo, G::TOX), tiles_y = cdiv(Ho, G::TOY), tiles = tiles_x * tiles_y;
  dim3 grid(tiles, cdiv(C, QC * 4), B * MLIIS_NZ);
  MLIIS_COUNT(), dw_wgrad_kernel<K, S><<<grid, G::NT, G::wgrad_smem_bytes(), s>>>(x, a, b, dy, partials, H, W, C, Ho, Wo, pad_t, pad_l,
                                                             tiles_x, tiles, B, MLIIS_ZS);
  reduce_partials(partials, B * tiles, K * K * C, dw, s);
}

void dw_bwd_weight(const float* x, const float* a, const float* b, const float* dy, float* partials, float* dw, int B,
                   int H, int W, int C, int k, int stride, int Ho, int Wo, int pad_t, int pad_l, cudaStream_t s) {
  if (k == 3 && stride == 1) dw_wgrad_launch<3, 1>(x, a, b, dy, partials, dw, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
  else if (k == 3 && stride == 2) dw_wgrad_launch<3, 2>(x, a, b, dy, partials, dw, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
  else if (k == 5 && stride == 1) dw_wgrad_launch<5, 1>(x, a, b, dy, partials, dw, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
  else dw_wgrad_launch<5, 2>(x, a, b, dy, partials, dw, B, H, W, C, Ho, Wo, pad_t, pad_l, s);
}

}  // namespace mliis
