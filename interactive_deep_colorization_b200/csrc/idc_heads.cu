// Bandwidth-bound kernels either side of the conv trunk:
//   conv1_1_kernel    input pack (cat(L/100, ab/110, mask-maskcent), model.py:142-148) fused with
//                     model1.0 (4->64 conv3x3 + ReLU, model.py:13-14)
//   out_head_kernel   model_out 1x1 128->2 + tanh * 110 (model.py:108-109,175)  [unfused variant]
//   softmax529_kernel softmax(0.2 * logits) over the 529 ab bins (model.py:131,160), NHWC->NCHW
//   lab2rgb_kernel    lab2rgb_transpose (data/colorize_image.py:20-28): Lab -> sRGB uint8
//   global_mlp_kernel global-hints branch (models/global_model/deploy_nodist.prototxt:38-172)
//   act<->NCHW        test hooks
#include "idc_internal.h"

namespace idc {

// ------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// ------------------------------------------------------------------------------------------
// conv1_1: one thread per pixel, 64 output channels in registers.  The 36x64 weights travel as a
// __grid_constant__ kernel parameter, so every FFMA takes its weight straight from the constant
// bank (warp-uniform operand, no LDS/LDG in the inner loop): the first version read them from
// shared memory and was LSU-bound at 1.6 ms per 64-image batch (round-1 profile).
// ------------------------------------------------------------------------------------------
// a / d for finite, normal-range a: reciprocal multiply + one FMA residual correction (the compiler's
// own division sequence without its special-case slow path; inputs are bounded Lab values)
__device__ __forceinline__ float div_corrected(float a, float d, float rd) {
  const float q = a * rd;
  return fmaf(fmaf(-q, d, a), rd, q);
}

template <bool SPLIT>
__global__ void __launch_bounds__(128, 4) conv1_1_kernel(const __grid_constant__ Conv11Weights W,
                                                      const float* __restrict__ L, const float* __restrict__ ab,
                                                      const float* __restrict__ mask, float maskcent, int N, int H,
                                                      int Wd, float* __restrict__ outf, __half* __restrict__ ohi,
                                                      __half* __restrict__ olo) {
  pdl_prologue_done();
  const size_t HW = (size_t)H * Wd;
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pix < (size_t)N * HW;                   // N*HW is a multiple of 64, blocks are 128 wide
  const size_t pixc = live ? pix : 0;
  const int n = (int)(pixc / HW);
  const int r = (int)(pixc - (size_t)n * HW);
  const int y = r / Wd, x = r - y * Wd;
  float in[36];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = y + ky - 1, ix = x + kx - 1;
      const bool ok = live && iy >= 0 && iy < H && ix >= 0 && ix < Wd;
      const size_t o = (size_t)iy * Wd + ix;
      const int t = (ky * 3 + kx) * 4;
      // zero padding applies to the concatenated, normalised input (model.py:148 then Conv2d pad)
      in[t + 0] = ok ? div_corrected(__ldg(L + (size_t)n * HW + o), 100.0f, 0.01f) : 0.f;
      in[t + 1] = ok ? div_corrected(__ldg(ab + (size_t)n * 2 * HW + o), 110.0f, 1.0f / 110.0f) : 0.f;
      in[t + 2] = ok ? div_corrected(__ldg(ab + (size_t)n * 2 * HW + HW + o), 110.0f, 1.0f / 110.0f) : 0.f;
      in[t + 3] = ok ? (__ldg(mask + (size_t)n * HW + o) - maskcent) : 0.f;
    }
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = W.b[c];
#pragma unroll
  for (int k = 0; k < 36; ++k) {
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = fmaf(in[k], W.w[k * 64 + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = fmaxf(acc[c], 0.f);
  if (!SPLIT) {
    if (!live) return;
    float4* op = reinterpret_cast<float4*>(outf + pix * 64);
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) op[c4] = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
  } else {
    // The warp's 32 pixels x 64 channels are one contiguous 4 KB block per plane in NHWC.  Writing 16 bytes per
    // lane at a 128-byte stride costs one L1 transaction per lane; instead transpose through a swizzled smem
    // tile so every store instruction writes 512 contiguous bytes.
    __shared__ __align__(16) uint4 tile[4][32 * 8];          // per warp: 32 rows x 8 chunks of 16 B
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wpix0 = pix - lane;                          // first pixel of this warp (blocks are 128-aligned)
    const size_t total = (size_t)N * HW;
#pragma unroll
    for (int plane = 0; plane < 2; ++plane) {
      if (plane == 1 && !olo) break;                          // IDC_FLAG_FAST_FP16: no lo plane
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        __align__(16) __half h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __half hi, lo;
          split_half(acc[c8 * 8 + j] * kActScale, hi, lo);
          h[j] = plane == 0 ? hi : lo;
        }
        tile[warp][lane * 8 + (c8 ^ (lane & 7))] = *reinterpret_cast<uint4*>(h);
      }
      __syncwarp();
      __half* gbase = (plane == 0 ? ohi : olo) + wpix0 * 64;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = i * 4 + (lane >> 3), c = lane & 7;
        if (wpix0 + rl < total)
          reinterpret_cast<uint4*>(gbase)[i * 32 + lane] = tile[warp][rl * 8 + (c ^ (rl & 7))];
      }
      __syncwarp();
    }
  }
}

cudaError_t launch_conv1_1(Ctx* c, int n, const float* L, const float* ab, const float* mask, float maskcent,
                           cudaStream_t st, int img0) {
  const ActBuf& o = c->bufs[c->buf_index.at("a1_1")];
  const size_t HW = (size_t)o.H * o.W, npix = (size_t)n * HW, ooff = (size_t)img0 * HW * o.C;
  const int grid = (int)((npix + 127) / 128);
  L += img0 * HW; ab += img0 * 2 * HW; mask += img0 * HW;
  cudaError_t e;
  if (c->simt)
    e = launch_k(c, conv1_1_kernel<false>, dim3(grid), dim3(128), 0, st, c->h_w11, L, ab, mask, maskcent, n, o.H, o.W,
                 static_cast<float*>(o.p0) + ooff, (__half*)nullptr, (__half*)nullptr);
  else
    e = launch_k(c, conv1_1_kernel<true>, dim3(grid), dim3(128), 0, st, c->h_w11, L, ab, mask, maskcent, n, o.H, o.W,
                 (float*)nullptr, static_cast<__half*>(o.p0) + ooff,
                 o.p1 ? static_cast<__half*>(o.p1) + ooff : (__half*)nullptr);   // FAST_FP16: no lo plane
  c->launch_count++;
  return e;
}

// ------------------------------------------------------------------------------------------
// unfused regression head: 8 lanes per pixel, 16 channels each
// ------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) out_head_kernel(const float* __restrict__ inf, const __half* __restrict__ ihi,
                                                       const __half* __restrict__ ilo, const float* __restrict__ w,
                                                       const float* __restrict__ b, int N, int H, int W,
                                                       float* __restrict__ out, float out_scale) {
  __shared__ float ws[256];
  ws[threadIdx.x] = w[threadIdx.x];                     // static weights: before the dependency wait
  pdl_prologue_done();
  __syncthreads();
  const size_t HW = (size_t)H * W;
  const size_t pix = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int part = threadIdx.x & 7;
  const bool valid = pix < (size_t)N * HW;
  float s0 = 0.f, s1 = 0.f;
  if (valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int ch = part * 16 + j;
      float v;
      if (SPLIT) v = (__half2float(ihi[pix * 128 + ch]) + (ilo ? __half2float(ilo[pix * 128 + ch]) : 0.f)) * kActInvScale;
      else v = inf[pix * 128 + ch];
      s0 = fmaf(v, ws[ch], s0);
      s1 = fmaf(v, ws[128 + ch], s1);
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if (valid && part == 0) {
    const int n = (int)(pix / HW);
    const size_t r = pix - (size_t)n * HW;
    out[(size_t)n * 2 * HW + r] = tanhf(s0 + b[0]) * out_scale;
    out[(size_t)n * 2 * HW + HW + r] = tanhf(s1 + b[1]) * out_scale;
  }
}

cudaError_t launch_out_head(Ctx* c, int n, float* out_ab, cudaStream_t st) {
  const ActBuf& in = c->bufs[c->buf_index.at("conv10_2")];
  const size_t npix = (size_t)n * in.H * in.W;
  const int grid = (int)((npix * 8 + 255) / 256);
  cudaError_t e;
  if (c->simt)
    e = launch_k(c, out_head_kernel<false>, dim3(grid), dim3(256), 0, st, static_cast<const float*>(in.p0),
                 (const __half*)nullptr, (const __half*)nullptr, c->wout, c->bout, n, in.H, in.W, out_ab,
                 (float)c->opt.tanh_scale);
  else
    e = launch_k(c, out_head_kernel<true>, dim3(grid), dim3(256), 0, st, (const float*)nullptr,
                 static_cast<const __half*>(in.p0), static_cast<const __half*>(in.p1), c->wout, c->bout, n, in.H, in.W,
                 out_ab, (float)c->opt.tanh_scale);
  c->launch_count++;
  return e;
}

// ------------------------------------------------------------------------------------------
// softmax over 529 bins: block = 32 pixels; warp-reduce per pixel, transpose through smem so the
// NCHW store is 128-byte coalesced.
// ------------------------------------------------------------------------------------------
constexpr int kBins = 529;
// one warp, one pixel: v[j] = softmax(0.2 * logits)[lane + 32 j].  Shared by the full-map kernel and the click's
// single-pixel kernel so that both produce the same bits (same instruction sequence, same reduction order).
__device__ __forceinline__ void softmax529_row(const float* __restrict__ row, int lane, float (&v)[17]) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 17; ++j) {
    const int ch = lane + 32 * j;
    v[j] = ch < kBins ? row[ch] * 0.2f : -INFINITY;   // model.py:160 "* .2"
    mx = fmaxf(mx, v[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 17; ++j) {
    v[j] = (lane + 32 * j) < kBins ? expf(v[j] - mx) : 0.f;
    sum += v[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < 17; ++j) v[j] *= inv;
}

__global__ void __launch_bounds__(256) softmax529_kernel(const float* __restrict__ logits, int ld, int M, int HW4,
                                                         float* __restrict__ out) {
  extern __shared__ float tile[];  // [529][33]
  pdl_prologue_done();
  const int p0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = 0; q < 4; ++q) {
    const int pl = warp * 4 + q;
    const int p = p0 + pl;
    if (p >= M) continue;
    float v[17];
    softmax529_row(logits + (size_t)p * ld, lane, v);
#pragma unroll
    for (int j = 0; j < 17; ++j) {
      const int ch = lane + 32 * j;
      if (ch < kBins) tile[ch * 33 + pl] = v[j];
    }
  }
  __syncthreads();
  // store: lanes run over the 32 pixels, warps over channels
  const int p = p0 + lane;
  if (p < M) {
    const int n = p / HW4;
    const int r = p - n * HW4;
    float* ob = out + (size_t)n * kBins * HW4 + r;
    for (int ch = warp; ch < kBins; ch += 8) ob[(size_t)ch * HW4] = tile[ch * 33 + lane];
  }
}

cudaError_t launch_softmax529(Ctx* c, int n, float* out_dist, cudaStream_t st) {
  const int HW4 = (c->H / 4) * (c->W / 4);
  const int M = n * HW4;
  int ld = 0;
  for (auto& op : c->ops)
    if (op.kind == OP_CLASS) ld = op.cout_pad;
  const size_t smem = (size_t)kBins * 33 * sizeof(float);
  static unsigned long long attr_devs = 0;       // the opt-in is per device: one bit per device ordinal
  if (c->dev >= 64 || !(attr_devs & (1ull << c->dev))) {
    cudaError_t e = cudaFuncSetAttribute(softmax529_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (c->dev < 64) attr_devs |= 1ull << c->dev;
  }
  cudaError_t e = launch_k(c, softmax529_kernel, dim3(ceil_div(M, 32)), dim3(256), smem, st, c->logits, ld, M, HW4, out_dist);
  c->launch_count++;
  return e;
}

// ------------------------------------------------------------------------------------------
// Lab -> sRGB uint8, float64 math like the reference's numpy/skimage path.
// skimage 0.13 color.lab2rgb (lab2xyz + xyz2rgb), clip, *255, truncating cast
// (data/colorize_image.py:27).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double srgb_inv_gamma(double c) {
  return c > 0.04045 ? pow((c + 0.055) / 1.055, 2.4) : c / 12.92;
}
__device__ __forceinline__ double lab_f(double t) { return t > 0.008856 ? cbrt(t) : 7.787 * t + 16.0 / 116.0; }

__device__ __forceinline__ double lab_finv(double t) {
  return t > 0.2068966 ? t * t * t : (t - 16.0 / 116.0) / 7.787;
}
__device__ __forceinline__ double srgb_gamma(double c) {
  return c > 0.0031308 ? 1.055 * pow(c, 1.0 / 2.4) - 0.055 : 12.92 * c;
}
// abq (optional): the reference's quantised `output_ab` = rgb2lab(uint8 RGB)[1:] (data/colorize_image.py:196-198,
// row a11) computed from the just-quantised pixel in the same thread, [N,2,HW] float64.
__global__ void lab2rgb_kernel(const float* __restrict__ L, float l_offset, const float* __restrict__ ab, int N,
                               int HW, uint8_t* __restrict__ rgb, double* __restrict__ abq) {
  pdl_prologue_done();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * HW) return;
  const int n = (int)(i / HW);
  const size_t r = i - (size_t)n * HW;
  const double l = (double)L[i] + (double)l_offset;
  const double a = (double)ab[(size_t)n * 2 * HW + r];
  const double b = (double)ab[(size_t)n * 2 * HW + HW + r];
  const double fy = (l + 16.0) / 116.0;
  const double fx = a / 500.0 + fy;
  double fz = fy - b / 200.0;
  if (fz < 0.0) fz = 0.0;
  const double X = lab_finv(fx) * 0.95047, Y = lab_finv(fy) * 1.0, Z = lab_finv(fz) * 1.08883;
  // inverse of the sRGB->XYZ matrix used by skimage (xyz_from_rgb), float64
  const double m00 = 3.240481343200526, m01 = -1.5371515162713185, m02 = -0.4985363261688878;
  const double m10 = -0.9692549499965682, m11 = 1.8759900014898907, m12 = 0.04155592655829284;
  const double m20 = 0.05564663913517716, m21 = -0.20404133836651123, m22 = 1.0573110696453443;
  double R = m00 * X + m01 * Y + m02 * Z;
  double G = m10 * X + m11 * Y + m12 * Z;
  double B = m20 * X + m21 * Y + m22 * Z;
  R = srgb_gamma(R); G = srgb_gamma(G); B = srgb_gamma(B);
  R = fmin(fmax(R, 0.0), 1.0) * 255.0;
  G = fmin(fmax(G, 0.0), 1.0) * 255.0;
  B = fmin(fmax(B, 0.0), 1.0) * 255.0;
  const uint8_t r8 = (uint8_t)R, g8 = (uint8_t)G, b8 = (uint8_t)B;
  rgb[i * 3 + 0] = r8;
  rgb[i * 3 + 1] = g8;
  rgb[i * 3 + 2] = b8;
  if (abq) {   // same arithmetic as rgb2lab_kernel below
    const double rl = srgb_inv_gamma(r8 / 255.0), gl = srgb_inv_gamma(g8 / 255.0), bl = srgb_inv_gamma(b8 / 255.0);
    const double fx2 = lab_f((0.412453 * rl + 0.357580 * gl + 0.180423 * bl) / 0.95047);
    const double fy2 = lab_f((0.212671 * rl + 0.715160 * gl + 0.072169 * bl) / 1.0);
    const double fz2 = lab_f((0.019334 * rl + 0.119193 * gl + 0.950227 * bl) / 1.08883);
    abq[(size_t)n * 2 * HW + r] = 500.0 * (fx2 - fy2);
    abq[(size_t)n * 2 * HW + HW + r] = 200.0 * (fy2 - fz2);
  }
}

// ------------------------------------------------------------------------------------------
// f1: the steps either side of the network on the GPU, float64 like numpy/skimage/scipy.
//   rgb2lab_kernel       uint8 RGB -> Lab planes (skimage rgb2lab; data/colorize_image.py:31-36,172-178,196-198)
//   zoom_lab2rgb_kernel  scipy.ndimage.zoom(order=1) of the ab planes to the full-resolution grid +
//                        lab2rgb_transpose with the full-resolution L (get_img_fullres, :123-131)
// ------------------------------------------------------------------------------------------
__global__ void rgb2lab_kernel(const uint8_t* __restrict__ rgb, int N, int HW, double* __restrict__ lab) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * HW) return;
  const int n = (int)(i / HW);
  const size_t r = i - (size_t)n * HW;
  const double R = srgb_inv_gamma(rgb[i * 3 + 0] / 255.0), G = srgb_inv_gamma(rgb[i * 3 + 1] / 255.0),
               B = srgb_inv_gamma(rgb[i * 3 + 2] / 255.0);
  const double X = (0.412453 * R + 0.357580 * G + 0.180423 * B) / 0.95047;
  const double Y = (0.212671 * R + 0.715160 * G + 0.072169 * B) / 1.0;
  const double Z = (0.019334 * R + 0.119193 * G + 0.950227 * B) / 1.08883;
  const double fx = lab_f(X), fy = lab_f(Y), fz = lab_f(Z);
  double* o = lab + (size_t)n * 3 * HW + r;
  o[0] = 116.0 * fy - 16.0;
  o[HW] = 500.0 * (fx - fy);
  o[2 * (size_t)HW] = 200.0 * (fy - fz);
}

__device__ __forceinline__ void lab_to_rgb_u8(double l, double a, double b, uint8_t* out) {
  const double fy = (l + 16.0) / 116.0;
  const double fx = a / 500.0 + fy;
  double fz = fy - b / 200.0;
  if (fz < 0.0) fz = 0.0;
  const double X = lab_finv(fx) * 0.95047, Y = lab_finv(fy) * 1.0, Z = lab_finv(fz) * 1.08883;
  double R = 3.240481343200526 * X + -1.5371515162713185 * Y + -0.4985363261688878 * Z;
  double G = -0.9692549499965682 * X + 1.8759900014898907 * Y + 0.04155592655829284 * Z;
  double B = 0.05564663913517716 * X + -0.20404133836651123 * Y + 1.0573110696453443 * Z;
  R = srgb_gamma(R); G = srgb_gamma(G); B = srgb_gamma(B);
  out[0] = (uint8_t)(fmin(fmax(R, 0.0), 1.0) * 255.0);
  out[1] = (uint8_t)(fmin(fmax(G, 0.0), 1.0) * 255.0);
  out[2] = (uint8_t)(fmin(fmax(B, 0.0), 1.0) * 255.0);
}

// scipy.ndimage.zoom(order=1, grid_mode=False): output o samples input coordinate o * (in - 1) / (out - 1)
__global__ void zoom_lab2rgb_kernel(const double* __restrict__ ab, int hin, int win, const double* __restrict__ Lfull,
                                    int H, int W, uint8_t* __restrict__ rgb) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)H * W) return;
  const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
  const double cy = H > 1 ? (double)y * (double)(hin - 1) / (double)(H - 1) : 0.0;
  const double cx = W > 1 ? (double)x * (double)(win - 1) / (double)(W - 1) : 0.0;
  int y0 = (int)floor(cy), x0 = (int)floor(cx);
  if (y0 > hin - 2) y0 = hin - 2 < 0 ? 0 : hin - 2;
  if (x0 > win - 2) x0 = win - 2 < 0 ? 0 : win - 2;
  const double ty = cy - y0, tx = cx - x0;
  const int y1 = y0 + 1 < hin ? y0 + 1 : y0, x1 = x0 + 1 < win ? x0 + 1 : x0;
  double v[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const double* p = ab + (size_t)c * hin * win;
    const double top = (1.0 - tx) * p[(size_t)y0 * win + x0] + tx * p[(size_t)y0 * win + x1];
    const double bot = (1.0 - tx) * p[(size_t)y1 * win + x0] + tx * p[(size_t)y1 * win + x1];
    v[c] = (1.0 - ty) * top + ty * bot;
  }
  lab_to_rgb_u8(Lfull[i], v[0], v[1], rgb + i * 3);
}

// ------------------------------------------------------------------------------------------
// f3: global statistics of a reference image = the glob vector of BASELINE config 4
// (models/global_model/global_stats.prototxt: BGR2Lab -> 4x4 average pool of ab -> NNEncLayer with NN=1,
//  sigma=5 (caffe_files/caffe_traininglayers.py:161-196: hard assignment to the nearest of the 313 bins) ->
//  global average = histogram; BGR2HSV -> global average of S).  One thread per pooled 4x4 cell.
//  out[316] = [313 histogram, 1, mean saturation, 1]   (indicators as data/colorize_image.py:452-463 sets them)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) global_stats_kernel(const uint8_t* __restrict__ rgb, int H, int W,
                                                           const float* __restrict__ pts, float* __restrict__ out) {
  __shared__ int hist[313];
  __shared__ double ssum[8];
  for (int i = threadIdx.x; i < 313; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int H4 = H / 4, W4 = W / 4;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  double sat = 0.0;
  if (cell < H4 * W4) {
    const int cy = cell / W4, cx = cell - cy * W4;
    double sa = 0.0, sb = 0.0;
    for (int dy = 0; dy < 4; ++dy)
      for (int dx = 0; dx < 4; ++dx) {
        const uint8_t* px = rgb + ((size_t)(cy * 4 + dy) * W + (cx * 4 + dx)) * 3;
        const double r8 = px[0] / 255.0, g8 = px[1] / 255.0, b8 = px[2] / 255.0;
        const double R = srgb_inv_gamma(r8), G = srgb_inv_gamma(g8), B = srgb_inv_gamma(b8);
        const double X = (0.412453 * R + 0.357580 * G + 0.180423 * B) / 0.95047;
        const double Y = (0.212671 * R + 0.715160 * G + 0.072169 * B);
        const double Z = (0.019334 * R + 0.119193 * G + 0.950227 * B) / 1.08883;
        const double fx = lab_f(X), fy = lab_f(Y), fz = lab_f(Z);
        sa += 500.0 * (fx - fy);
        sb += 200.0 * (fy - fz);
        const double mx = fmax(r8, fmax(g8, b8)), mn = fmin(r8, fmin(g8, b8));
        sat += mx > 0.0 ? (mx - mn) / mx : 0.0;            // skimage rgb2hsv saturation
      }
    const float a = (float)(sa / 16.0), b = (float)(sb / 16.0);
    int best = 0;
    float bd = 3.4e38f;
    for (int k = 0; k < 313; ++k) {
      const float da = a - pts[2 * k], db = b - pts[2 * k + 1];
      const float d = da * da + db * db;
      if (d < bd) { bd = d; best = k; }
    }
    atomicAdd(&hist[best], 1);
  }
  // block reduction of the saturation sum (fixed order inside the block)
  for (int o = 16; o > 0; o >>= 1) sat += __shfl_xor_sync(0xffffffffu, sat, o);
  if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5] = sat;
  __syncthreads();
  const float inv_cells = 1.0f / (float)(H4 * W4);
  for (int i = threadIdx.x; i < 313; i += blockDim.x)
    if (hist[i]) atomicAdd(out + i, hist[i] * inv_cells);
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ssum[i];
    atomicAdd(out + 314, (float)(t / ((double)H * W)));
  }
}

cudaError_t launch_global_stats(int h, int w, const uint8_t* rgb, const float* pts, float* out316, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out316, 0, 316 * sizeof(float), st);
  if (e != cudaSuccess) return e;
  const int cells = (h / 4) * (w / 4);
  global_stats_kernel<<<(cells + 255) / 256, 256, 0, st>>>(rgb, h, w, pts, out316);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const float one = 1.0f;
  e = cudaMemcpyAsync(out316 + 313, &one, sizeof(float), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  return cudaMemcpyAsync(out316 + 315, &one, sizeof(float), cudaMemcpyHostToDevice, st);
}

cudaError_t launch_rgb2lab(int n, int h, int w, const uint8_t* rgb, double* lab, cudaStream_t st) {
  const size_t tot = (size_t)n * h * w;
  rgb2lab_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(rgb, n, h * w, lab);
  return cudaGetLastError();
}

cudaError_t launch_zoom_lab2rgb(const double* ab, int hin, int win, const double* Lfull, int H, int W, uint8_t* rgb,
                                cudaStream_t st) {
  const size_t tot = (size_t)H * W;
  zoom_lab2rgb_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(ab, hin, win, Lfull, H, W, rgb);
  return cudaGetLastError();
}

cudaError_t launch_lab2rgb(Ctx* c, int n, int h, int w, const float* L, float l_offset, const float* ab, uint8_t* rgb,
                           cudaStream_t st, double* abq) {
  const size_t tot = (size_t)n * h * w;
  return launch_k(c, lab2rgb_kernel, dim3((unsigned)((tot + 127) / 128)), dim3(128), 0, st, L, l_offset, ab, n, h * w, rgb, abq);
}

// ------------------------------------------------------------------------------------------
// Caffe-spec 313-bin decode (deploy_nopred.prototxt:776-850): the two grouped x2 "bilinear" deconvolutions
// (kernel outer([.5,1,.5,0]), stride 2, pad 1) compose to a x4 upsample whose output 4i+r is
//   w0[r]*a[i] + w1[r]*a[i+1],  w0 = {1,.75,.5,.25}, w1 = {0,.25,.5,.75},  a[len] = 0 (zero padding),
// separably in y and x.  One warp per source cell (i, j) = 16 output pixels; lanes run over the 313 bins.
// ------------------------------------------------------------------------------------------
constexpr int kBins313 = 313;
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void load_cell313(const float* __restrict__ logits, int ld, int n, int H4, int W4, int i, int j,
                                             int lane, float (&a)[4][10]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ii = i + (k >> 1), jj = j + (k & 1);
    const bool ok = ii < H4 && jj < W4;
    const float* row = logits + ((size_t)(n * H4 + (ok ? ii : 0)) * W4 + (ok ? jj : 0)) * ld;
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      const int b = lane + 32 * q;
      a[k][q] = (ok && b < kBins313) ? __ldg(row + b) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) decode313_kernel(const float* __restrict__ logits, int ld, int N, int H4, int W4,
                                                        const float* __restrict__ pts, float T, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x * 8 + warp;
  if (cell >= N * H4 * W4) return;
  const int n = cell / (H4 * W4);
  const int r = cell - n * H4 * W4;
  const int i = r / W4, j = r - i * W4;
  float a[4][10];
  load_cell313(logits, ld, n, H4, W4, i, j, lane, a);
  float pa[10], pb[10];
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    const int b = lane + 32 * q;
    pa[q] = b < kBins313 ? pts[2 * b] : 0.f;
    pb[q] = b < kBins313 ? pts[2 * b + 1] : 0.f;
  }
  const int H = H4 * 4, W = W4 * 4;
  const float w0[4] = {1.f, .75f, .5f, .25f}, w1[4] = {0.f, .25f, .5f, .75f};
#pragma unroll
  for (int ry = 0; ry < 4; ++ry)
#pragma unroll
    for (int rx = 0; rx < 4; ++rx) {
      float v[10], mx = -INFINITY;
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const float top = w0[rx] * a[0][q] + w1[rx] * a[1][q];
        const float bot = w0[rx] * a[2][q] + w1[rx] * a[3][q];
        v[q] = (lane + 32 * q) < kBins313 ? T * (w0[ry] * top + w1[ry] * bot) : -INFINITY;
        mx = fmaxf(mx, v[q]);
      }
      mx = warp_max(mx);
      float s = 0.f, sa = 0.f, sb = 0.f;
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const float e = (lane + 32 * q) < kBins313 ? expf(v[q] - mx) : 0.f;
        s += e; sa = fmaf(e, pa[q], sa); sb = fmaf(e, pb[q], sb);
      }
      s = warp_sum(s); sa = warp_sum(sa); sb = warp_sum(sb);
      if (lane == 0) {
        const size_t o = (size_t)n * 2 * H * W + (size_t)(4 * i + ry) * W + (4 * j + rx);
        out[o] = sa / s;
        out[o + (size_t)H * W] = sb / s;
      }
    }
}

__global__ void dist313_pixel_kernel(const float* __restrict__ logits, int ld, int H4, int W4, int img, int y, int x,
                                     float S, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int i = y >> 2, j = x >> 2, ry = y & 3, rx = x & 3;
  float a[4][10];
  load_cell313(logits, ld, img, H4, W4, i, j, lane, a);
  const float w0[4] = {1.f, .75f, .5f, .25f}, w1[4] = {0.f, .25f, .5f, .75f};
  float v[10], mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    const float top = w0[rx] * a[0][q] + w1[rx] * a[1][q];
    const float bot = w0[rx] * a[2][q] + w1[rx] * a[3][q];
    v[q] = (lane + 32 * q) < kBins313 ? S * (w0[ry] * top + w1[ry] * bot) : -INFINITY;
    mx = fmaxf(mx, v[q]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    v[q] = (lane + 32 * q) < kBins313 ? expf(v[q] - mx) : 0.f;
    s += v[q];
  }
  s = warp_sum(s);
#pragma unroll
  for (int q = 0; q < 10; ++q)
    if ((lane + 32 * q) < kBins313) out[lane + 32 * q] = v[q] / s;
}

cudaError_t launch_decode313(Ctx* c, int n, float T, float* out_ab, cudaStream_t st) {
  const int H4 = c->H / 4, W4 = c->W / 4;
  const int cells = n * H4 * W4;
  decode313_kernel<<<ceil_div(cells, 8), 256, 0, st>>>(c->logits313, 320, n, H4, W4, c->pts313, T, out_ab);
  c->launch_count++;
  return cudaGetLastError();
}

cudaError_t launch_dist313_pixel(Ctx* c, int img, int y, int x, float S, float* out313_dev, cudaStream_t st) {
  dist313_pixel_kernel<<<1, 32, 0, st>>>(c->logits313, 320, c->H / 4, c->W / 4, img, y, x, S, out313_dev);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// global-hints MLP: 4 x (1x1 conv = dense layer, ReLU, BN).  One warp per output neuron.
// Layer 0 consumes [hist313, ind] (glob_conv1_1) and [s_avg, ind] (glob_s_conv1_1) summed.
// ------------------------------------------------------------------------------------------
__global__ void dense_relu_bn_kernel(const float* __restrict__ x, int xin, int xld, const float* __restrict__ w,
                                     const float* __restrict__ b, const float* __restrict__ scale,
                                     const float* __restrict__ shift, int cout, float* __restrict__ y, int yld) {
  pdl_prologue_done();
  const int n = blockIdx.y;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= cout) return;
  float s = 0.f;
  for (int i = lane; i < xin; i += 32) s = fmaf(x[(size_t)n * xld + i], w[(size_t)o * xin + i], s);
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if (lane == 0) {
    float v = fmaxf(s + b[o], 0.f);
    y[(size_t)n * yld + o] = v * scale[o] + shift[o];
  }
}

cudaError_t launch_global_mlp(Ctx* c, int n, const float* glob, cudaStream_t st) {
  // layer 0: input 316 = [313 hist, 1 ind, 1 sat, 1 ind]; weight [512][316] (both branches concatenated)
  const float* x = glob;
  int xin = 316, xld = 316;
  for (int l = 0; l < 4; ++l) {
    float* y = (l == 3) ? c->gvec : c->gtmp + (size_t)(l & 1) * c->max_n * 512;
    dim3 grid(512 / 8, n);
    cudaError_t e = launch_k(c, dense_relu_bn_kernel, grid, dim3(256), 0, st, x, xin, xld, (const float*)c->gw[l],
                             (const float*)c->gb[l], (const float*)c->gscale[l], (const float*)c->gshift[l], 512, y, 512);
    if (e != cudaSuccess) return e;
    c->launch_count++;
    x = y; xin = 512; xld = 512;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// colour suggestions (data/colorize_image.py:322-354): the reference draws 25 000 samples
// from one pixel's 529-bin pmf and k-means them.  The N -> infinity limit of that procedure is weighted
// k-means over the 529 gamut points with the pmf as weights; this kernel runs it deterministically in one
// CTA: greedy farthest-point seeding (first seed = heaviest bin, next = argmax w * d^2), Lloyd iterations in
// FP64 until the assignment is stable, clusters ordered by mass.  Warp k owns cluster k.
// ------------------------------------------------------------------------------------------
constexpr int kReccBins = 529, kReccMaxK = 32;

__device__ __forceinline__ int block_argmax(double v, int idx, double* rv, int* ri) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  if (lane == 0) { rv[warp] = v; ri[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    v = rv[lane]; idx = ri[lane];
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, off);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) ri[0] = idx;
  }
  __syncthreads();
  const int r = ri[0];
  __syncthreads();
  return r;
}

// The click of the interactive path (idc_set_click): `click` = {img, y4, x4, K, seq, ...} in mapped host memory, read
// when the graph RUNS (the graph itself never changes).  One warp computes the clicked pixel's softmax straight from the
// class logits -- the same per-row routine as softmax529_kernel, so the 529 floats are bit-identical to
// dist[img, :, y4, x4] without waiting for the full-map softmax -- behind an 8-int header that echoes the click, so the
// host can tell which pixel the block belongs to.
__global__ void __launch_bounds__(32) click_pmf_kernel(const float* __restrict__ logits, int ld, const int* __restrict__ click,
                                                       int n_img, int H4, int W4, int* __restrict__ out_hdr,
                                                       float* __restrict__ out_pmf) {
  const int lane = threadIdx.x;
  const int img = click[0], y4 = click[1], x4 = click[2];
  const bool ok = img >= 0 && img < n_img && y4 >= 0 && y4 < H4 && x4 >= 0 && x4 < W4;
  if (lane < 8) out_hdr[lane] = lane == 7 ? (ok ? 1 : 0) : click[lane];
  if (!ok) return;
  float v[17];
  softmax529_row(logits + ((size_t)(img * H4 + y4) * W4 + x4) * ld, lane, v);
#pragma unroll
  for (int j = 0; j < 17; ++j) {
    const int ch = lane + 32 * j;
    if (ch < kBins) out_pmf[ch] = v[j];
  }
}

cudaError_t launch_click_pmf(Ctx* c, const int* click_dev, int n_img, int* out_hdr, float* out_pmf, cudaStream_t st) {
  int ld = 0;
  for (auto& op : c->ops)
    if (op.kind == OP_CLASS) ld = op.cout_pad;
  click_pmf_kernel<<<1, 32, 0, st>>>(c->logits, ld, click_dev, n_img, c->H / 4, c->W / 4, out_hdr, out_pmf);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(1024) ab_reccs_kernel(const float* __restrict__ pmf, size_t bin_stride,
                                                        const float* __restrict__ pts, int K, int max_iter,
                                                        double* __restrict__ out_all, const int* __restrict__ dyn) {
  // CTA v = restart v: its first seed is the bin of weight-rank v (0 = heaviest); out_all[v] = [K][2]
  // centres, [K] mass, iterations, inertia
  if (dyn) {                       // click graph: K comes from the click header written by click_pmf_kernel
    K = dyn[3];
    if (!dyn[7] || K < 1 || K > 32) return;
  }
  double* out = out_all + (size_t)blockIdx.x * (3 * K + 2);
  __shared__ double w[kReccBins], mind[kReccBins];
  __shared__ double px[kReccBins], py[kReccBins];
  __shared__ int label[kReccBins];
  __shared__ double cx[kReccMaxK], cy[kReccMaxK], mass[kReccMaxK], rv[32];
  __shared__ int ri[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool live = tid < kReccBins;

  // normalised weights (fixed-order tree sum)
  double v = live ? (double)pmf[(size_t)tid * bin_stride] : 0.0;
  double s = v;
#pragma unroll
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) rv[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = rv[lane];
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) rv[0] = s;
  }
  __syncthreads();
  const double total = rv[0];
  __syncthreads();
  if (live) {
    w[tid] = v / total;
    px[tid] = (double)pts[2 * tid];
    py[tid] = (double)pts[2 * tid + 1];
    label[tid] = -1;
  }
  __syncthreads();

  // greedy seeding.  Restart v starts from the bin of weight-rank v (order: weight descending, index ascending):
  // v + 1 arg-max rounds with the winners taken out, instead of ranking all 529 bins against each other
  int first = 0;
  {
    bool taken = false;
    for (int r = 0; r <= (int)blockIdx.x; ++r) {
      first = block_argmax(live && !taken ? w[tid] : -1.0, live ? tid : 0x7fffffff, rv, ri);
      taken = taken || tid == first;
    }
  }
  for (int j = 0; j < K; ++j) {
    double score = -1.0;
    if (live) score = j == 0 ? (tid == first ? 2.0 : -1.0) : w[tid] * mind[tid];
    const int pick = block_argmax(score, live ? tid : 0x7fffffff, rv, ri);
    if (tid == 0) { cx[j] = px[pick]; cy[j] = py[pick]; }
    __syncthreads();
    if (live) {
      const double dx = px[tid] - cx[j], dy = py[tid] - cy[j];
      const double d = dx * dx + dy * dy;
      mind[tid] = j == 0 ? d : fmin(mind[tid], d);
    }
  }
  __syncthreads();

  // Lloyd
  int iters = 0;
  for (; iters < max_iter; ++iters) {
    int changed = 0;
    if (live) {
      int best = 0;
      double bd = INFINITY;
      for (int k = 0; k < K; ++k) {
        const double dx = px[tid] - cx[k], dy = py[tid] - cy[k];
        const double d = dx * dx + dy * dy;
        if (d < bd) { bd = d; best = k; }
      }
      changed = best != label[tid];
      label[tid] = best;
    }
    if (!__syncthreads_or(changed)) break;
    if (warp < K) {
      double sw = 0.0, sx = 0.0, sy = 0.0;
      for (int i = lane; i < kReccBins; i += 32)
        if (label[i] == warp) { sw += w[i]; sx += w[i] * px[i]; sy += w[i] * py[i]; }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        sw += __shfl_xor_sync(0xffffffffu, sw, off);
        sx += __shfl_xor_sync(0xffffffffu, sx, off);
        sy += __shfl_xor_sync(0xffffffffu, sy, off);
      }
      if (lane == 0) {
        mass[warp] = sw;
        if (sw > 0.0) { cx[warp] = sx / sw; cy[warp] = sy / sw; }   // an empty cluster keeps its centre
      }
    }
    __syncthreads();
  }

  // inertia = sum_i w_i * min_k d(i, k) with the final centres (fixed-order tree sum)
  double e = 0.0;
  if (live) {
    double bd = INFINITY;
    for (int k = 0; k < K; ++k) {
      const double dx = px[tid] - cx[k], dy = py[tid] - cy[k];
      bd = fmin(bd, dx * dx + dy * dy);
    }
    e = w[tid] * bd;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
  if (lane == 0) rv[warp] = e;
  __syncthreads();
  if (warp == 0) {
    e = rv[lane];
#pragma unroll
    for (int off = 16; off; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if (lane == 0) out[3 * K + 1] = e;
  }

  // order by mass, descending, stable
  if (tid == 0) {
    int order[kReccMaxK];
    for (int k = 0; k < K; ++k) order[k] = k;
    for (int a = 1; a < K; ++a) {
      const int o = order[a];
      int b = a - 1;
      while (b >= 0 && mass[order[b]] < mass[o]) { order[b + 1] = order[b]; --b; }
      order[b + 1] = o;
    }
    for (int k = 0; k < K; ++k) {
      out[2 * k] = cx[order[k]];
      out[2 * k + 1] = cy[order[k]];
      out[2 * K + k] = mass[order[k]];
    }
    out[3 * K] = (double)iters;
  }
}

cudaError_t launch_ab_reccs(const float* pmf, size_t bin_stride, const float* pts_dev, int K, int max_iter,
                            int n_init, double* out_dev, cudaStream_t st, const int* dyn) {
  ab_reccs_kernel<<<n_init, 1024, 0, st>>>(pmf, bin_stride, pts_dev, K, max_iter, out_dev, dyn);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// test hooks: activation <-> NCHW fp32
// ------------------------------------------------------------------------------------------
__global__ void act_to_nchw_kernel(const float* f, const __half* hi, const __half* lo, int N, int H, int W, int C,
                                   float* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)N * H * W * C;
  if (i >= tot) return;
  const int cch = (int)(i % C);
  const size_t p = i / C;
  const int x = (int)(p % W);
  const int y = (int)((p / W) % H);
  const int n = (int)(p / ((size_t)W * H));
  const float v = f ? f[i] : (__half2float(hi[i]) + (lo ? __half2float(lo[i]) : 0.f)) * kActInvScale;
  out[(((size_t)n * C + cch) * H + y) * W + x] = v;
}
__global__ void nchw_to_act_kernel(const float* in, int N, int H, int W, int C, float* f, __half* hi, __half* lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)N * H * W * C;
  if (i >= tot) return;
  const int cch = (int)(i % C);
  const size_t p = i / C;
  const int x = (int)(p % W);
  const int y = (int)((p / W) % H);
  const int n = (int)(p / ((size_t)W * H));
  const float v = in[(((size_t)n * C + cch) * H + y) * W + x];
  if (f) f[i] = v;
  else {
    __half h, l;
    split_half(v * kActScale, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

cudaError_t launch_act_to_nchw(Ctx* c, const ActBuf& b, int n, float* out, cudaStream_t st) {
  const size_t tot = (size_t)n * b.H * b.W * b.C;
  if (c->simt)
    act_to_nchw_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(static_cast<const float*>(b.p0), nullptr, nullptr, n,
                                                               b.H, b.W, b.C, out);
  else
    act_to_nchw_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(nullptr, static_cast<const __half*>(b.p0),
                                                               static_cast<const __half*>(b.p1), n, b.H, b.W, b.C, out);
  return cudaGetLastError();
}

cudaError_t launch_nchw_to_act(Ctx* c, const ActBuf& b, int n, const float* in, cudaStream_t st) {
  const size_t tot = (size_t)n * b.H * b.W * b.C;
  if (c->simt)
    nchw_to_act_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(in, n, b.H, b.W, b.C, static_cast<float*>(b.p0),
                                                               nullptr, nullptr);
  else
    nchw_to_act_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(in, n, b.H, b.W, b.C, nullptr,
                                                               static_cast<__half*>(b.p0), static_cast<__half*>(b.p1));
  return cudaGetLastError();
}

}  // namespace idc
