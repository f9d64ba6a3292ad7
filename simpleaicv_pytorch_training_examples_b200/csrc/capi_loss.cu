// SAMLoss on full-resolution mask logits (SimpleAICV/interactive_segmentation/losses.py:11-198) as two streaming
// passes.  The reference evaluates sigmoid focal loss, dice loss and the IoU-prediction target with ~25 full-size
// elementwise torch kernels per prompt iteration on [B, M, 1024, 1024] fp32 maps (x 5 iterations per step); here ONE
// pass reduces the six per-(b, m) sums all three terms need, and ONE pass writes the logit gradient:
//   sums[b*M+m] = { S_f = sum focal_weight * bce,  S_pt = sum sigmoid(x) t,  S_p = sum sigmoid(x),  S_t = sum t,
//                   I = #(x > thr & t > thr),  U = #(x > thr | t > thr) }
//   focal[b,m] = S_f / (N B)   dice[b,m] = (1 - (2 S_pt + 1) / (S_p + S_t + 1)) / B   gt_iou = clamp(I / max(U, 1e-6), 0, 1)
//   dL/dx = cf[b,m] * d(focal_weight * bce)/dx + sigmoid'(x) * (c1[b,m] * t + c2[b,m])
// HBM-bound: 8 B/pixel forward (fp32 logits + fp32 target; the target plane is shared by the M masks of an image and
// mostly L2-resident), 12 B/pixel backward.  Reductions are per-block partials folded in a fixed order (bit-reproducible).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/saicv_b200.h"
#include "host_util.h"

namespace saicv {
namespace {

constexpr int kLossThreads = 256;
constexpr int kMaxBlocksPerPlane = 64;

__device__ __forceinline__ float4 load4(const void* p, long long i4, bool bf16) {
  if (!bf16) return __ldg(reinterpret_cast<const float4*>(p) + i4);
  const uint2 w = __ldg(reinterpret_cast<const uint2*>(p) + i4);
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                     __uint_as_float(w.y & 0xffff0000u));
}

struct PixelTerms {
  float focal, sig;
};
// focal_weight * bce and sigmoid of one pixel (losses.py:126-141)
__device__ __forceinline__ PixelTerms pixel_terms(float x, float t, float alpha, float gamma) {
  const float e = __expf(-fabsf(x));
  const float inv = __fdividef(1.f, 1.f + e);
  const float sig = x >= 0.f ? inv : e * inv;
  const float bce = fmaxf(x, 0.f) - x * t + log1pf(e);
  const float pt = sig * t + (1.f - sig) * (1.f - t);
  const float af = alpha * t + (1.f - alpha) * (1.f - t);
  const float om = 1.f - pt;
  const float w = gamma == 2.f ? om * om : powf(om, gamma);
  return {af * w * bce, sig};
}

__global__ void __launch_bounds__(kLossThreads)
sam_loss_sums_kernel(const void* __restrict__ logits, int logits_bf16, const float* __restrict__ targets, float* __restrict__ partials,
                     int M, long long n4, float alpha, float gamma, float thr) {
  const int plane = blockIdx.y;                    // b * M + m
  const int b = plane / M;
  const size_t esz = logits_bf16 ? 2 : 4;
  const char* lp = reinterpret_cast<const char*>(logits) + (size_t)plane * n4 * 4 * esz;
  const float4* tp = reinterpret_cast<const float4*>(targets) + (long long)b * n4;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = load4(lp, i, logits_bf16 != 0);
    const float4 t = __ldg(tp + i);
    const float xs[4] = {x.x, x.y, x.z, x.w}, ts[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const PixelTerms pt = pixel_terms(xs[k], ts[k], alpha, gamma);
      s[0] += pt.focal;
      s[1] += pt.sig * ts[k];
      s[2] += pt.sig;
      s[3] += ts[k];
      const bool px = xs[k] > thr, tx = ts[k] > thr;
      s[4] += (px && tx) ? 1.f : 0.f;
      s[5] += (px || tx) ? 1.f : 0.f;
    }
  }
  __shared__ float red[kLossThreads / 32][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][threadIdx.x];
    partials[((long long)plane * gridDim.x + blockIdx.x) * 6 + threadIdx.x] = v;
  }
}

__global__ void sam_loss_fold_kernel(const float* __restrict__ partials, float* __restrict__ sums, int planes, int nblk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // plane * 6 + k
  if (i >= planes * 6) return;
  const int plane = i / 6, k = i % 6;
  float v = 0.f;
  for (int b = 0; b < nblk; ++b) v += partials[((long long)plane * nblk + b) * 6 + k];
  sums[i] = v;
}

// dlogits = cf * d(focal_weight * bce)/dx + sigmoid'(x) * (c1 * t + c2), coefficients per plane in coef[plane][3]
__global__ void __launch_bounds__(kLossThreads)
sam_loss_bwd_kernel(const void* __restrict__ logits, int logits_bf16, const float* __restrict__ targets, const float* __restrict__ coef,
                    void* __restrict__ dlogits, int dl_bf16, int M, long long n4, float alpha, float gamma) {
  const int plane = blockIdx.y;
  const int b = plane / M;
  const size_t esz = logits_bf16 ? 2 : 4;
  const char* lp = reinterpret_cast<const char*>(logits) + (size_t)plane * n4 * 4 * esz;
  const float4* tp = reinterpret_cast<const float4*>(targets) + (long long)b * n4;
  const float cf = coef[plane * 3], c1 = coef[plane * 3 + 1], c2 = coef[plane * 3 + 2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = load4(lp, i, logits_bf16 != 0);
    const float4 t = __ldg(tp + i);
    const float xs[4] = {x.x, x.y, x.z, x.w}, ts[4] = {t.x, t.y, t.z, t.w};
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xv = xs[k], tv = ts[k];
      const float e = __expf(-fabsf(xv));
      const float inv = __fdividef(1.f, 1.f + e);
      const float sig = xv >= 0.f ? inv : e * inv;
      const float ds = sig * (1.f - sig);
      const float bce = fmaxf(xv, 0.f) - xv * tv + log1pf(e);
      const float pt = sig * tv + (1.f - sig) * (1.f - tv);
      const float af = alpha * tv + (1.f - alpha) * (1.f - tv);
      const float om = 1.f - pt;
      float w, dw;                                  // (1 - pt)^gamma and its derivative w.r.t. pt, negated below
      if (gamma == 2.f) {
        w = om * om;
        dw = 2.f * om;
      } else {
        w = powf(om, gamma);
        dw = om > 0.f ? gamma * powf(om, gamma - 1.f) : 0.f;
      }
      const float dpt = ds * (2.f * tv - 1.f);
      const float dfocal = af * (-dw * dpt * bce + w * (sig - tv));
      g[k] = cf * dfocal + ds * (c1 * tv + c2);
    }
    if (dl_bf16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(g[0], g[1]), c = __floats2bfloat162_rn(g[2], g[3]);
      reinterpret_cast<uint2*>(reinterpret_cast<char*>(dlogits) + (size_t)plane * n4 * 8)[i] =
          make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&c));
    } else {
      reinterpret_cast<float4*>(reinterpret_cast<char*>(dlogits) + (size_t)plane * n4 * 16)[i] = make_float4(g[0], g[1], g[2], g[3]);
    }
  }
}

int blocks_per_plane(long long n4) {
  long long b = (n4 + kLossThreads * 4 - 1) / (kLossThreads * 4);
  if (b < 1) b = 1;
  if (b > kMaxBlocksPerPlane) b = kMaxBlocksPerPlane;
  return (int)b;
}

}  // namespace
}  // namespace saicv

using namespace saicv;
#define ST ((cudaStream_t)stream)

extern "C" {

int saicv_sam_loss_partial_floats(int b, int m, long long n) { return b * m * blocks_per_plane(n / 4) * 6; }

int saicv_sam_loss_sums(const void* logits, int logits_bf16, const float* targets, float* partials, float* sums, int b, int m,
                        long long n, float alpha, float gamma, float mask_threshold, void* stream) {
  if (n % 4 || n <= 0) return set_error("saicv_sam_loss_sums: the plane size must be a positive multiple of 4 (got %lld)", n);
  const int nblk = blocks_per_plane(n / 4);
  sam_loss_sums_kernel<<<dim3(nblk, b * m), kLossThreads, 0, ST>>>(logits, logits_bf16, targets, partials, m, n / 4, alpha, gamma,
                                                                    mask_threshold);
  if (int e = check_launch("sam_loss_sums_kernel")) return e;
  sam_loss_fold_kernel<<<(b * m * 6 + 127) / 128, 128, 0, ST>>>(partials, sums, b * m, nblk);
  return check_launch("sam_loss_fold_kernel");
}

int saicv_sam_loss_bwd(const void* logits, int logits_bf16, const float* targets, const float* coef, void* dlogits, int dl_bf16, int b,
                       int m, long long n, float alpha, float gamma, void* stream) {
  if (n % 4 || n <= 0) return set_error("saicv_sam_loss_bwd: the plane size must be a positive multiple of 4 (got %lld)", n);
  const int nblk = blocks_per_plane(n / 4);
  sam_loss_bwd_kernel<<<dim3(nblk, b * m), kLossThreads, 0, ST>>>(logits, logits_bf16, targets, coef, dlogits, dl_bf16, m, n / 4, alpha,
                                                                   gamma);
  return check_launch("sam_loss_bwd_kernel");
}

}  // extern "C"
