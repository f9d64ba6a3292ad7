// SAM image-encoder specific kernels (SimpleAICV/interactive_segmentation/models/segment_anything/image_encoder.py):
//   window partition / unpartition with zero padding (:32-79),
//   decomposed relative-position bias (:82-144): the bias rel_h[q, kh] + rel_w[q, kw] is turned into extra score
//   columns so that the tcgen05 attention kernel adds it on the tensor cores:
//       Qe[l] = [ q[l] * scale | rel_h[l, 0..Sh) | rel_w[l, 0..Sw) | 0 ]      rel_h[l, kh] = q[l] . Rh[qh(l) - kh + Sh - 1]
//       Ke[l] = [ k[l]         | onehot(kh(l))   | onehot(kw(l))   | 0 ]
//       Qe . Ke^T = scale * q.k + rel_h[q, kh(k)] + rel_w[q, kw(k)]            (= image_encoder.py:171-174)
//   and the matching backward (dq from dQe, d rel_pos_h / d rel_pos_w with a fixed-order two-stage reduction),
//   the positional-embedding broadcast add (:315).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/saicv_b200.h"
#include "host_util.h"
#include "vec8.cuh"

namespace saicv {
namespace {


int grid1d(long long items, int per_block = 256, int cap = 148 * 16) {
  long long b = (items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// windows[(b*nwy+wy)*nwx+wx][iy*ws+ix][c] = x[b][wy*ws+iy][wx*ws+ix][c]  (0 outside H x W)
__global__ void window_partition_kernel(const V8* __restrict__ x, V8* __restrict__ win, int B, int H, int W, int C, int ws, int nwy,
                                        int nwx) {
  const int vpr = C >> 3;
  const long long total = (long long)B * nwy * nwx * ws * ws * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    long long t = i / vpr;
    const int ix = (int)(t % ws); t /= ws;
    const int iy = (int)(t % ws); t /= ws;
    const int wx = (int)(t % nwx); t /= nwx;
    const int wy = (int)(t % nwy);
    const long long b = t / nwy;
    const int h = wy * ws + iy, w = wx * ws + ix;
    V8 val;
    val.zero();
    if (h < H && w < W) val = x[((b * H + h) * W + w) * vpr + v];
    win[i] = val;
  }
}
// x[b][h][w][c] = windows[...] (padding rows dropped)
__global__ void window_unpartition_kernel(const V8* __restrict__ win, V8* __restrict__ x, int B, int H, int W, int C, int ws,
                                          int nwy, int nwx) {
  const int vpr = C >> 3;
  const long long total = (long long)B * H * W * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    long long t = i / vpr;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const long long b = t / H;
    const int wy = h / ws, iy = h % ws, wx = w / ws, ix = w % ws;
    x[i] = win[((((b * nwy + wy) * nwx + wx) * ws + iy) * ws + ix) * vpr + v];
  }
}

// x[b][r][c] += pos[r][c]   (fp32, in place)
__global__ void add_pos_kernel(float4* __restrict__ x, const float4* __restrict__ pos, long long per_batch4, long long total4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = x[i];
    const float4 p = pos[i % per_batch4];
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    x[i] = a;
  }
}

// ---- rel-pos score columns.  qkv: [Bw][L][3][H][HD] bf16 (L = Sh*Sw tokens, row-major (qh, qw)).
// One thread per (bw, head, token).  rel_pos tables live in shared memory as fp32 rounded to bf16 (the
// reference's einsum consumes them in bf16 under autocast).
template <int HD>
__global__ void __launch_bounds__(128)
relpos_build_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rph, const float* __restrict__ rpw,
                    __nv_bfloat16* __restrict__ qe, __nv_bfloat16* __restrict__ ke, int Bw, int H, int Sh, int Sw, int DQK,
                    float scale) {
  extern __shared__ float stab[];   // [2Sh-1][HD] then [2Sw-1][HD]
  float* th = stab;
  float* tw = stab + (2 * Sh - 1) * HD;
  for (int i = threadIdx.x; i < (2 * Sh - 1) * HD; i += blockDim.x) th[i] = __bfloat162float(__float2bfloat16_rn(rph[i]));
  for (int i = threadIdx.x; i < (2 * Sw - 1) * HD; i += blockDim.x) tw[i] = __bfloat162float(__float2bfloat16_rn(rpw[i]));
  __syncthreads();
  const int L = Sh * Sw;
  const long long total = (long long)Bw * H * L;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(r % L);
    const long long bh = r / L;
    const int h = (int)(bh % H);
    const long long bw = bh / H;
    const int qh = l / Sw, qw = l % Sw;
    const __nv_bfloat16* qrow = qkv + ((bw * L + l) * 3) * (long long)H * HD + (long long)h * HD;
    const __nv_bfloat16* krow = qrow + (long long)H * HD;
    float q[HD];
#pragma unroll
    for (int i = 0; i < HD; i += 8) {
      const V8 v = *reinterpret_cast<const V8*>(qrow + i);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(v.h(k));
        q[i + 2 * k] = f.x;
        q[i + 2 * k + 1] = f.y;
      }
    }
    __nv_bfloat16* qo = qe + r * DQK;
    __nv_bfloat16* ko = ke + r * DQK;
#pragma unroll
    for (int i = 0; i < HD; i += 8) {
      V8 o;
#pragma unroll
      for (int k = 0; k < 4; ++k) o.set(k, __floats2bfloat162_rn(q[i + 2 * k] * scale, q[i + 2 * k + 1] * scale));
      *reinterpret_cast<V8*>(qo + i) = o;
      *reinterpret_cast<V8*>(ko + i) = *reinterpret_cast<const V8*>(krow + i);
    }
    for (int j = 0; j < Sh + Sw; ++j) {
      const float* t = j < Sh ? th + (qh - j + Sh - 1) * HD : tw + (qw - (j - Sh) + Sw - 1) * HD;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < HD; i += 4) {
        const float4 tv = *reinterpret_cast<const float4*>(t + i);
        acc = fmaf(q[i], tv.x, acc);
        acc = fmaf(q[i + 1], tv.y, acc);
        acc = fmaf(q[i + 2], tv.z, acc);
        acc = fmaf(q[i + 3], tv.w, acc);
      }
      qo[HD + j] = __float2bfloat16_rn(acc);
      ko[HD + j] = __float2bfloat16_rn((j < Sh ? (j == qh) : (j - Sh == qw)) ? 1.f : 0.f);
    }
    for (int j = HD + Sh + Sw; j < DQK; ++j) {
      qo[j] = __float2bfloat16_rn(0.f);
      ko[j] = __float2bfloat16_rn(0.f);
    }
  }
}

// dq[l] = scale * dQe[l][0..HD) + sum_kh dQe[l][HD+kh] * Rh[qh-kh+Sh-1] + sum_kw dQe[l][HD+Sh+kw] * Rw[qw-kw+Sw-1]
// written into the q slot of dqkv [Bw][L][3][H][HD].
template <int HD>
__global__ void __launch_bounds__(128)
relpos_bwd_dq_kernel(const __nv_bfloat16* __restrict__ dqe, const float* __restrict__ rph, const float* __restrict__ rpw,
                     __nv_bfloat16* __restrict__ dqkv, int Bw, int H, int Sh, int Sw, int DQK, float scale) {
  extern __shared__ float stab[];
  float* th = stab;
  float* tw = stab + (2 * Sh - 1) * HD;
  for (int i = threadIdx.x; i < (2 * Sh - 1) * HD; i += blockDim.x) th[i] = __bfloat162float(__float2bfloat16_rn(rph[i]));
  for (int i = threadIdx.x; i < (2 * Sw - 1) * HD; i += blockDim.x) tw[i] = __bfloat162float(__float2bfloat16_rn(rpw[i]));
  __syncthreads();
  const int L = Sh * Sw;
  const long long total = (long long)Bw * H * L;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(r % L);
    const long long bh = r / L;
    const int h = (int)(bh % H);
    const long long bw = bh / H;
    const int qh = l / Sw, qw = l % Sw;
    const __nv_bfloat16* g = dqe + r * DQK;
    float acc[HD];
#pragma unroll
    for (int i = 0; i < HD; i += 8) {
      const V8 v = *reinterpret_cast<const V8*>(g + i);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(v.h(k));
        acc[i + 2 * k] = f.x * scale;
        acc[i + 2 * k + 1] = f.y * scale;
      }
    }
    for (int j = 0; j < Sh + Sw; ++j) {
      const float* t = j < Sh ? th + (qh - j + Sh - 1) * HD : tw + (qw - (j - Sh) + Sw - 1) * HD;
      const float d = __bfloat162float(g[HD + j]);
#pragma unroll
      for (int i = 0; i < HD; i += 4) {
        const float4 tv = *reinterpret_cast<const float4*>(t + i);
        acc[i] = fmaf(d, tv.x, acc[i]);
        acc[i + 1] = fmaf(d, tv.y, acc[i + 1]);
        acc[i + 2] = fmaf(d, tv.z, acc[i + 2]);
        acc[i + 3] = fmaf(d, tv.w, acc[i + 3]);
      }
    }
    __nv_bfloat16* out = dqkv + ((bw * L + l) * 3) * (long long)H * HD + (long long)h * HD;
#pragma unroll
    for (int i = 0; i < HD; i += 8) {
      V8 o;
#pragma unroll
      for (int k = 0; k < 4; ++k) o.set(k, __floats2bfloat162_rn(acc[i + 2 * k], acc[i + 2 * k + 1]));
      *reinterpret_cast<V8*>(out + i) = o;
    }
  }
}

// Table gradients as ONE tensor-core GEMM.  d rel_pos_h[idx][c] = sum over rows r and key rows kh with
// qh(r) - kh + Sh - 1 == idx of dT_h[r][kh] * q[r][c] (likewise for the width table).  Re-indexing the bias gradients per
// row, ef[r][idx] = dT_h[r][qh(r) - idx + Sh - 1] and ef[r][nh + idx] = dT_w[r][qw(r) - idx + Sw - 1] (zero where the
// key index falls outside the grid, and in the padding columns up to NIP), turns both sums into
//     [d rel_pos_h ; d rel_pos_w] = ef^T [NIP x rows] * qc [rows x HD],
// the weight-gradient form of the GEMM engine (saicv_linear_wgrad).  This kernel writes ef and the compact copy qc of
// the (unscaled) q rows; rows are ordered (window, head, token).  One thread per 16-byte piece.
template <int HD>
__global__ void __launch_bounds__(256)
relpos_shift_kernel(const __nv_bfloat16* __restrict__ dqe, const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ ef,
                    __nv_bfloat16* __restrict__ qc, int Bw, int H, int Sh, int Sw, int DQK, int NIP) {
  const int L = Sh * Sw;
  const int ef_pieces = NIP >> 3, pieces = ef_pieces + HD / 8;
  const int nh = 2 * Sh - 1, nw = 2 * Sw - 1;
  const long long total = (long long)Bw * H * L * pieces;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pc = (int)(i % pieces);
    const long long r = i / pieces;
    const int l = (int)(r % L);
    if (pc >= ef_pieces) {
      const long long bh = r / L;
      const int h = (int)(bh % H);
      const long long bw = bh / H;
      const int c0 = (pc - ef_pieces) * 8;
      *reinterpret_cast<uint4*>(qc + r * HD + c0) =
          __ldg(reinterpret_cast<const uint4*>(qkv + ((bw * L + l) * 3) * (long long)H * HD + (long long)h * HD + c0));
      continue;
    }
    const int qh = l / Sw, qw = l % Sw;
    const __nv_bfloat16* g = dqe + r * DQK + HD;
    const unsigned short* gs = reinterpret_cast<const unsigned short*>(g);
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned short v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = pc * 8 + 2 * k + e;
        int src = -1;
        if (idx < nh) {
          const int kh = qh - idx + Sh - 1;
          if (kh >= 0 && kh < Sh) src = kh;
        } else if (idx < nh + nw) {
          const int kw = qw - (idx - nh) + Sw - 1;
          if (kw >= 0 && kw < Sw) src = Sh + kw;
        }
        v[e] = src >= 0 ? __ldg(gs + src) : (unsigned short)0;
      }
      w[k] = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
    }
    *reinterpret_cast<uint4*>(ef + r * NIP + pc * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace
}  // namespace saicv

using namespace saicv;
#define ST ((cudaStream_t)stream)

extern "C" {

int saicv_window_partition(const void* x, void* windows, int b, int h, int w, int c, int ws, void* stream) {
  if (c % 8 || ws < 1) return set_error("saicv_window_partition: C %% 8 != 0 or bad window size");
  const int nwy = (h + ws - 1) / ws, nwx = (w + ws - 1) / ws;
  window_partition_kernel<<<grid1d((long long)b * nwy * nwx * ws * ws * (c / 8)), 256, 0, ST>>>(
      reinterpret_cast<const V8*>(x), reinterpret_cast<V8*>(windows), b, h, w, c, ws, nwy, nwx);
  return check_launch("window_partition_kernel");
}

int saicv_window_unpartition(const void* windows, void* x, int b, int h, int w, int c, int ws, void* stream) {
  if (c % 8 || ws < 1) return set_error("saicv_window_unpartition: C %% 8 != 0 or bad window size");
  const int nwy = (h + ws - 1) / ws, nwx = (w + ws - 1) / ws;
  window_unpartition_kernel<<<grid1d((long long)b * h * w * (c / 8)), 256, 0, ST>>>(
      reinterpret_cast<const V8*>(windows), reinterpret_cast<V8*>(x), b, h, w, c, ws, nwy, nwx);
  return check_launch("window_unpartition_kernel");
}

int saicv_add_pos_embed(float* x, const float* pos, int b, long long per_batch, void* stream) {
  if (per_batch % 4) return set_error("saicv_add_pos_embed: per-batch size %% 4 != 0");
  add_pos_kernel<<<grid1d((long long)b * per_batch / 4), 256, 0, ST>>>(reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(pos),
                                                                      per_batch / 4, (long long)b * per_batch / 4);
  return check_launch("add_pos_kernel");
}

#define SAM_HD_DISPATCH(KERNEL, ...)                                                                      \
  switch (hd) {                                                                                           \
    case 32: KERNEL<32> __VA_ARGS__; break;                                                               \
    case 64: KERNEL<64> __VA_ARGS__; break;                                                               \
    case 80: KERNEL<80> __VA_ARGS__; break;                                                               \
    default: return set_error("SAM rel-pos kernels: unsupported head dim %d (32, 64, 80)", hd);            \
  }

int saicv_relpos_build(const void* qkv, const float* rel_pos_h, const float* rel_pos_w, void* qe, void* ke, int bw, int heads,
                       int hd, int sh, int sw, int dqk, float scale, void* stream) {
  if (dqk < hd + sh + sw || dqk % 8) return set_error("saicv_relpos_build: dqk %d too small for %d + %d + %d", dqk, hd, sh, sw);
  const size_t smem = (size_t)(2 * sh - 1 + 2 * sw - 1) * hd * 4;
  const long long rows = (long long)bw * heads * sh * sw;
  const int grid = grid1d(rows, 128, 148 * 8);
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(qe);
  __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(ke);
  if (smem > 48 * 1024) {
    if (hd == 32) cudaFuncSetAttribute(relpos_build_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (hd == 64) cudaFuncSetAttribute(relpos_build_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (hd == 80) cudaFuncSetAttribute(relpos_build_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  SAM_HD_DISPATCH(relpos_build_kernel, <<<grid, 128, smem, ST>>>(q, rel_pos_h, rel_pos_w, o1, o2, bw, heads, sh, sw, dqk, scale))
  return check_launch("relpos_build_kernel");
}

int saicv_relpos_bwd(const void* dqe, const void* qkv, const float* rel_pos_h, const float* rel_pos_w, void* dqkv, void* ef,
                     void* qc, int nip, int bw, int heads, int hd, int sh, int sw, int dqk, float scale, void* stream) {
  if (sh + sw > 128) return set_error("saicv_relpos_bwd: Sh + Sw must be <= 128");
  if (nip % 8 || nip < 2 * sh - 1 + 2 * sw - 1) return set_error("saicv_relpos_bwd: nip must be a multiple of 8 covering both tables");
  const size_t smem = (size_t)(2 * sh - 1 + 2 * sw - 1) * hd * 4;
  const long long rows = (long long)bw * heads * sh * sw;
  const int grid = grid1d(rows, 128, 148 * 8);
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(dqe);
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
  if (smem > 48 * 1024) {
    if (hd == 32) cudaFuncSetAttribute(relpos_bwd_dq_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (hd == 64) cudaFuncSetAttribute(relpos_bwd_dq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (hd == 80) cudaFuncSetAttribute(relpos_bwd_dq_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  SAM_HD_DISPATCH(relpos_bwd_dq_kernel, <<<grid, 128, smem, ST>>>(g, rel_pos_h, rel_pos_w, dq, bw, heads, sh, sw, dqk, scale))
  if (int e = check_launch("relpos_bwd_dq_kernel")) return e;
  __nv_bfloat16* efp = reinterpret_cast<__nv_bfloat16*>(ef);
  __nv_bfloat16* qcp = reinterpret_cast<__nv_bfloat16*>(qc);
  const int sgrid = grid1d(rows * (nip / 8 + hd / 8));
  SAM_HD_DISPATCH(relpos_shift_kernel, <<<sgrid, 256, 0, ST>>>(g, q, efp, qcp, bw, heads, sh, sw, dqk, nip))
  return check_launch("relpos_shift_kernel");
}

}  // extern "C"
