// ViT-specific kernels of the hot path (SimpleAICV/classification/backbones/vit.py): LayerNorm
// forward/backward with the residual-stream add fused, exact-erf GELU, token assembly
// (cls + positional embedding), token pooling, and fused multi-head attention forward/backward
// (softmax(QK^T*scale)V without materialising the L x L matrix) for short sequences (L <= 256,
// head_dim 64) on mma.sync tensor-core tiles.
//
// Numerics follow the reference under autocast: the residual stream, LayerNorm
// statistics and softmax are fp32; GEMM / attention operands are bf16 with fp32 accumulation.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/saicv_b200.h"
#include "host_util.h"

namespace saicv {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
}

// ----------------------------------------------------------------------------- LayerNorm
// One warp per row; each lane owns NCH float4 chunks (columns i*128 + lane*4 .. +3).
template <int NCH>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ stats, long long M, float eps) {
  constexpr int C = NCH * 128;
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  float4 g[NCH], b[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    g[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    b[i] = *reinterpret_cast<const float4*>(beta + i * 128 + lane * 4);
  }
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps) {
    float4 v[NCH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      v[i] = *reinterpret_cast<const float4*>(x + row * C + i * 128 + lane * 4);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + bb * bb + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    if (lane == 0) {
      stats[row] = mean;
      stats[M + row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const uint32_t lo = pack2((v[i].x - mean) * rstd * g[i].x + b[i].x, (v[i].y - mean) * rstd * g[i].y + b[i].y);
      const uint32_t hi = pack2((v[i].z - mean) * rstd * g[i].z + b[i].z, (v[i].w - mean) * rstd * g[i].w + b[i].w);
      *reinterpret_cast<uint2*>(y + row * C + i * 128 + lane * 4) = make_uint2(lo, hi);
    }
  }
}

// dx = dres + rstd*(dxhat - mean(dxhat) - xhat*mean(dxhat*xhat)), dxhat = dy*gamma; dgamma/dbeta
// accumulated per warp in registers, folded through shared memory in a fixed order into one
// partial row per block ([gridDim.x][2][C]); ln_fold_kernel sums the rows, again in a fixed order,
// so the parameter gradients are bit-reproducible (no atomics).
template <int NCH>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ stats, const float* __restrict__ dres, float* __restrict__ dx,
              __nv_bfloat16* __restrict__ dx_bf16, const float* __restrict__ row_scale, int rows_per_scale,
              float* __restrict__ partials, long long M) {
  constexpr int C = NCH * 128;
  __shared__ float red[8][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  float4 g[NCH], dg[NCH], db[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    g[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp; row < M; row += warps) {
    const float mean = stats[row], rstd = stats[M + row];
    const float bsc = row_scale ? row_scale[row / rows_per_scale] : 1.f;
    float4 xh[NCH], d[NCH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float4 xv = *reinterpret_cast<const float4*>(x + row * C + i * 128 + lane * 4);
      const uint2 dv = *reinterpret_cast<const uint2*>(dy + row * C + i * 128 + lane * 4);
      const float2 d01 = unpack2(dv.x), d23 = unpack2(dv.y);
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      db[i].x += d01.x; db[i].y += d01.y; db[i].z += d23.x; db[i].w += d23.y;
      dg[i].x += d01.x * xh[i].x; dg[i].y += d01.y * xh[i].y; dg[i].z += d23.x * xh[i].z; dg[i].w += d23.y * xh[i].w;
      d[i] = make_float4(d01.x * g[i].x, d01.y * g[i].y, d23.x * g[i].z, d23.y * g[i].w);
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
    }
    const float c1 = warp_sum(s1) * (1.f / C), c2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float4 o = make_float4(rstd * (d[i].x - c1 - xh[i].x * c2), rstd * (d[i].y - c1 - xh[i].y * c2),
                             rstd * (d[i].z - c1 - xh[i].z * c2), rstd * (d[i].w - c1 - xh[i].w * c2));
      if (dres) {
        const float4 r = *reinterpret_cast<const float4*>(dres + row * C + i * 128 + lane * 4);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(dx + row * C + i * 128 + lane * 4) = o;
      if (dx_bf16)
        *reinterpret_cast<uint2*>(dx_bf16 + row * C + i * 128 + lane * 4) =
            make_uint2(pack2(o.x * bsc, o.y * bsc), pack2(o.z * bsc, o.w * bsc));
    }
  }
  // fold the 8 warps' partial dgamma / dbeta
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      *reinterpret_cast<float4*>(&red[warp][i * 128 + lane * 4]) = pass == 0 ? dg[i] : db[i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][c];
      partials[((long long)blockIdx.x * 2 + pass) * C + c] = s;
    }
  }
}

// dgamma / dbeta (+)= sum over the per-block partial rows, fixed order
__global__ void ln_fold_kernel(const float* __restrict__ partials, float* __restrict__ dgamma, float* __restrict__ dbeta,
                               int nblk, int C, int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 2C-1
  if (j >= 2 * C) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partials[(long long)b * 2 * C + j];
  float* dst = j < C ? dgamma + j : dbeta + (j - C);
  *dst = accumulate ? *dst + s : s;
}

// ----------------------------------------------------------------------------- GELU (exact, erf)
__global__ void gelu_fwd_kernel(const uint4* __restrict__ u, uint4* __restrict__ h, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = u[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2(w[k]);
      w[k] = pack2(0.5f * f.x * (1.f + erff(f.x * 0.70710678118654752f)), 0.5f * f.y * (1.f + erff(f.y * 0.70710678118654752f)));
    }
    h[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
// du = dh * gelu'(u)
__global__ void gelu_bwd_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ u, uint4* __restrict__ du,
                                long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 a = dh[i], b = u[i];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 d = unpack2(aw[k]), x = unpack2(bw[k]);
      o[k] = pack2(d.x * gelu_grad(x.x), d.y * gelu_grad(x.y));
    }
    du[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ----------------------------------------------------------------------------- tokens
// x[b, 0] = cls + pos[0];  x[b, 1+i] = patch[b*NP+i] + pos[1+i]      (vit.py:242-243)
__global__ void assemble_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float* __restrict__ x, int B, int NP, int C) {
  const int c4 = C >> 2;
  const long long total = (long long)B * (NP + 1) * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    const long long t = i / c4;
    const int l = (int)(t % (NP + 1));
    const long long b = t / (NP + 1);
    const float4 p = reinterpret_cast<const float4*>(pos)[(long long)l * c4 + c];
    const float4 s = l == 0 ? reinterpret_cast<const float4*>(cls)[c]
                            : reinterpret_cast<const float4*>(patch)[(b * NP + l - 1) * c4 + c];
    reinterpret_cast<float4*>(x)[i] = make_float4(s.x + p.x, s.y + p.y, s.z + p.z, s.w + p.w);
  }
}
// dpos[l] (+)= sum_b dx[b,l]; dcls (+)= sum_b dx[b,0]; dpatch[b*NP+i] = bf16(dx[b,1+i])
__global__ void assemble_tokens_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dpos, float* __restrict__ dcls,
                                           __nv_bfloat16* __restrict__ dpatch, int B, int NP, int C, int accumulate) {
  const long long total = (long long)(NP + 1) * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int l = (int)(i / C);
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float v = dx[((long long)b * (NP + 1) + l) * C + c];
      s += v;
      if (l > 0) dpatch[((long long)b * NP + l - 1) * C + c] = __float2bfloat16_rn(v);
    }
    dpos[i] = accumulate ? dpos[i] + s : s;
    if (l == 0) dcls[c] = accumulate ? dcls[c] + s : s;
  }
}
// pooled[b] = mean over tokens 1..L-1 (global_pool, vit.py:252-255) or x[b, 0] (cls token, :258)
__global__ void token_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ pooled, int B, int L, int C, int mean_pool) {
  const long long total = (long long)B * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long b = i / C;
    if (!mean_pool) {
      pooled[i] = x[b * L * C + c];
    } else {
      float s = 0.f;
      for (int l = 1; l < L; ++l) s += x[(b * L + l) * C + c];
      pooled[i] = s / (float)(L - 1);
    }
  }
}
__global__ void token_pool_bwd_kernel(const float* __restrict__ dpooled, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16,
                                      const float* __restrict__ row_scale, int B, int L, int C, int mean_pool) {
  const long long total = (long long)B * L * C;
  const float inv = 1.f / (float)(L - 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int l = (int)(t % L);
    const long long b = t / L;
    float v;
    if (mean_pool) v = l == 0 ? 0.f : dpooled[b * C + c] * inv;
    else v = l == 0 ? dpooled[b * C + c] : 0.f;
    dx[i] = v;
    if (dx_bf16) dx_bf16[i] = __float2bfloat16_rn(row_scale ? v * row_scale[b] : v);
  }
}

// ----------------------------------------------------------------------------- attention
// Tensor-core tiles: mma.sync.m16n8k16 (bf16 x bf16 -> fp32).  Fragment layouts (g = lane/4, t = lane%4):
//   A (16x16 row-major): a0=(g, 2t..) a1=(g+8, 2t..) a2=(g, 2t+8..) a3=(g+8, 2t+8..)
//   B (16x8, k-major per n): b0=(k=2t.., n=g) b1=(k=2t+8.., n=g)
//   C (16x8): c0,c1=(g, 2t,2t+1) c2,c3=(g+8, 2t,2t+1)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// B fragment of a 16(k) x 8(n) tile whose storage is [k][n] (n contiguous): transposing ldmatrix.
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

constexpr int HD = 64;      // head dim
constexpr int LDS = 72;     // smem row stride (bf16): 144 B keeps 32-bit fragment loads and ldmatrix conflict free

// Stage `rows` rows of one head's q / k / v (or of a [B, L, H*D] tensor) into smem, zero padded.
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, long long row_stride, int row0,
                                           int nrows, int L) {
  for (int i = threadIdx.x; i < nrows * (HD / 8); i += blockDim.x) {
    const int r = i / (HD / 8), v = i % (HD / 8);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (row0 + r < L) val = *reinterpret_cast<const uint4*>(src + (long long)(row0 + r) * row_stride + v * 8);
    *reinterpret_cast<uint4*>(dst + r * LDS + v * 8) = val;
  }
}
// A fragments (16 rows x 64 k) of the row-major tile starting at row0: one ldmatrix.x4 per k16 step
// (matrices: rows 0-7 / 8-15 x k 0-7 / 8-15 = a0, a1, a2, a3).
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const __nv_bfloat16* s, int row0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4(a[kk], s + (row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + kk * 16 + (lane >> 4) * 8);
}
// acc (16 x 8 tile) = A(16 x 64) * Bsrc[rows n0 .. n0+7][0..63]^T ; Bsrc row-major (k contiguous)
__device__ __forceinline__ void gemm_nt_16x8(float (&acc)[4], const uint32_t (&a)[4][4], const __nv_bfloat16* bsrc,
                                             int n0, int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = 0.f;
#pragma unroll
  for (int kk2 = 0; kk2 < 2; ++kk2) {
    uint32_t b[4];  // (b0, b1) of k16 step 2*kk2 and of step 2*kk2+1
    ldsm_x4(b, bsrc + (n0 + (lane & 7)) * LDS + kk2 * 32 + (lane >> 3) * 8);
    mma16816(acc, a[2 * kk2], b[0], b[1]);
    mma16816(acc, a[2 * kk2 + 1], b[2], b[3]);
  }
}
__device__ __forceinline__ void gemm_nt_16x16(float (&acc)[2][4], const uint32_t (&a)[4][4], const __nv_bfloat16* bsrc,
                                              int n0, int lane) {
  gemm_nt_16x8(acc[0], a, bsrc, n0, lane);
  gemm_nt_16x8(acc[1], a, bsrc, n0 + 8, lane);
}
// acc[dt] (16 x 64 output, 8 d-tiles) += P(16 x 16, as A fragment) * Bsrc[rows k0 .. k0+15][0..63]
// Bsrc is [k][n] (n contiguous): transposing ldmatrix.x4 yields (b0, b1) of two d-tiles at once.
__device__ __forceinline__ void gemm_pv_16x64(float (&acc)[8][4], const uint32_t (&pa)[4], const __nv_bfloat16* bsrc,
                                              int k0, int lane) {
#pragma unroll
  for (int dp = 0; dp < 4; ++dp) {
    uint32_t b[4];
    ldsm_x4_trans(b, bsrc + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (dp * 2 + (lane >> 4)) * 8);
    mma16816(acc[2 * dp], pa, b[0], b[1]);
    mma16816(acc[2 * dp + 1], pa, b[2], b[3]);
  }
}

// qkv: [B, L, 3, H, 64] bf16; out: [B, L, H*64] bf16; lse: [B, H, L] fp32 (log2 domain: m + log2(sum))
// One CTA per (batch, head), one warp per 16 query rows; K and V of the head are staged once
// (zero padded to a multiple of 64 keys) and consumed in 64-key blocks with an online softmax.
__global__ void __launch_bounds__(512)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int B,
                int L, int H, float scale_log2, int Lp, int Lk) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + Lp * LDS;
  __nv_bfloat16* sV = sK + Lk * LDS;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const long long rs = 3LL * H * HD;
  const __nv_bfloat16* base = qkv + (long long)b * L * rs + h * HD;
  stage_rows(sQ, base, rs, 0, Lp, L);
  stage_rows(sK, base + (long long)H * HD, rs, 0, Lk, L);
  stage_rows(sV, base + 2LL * H * HD, rs, 0, Lk, L);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint32_t qa[4][4];
  load_a_frags(qa, sQ, warp * 16, lane);
  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[dt][j] = 0.f;
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
  for (int kb = 0; kb < Lk; kb += 64) {
    float s[8][4];
    float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      gemm_nt_16x8(s[nt], qa, sK, kb + nt * 8, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kb + nt * 8 + 2 * t + (j & 1);
        s[nt][j] = key < L ? s[nt][j] * scale_log2 : -CUDART_INF_F;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: key kb is always < L
    const float al0 = exp2f(m0 - mn0), al1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m0); s[nt][1] = exp2f(s[nt][1] - m0);
      s[nt][2] = exp2f(s[nt][2] - m1); s[nt][3] = exp2f(s[nt][3] - m1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * al0 + rs0;
    l1 = l1 * al1 + rs1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= al0; o[dt][1] *= al0; o[dt][2] *= al1; o[dt][3] *= al1;
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t pa[4] = {pack2(s[2 * ks][0], s[2 * ks][1]), pack2(s[2 * ks][2], s[2 * ks][3]),
                              pack2(s[2 * ks + 1][0], s[2 * ks + 1][1]), pack2(s[2 * ks + 1][2], s[2 * ks + 1][3])};
      gemm_pv_16x64(o, pa, sV, kb + ks * 16, lane);
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  __nv_bfloat16* ob = out + (long long)b * L * H * HD + h * HD;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    if (r0 < L) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * H * HD + dt * 8 + 2 * t) = pack2(o[dt][0] * inv0, o[dt][1] * inv0);
    if (r1 < L) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * H * HD + dt * 8 + 2 * t) = pack2(o[dt][2] * inv1, o[dt][3] * inv1);
  }
  if (t == 0) {
    if (r0 < L) lse[(long long)bh * L + r0] = m0 + log2f(l0);
    if (r1 < L) lse[(long long)bh * L + r1] = m1 + log2f(l1);
  }
}

// A fragments (16 rows x 64 k) straight from global memory (rows >= L read as zero).
__device__ __forceinline__ void load_a_frags_global(uint32_t (&a)[4][4], const __nv_bfloat16* base, long long row_stride,
                                                    int row0, int L, int g, int t) {
  const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = r0 < L ? *reinterpret_cast<const uint32_t*>(base + (long long)r0 * row_stride + kk * 16 + 2 * t) : 0u;
    a[kk][1] = r1 < L ? *reinterpret_cast<const uint32_t*>(base + (long long)r1 * row_stride + kk * 16 + 2 * t) : 0u;
    a[kk][2] = r0 < L ? *reinterpret_cast<const uint32_t*>(base + (long long)r0 * row_stride + kk * 16 + 8 + 2 * t) : 0u;
    a[kk][3] = r1 < L ? *reinterpret_cast<const uint32_t*>(base + (long long)r1 * row_stride + kk * 16 + 8 + 2 * t) : 0u;
  }
}
// sum_d a[row][d] * b[row][d] over the 64 head dims, computed by the 4 lanes of a quad
__device__ __forceinline__ float row_dot64(const __nv_bfloat16* a, const __nv_bfloat16* b, int t) {
  float acc = 0.f;
  const uint4* pa = reinterpret_cast<const uint4*>(a + t * 16);
  const uint4* pb = reinterpret_cast<const uint4*>(b + t * 16);
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const uint4 x = pa[v], y = pb[v];
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fx = unpack2(xw[k]), fy = unpack2(yw[k]);
      acc += fx.x * fy.x + fx.y * fy.y;
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  return acc;
}

// Backward of the fused attention.  grid = (batch*heads, 2), Lp/16 warps per CTA, ~61 KB of shared
// memory so three CTAs share an SM:
//   blockIdx.y == 0: each warp owns 16 query rows -> dQ      (K, V staged in smem; Q, dO tiles from global)
//   blockIdx.y == 1: each warp owns 16 key rows   -> dK, dV  (Q, dO staged in smem; K, V tiles from global)
// P is recomputed from the saved log-sum-exp; dS = P * (dP - D) * scale with D = rowsum(dO * O).
// dqkv has the layout of qkv.
__global__ void __launch_bounds__(512)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv,
                int B, int L, int H, float scale, float scale_log2, int Lp, int Lq) {
  // Lp = L rounded up to 16 (one warp per 16 rows); Lq = L rounded up to 32 (rows staged, zero padded)
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(smem_attn);  // K (phase 0) or Q (phase 1)
  __nv_bfloat16* sY = sX + Lq * LDS;                                // V (phase 0) or dO (phase 1)
  float* sLse = reinterpret_cast<float*>(sY + Lq * LDS);
  float* sD = sLse + Lp;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int phase = blockIdx.y;
  const long long rs = 3LL * H * HD, os = (long long)H * HD;
  const __nv_bfloat16* qb = qkv + (long long)b * L * rs + h * HD;
  const __nv_bfloat16* kb_ = qb + (long long)H * HD;
  const __nv_bfloat16* vb = qb + 2LL * H * HD;
  const __nv_bfloat16* ob = out + (long long)b * L * os + h * HD;
  const __nv_bfloat16* dob = dout + (long long)b * L * os + h * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int tile0 = warp * 16;
  __nv_bfloat16* dbase = dqkv + (long long)b * L * rs + h * HD;
  if (phase == 0) {
    stage_rows(sX, kb_, rs, 0, Lq, L);
    stage_rows(sY, vb, rs, 0, Lq, L);
  } else {
    stage_rows(sX, qb, rs, 0, Lq, L);
    stage_rows(sY, dob, os, 0, Lq, L);
    for (int r = threadIdx.x; r < Lp; r += blockDim.x) sLse[r] = r < L ? lse[(long long)bh * L + r] : 0.f;
    // D[i] for every query: 4 lanes per row
    // (every lane of a warp runs the quad shuffles of row_dot64: rows >= L are clamped, not skipped;
    //  blockDim.x/4 divides Lp, so the trip count is warp-uniform)
    for (int r = (threadIdx.x >> 2); r < Lp; r += (blockDim.x >> 2)) {
      const int rc = min(r, L - 1);
      const float dsum = row_dot64(ob + (long long)rc * os, dob + (long long)rc * os, t);
      if (t == 0) sD[r] = r < L ? dsum : 0.f;
    }
  }
  __syncthreads();
  float acc1[8][4], acc2[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc1[dt][j] = acc2[dt][j] = 0.f;

  if (phase == 0) {
    // ---------------- dQ[tile] = sum_keys dS K
    uint32_t qa[4][4], da[4][4];
    load_a_frags_global(qa, qb, rs, tile0, L, g, t);
    load_a_frags_global(da, dob, os, tile0, L, g, t);
    const int r0 = tile0 + g, r1 = r0 + 8;
    const float ls0 = r0 < L ? lse[(long long)bh * L + r0] : 0.f, ls1 = r1 < L ? lse[(long long)bh * L + r1] : 0.f;
    const int r0c = min(r0, L - 1), r1c = min(r1, L - 1);  // clamp (not skip): warp-wide shuffles inside
    const float d0 = row_dot64(ob + (long long)r0c * os, dob + (long long)r0c * os, t);
    const float d1 = row_dot64(ob + (long long)r1c * os, dob + (long long)r1c * os, t);
    // 32 keys per step: 8 independent accumulation chains keep the tensor pipe busier
    for (int k0 = 0; k0 < Lq; k0 += 32) {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        gemm_nt_16x8(s[nt], qa, sX, k0 + nt * 8, lane);
        gemm_nt_16x8(dp[nt], da, sY, k0 + nt * 8, lane);
      }
      uint32_t pa[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float ds[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int key = k0 + nt * 8 + 2 * t + (j & 1);
          const float p = key < L ? exp2f(s[nt][j] * scale_log2 - (j < 2 ? ls0 : ls1)) : 0.f;
          ds[j] = p * (dp[nt][j] - (j < 2 ? d0 : d1)) * scale;
        }
        pa[nt >> 1][(nt & 1) * 2] = pack2(ds[0], ds[1]);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(ds[2], ds[3]);
      }
      gemm_pv_16x64(acc1, pa[0], sX, k0, lane);
      gemm_pv_16x64(acc1, pa[1], sX, k0 + 16, lane);
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      if (r0 < L) *reinterpret_cast<uint32_t*>(dbase + (long long)r0 * rs + dt * 8 + 2 * t) = pack2(acc1[dt][0], acc1[dt][1]);
      if (r1 < L) *reinterpret_cast<uint32_t*>(dbase + (long long)r1 * rs + dt * 8 + 2 * t) = pack2(acc1[dt][2], acc1[dt][3]);
    }
  } else {
    // ---------------- dV[tile] = sum_q P^T dO;  dK[tile] = sum_q dS^T Q
    uint32_t ka[4][4], va[4][4];
    load_a_frags_global(ka, kb_, rs, tile0, L, g, t);
    load_a_frags_global(va, vb, rs, tile0, L, g, t);
    const int key0 = tile0 + g, key1 = key0 + 8;
    for (int q0 = 0; q0 < Lp; q0 += 16) {
      float st[2][4], dpt[2][4];
      gemm_nt_16x16(st, ka, sX, q0, lane);     // S^T[key][query]
      gemm_nt_16x16(dpt, va, sY, q0, lane);    // dP^T[key][query]
      uint32_t pta[4], dsa[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float p[4], ds[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int qi = q0 + nt * 8 + 2 * t + (j & 1);
          const int key = j < 2 ? key0 : key1;
          p[j] = (qi < L && key < L) ? exp2f(st[nt][j] * scale_log2 - sLse[qi]) : 0.f;
          ds[j] = p[j] * (dpt[nt][j] - sD[qi]) * scale;
        }
        pta[nt * 2] = pack2(p[0], p[1]);
        pta[nt * 2 + 1] = pack2(p[2], p[3]);
        dsa[nt * 2] = pack2(ds[0], ds[1]);
        dsa[nt * 2 + 1] = pack2(ds[2], ds[3]);
      }
      gemm_pv_16x64(acc1, pta, sY, q0, lane);  // dV += P^T dO
      gemm_pv_16x64(acc2, dsa, sX, q0, lane);  // dK += dS^T Q
    }
    __nv_bfloat16* dk = dbase + (long long)H * HD;
    __nv_bfloat16* dv = dbase + 2LL * H * HD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      if (key0 < L) {
        *reinterpret_cast<uint32_t*>(dv + (long long)key0 * rs + dt * 8 + 2 * t) = pack2(acc1[dt][0], acc1[dt][1]);
        *reinterpret_cast<uint32_t*>(dk + (long long)key0 * rs + dt * 8 + 2 * t) = pack2(acc2[dt][0], acc2[dt][1]);
      }
      if (key1 < L) {
        *reinterpret_cast<uint32_t*>(dv + (long long)key1 * rs + dt * 8 + 2 * t) = pack2(acc1[dt][2], acc1[dt][3]);
        *reinterpret_cast<uint32_t*>(dk + (long long)key1 * rs + dt * 8 + 2 * t) = pack2(acc2[dt][2], acc2[dt][3]);
      }
    }
  }
}

int grid_1d(long long items, int per_block = 256, int cap = 148 * 16) {
  long long b = (items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

}  // namespace
}  // namespace saicv

using namespace saicv;
#define ST ((cudaStream_t)stream)

extern "C" {

int saicv_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* stats,
                        long long rows, int c, float eps, void* stream) {
  const int grid = grid_1d(rows, 8, 148 * 8);
  __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
  switch (c) {
    case 768: ln_fwd_kernel<6><<<grid, 256, 0, ST>>>(x, gamma, beta, yy, stats, rows, eps); break;
    case 1024: ln_fwd_kernel<8><<<grid, 256, 0, ST>>>(x, gamma, beta, yy, stats, rows, eps); break;
    case 1280: ln_fwd_kernel<10><<<grid, 256, 0, ST>>>(x, gamma, beta, yy, stats, rows, eps); break;
    case 256: ln_fwd_kernel<2><<<grid, 256, 0, ST>>>(x, gamma, beta, yy, stats, rows, eps); break;
    case 128: ln_fwd_kernel<1><<<grid, 256, 0, ST>>>(x, gamma, beta, yy, stats, rows, eps); break;
    default: return set_error("saicv_layernorm_fwd: unsupported width %d (128, 256, 768, 1024, 1280)", c);
  }
  return check_launch("ln_fwd_kernel");
}

int saicv_layernorm_bwd(const void* dy, const float* x, const float* gamma, const float* stats, const float* dres,
                        float* dx, void* dx_bf16, const float* bf16_row_scale, int rows_per_scale, float* partials,
                        float* dgamma, float* dbeta, long long rows, int c, int accumulate, void* stream) {
  if (bf16_row_scale && rows_per_scale <= 0) return set_error("saicv_layernorm_bwd: rows_per_scale must be > 0");
  if (!partials) return set_error("saicv_layernorm_bwd: needs a [SAICV_BN_PARTIAL_ROWS][2*c] fp32 workspace");
  const int grid = grid_1d(rows, 8 * 8, 148 * 2);
  const __nv_bfloat16* d = reinterpret_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  switch (c) {
    case 768: ln_bwd_kernel<6><<<grid, 256, 0, ST>>>(d, x, gamma, stats, dres, dx, xb, bf16_row_scale, rows_per_scale, partials, rows); break;
    case 1024: ln_bwd_kernel<8><<<grid, 256, 0, ST>>>(d, x, gamma, stats, dres, dx, xb, bf16_row_scale, rows_per_scale, partials, rows); break;
    case 1280: ln_bwd_kernel<10><<<grid, 256, 0, ST>>>(d, x, gamma, stats, dres, dx, xb, bf16_row_scale, rows_per_scale, partials, rows); break;
    case 256: ln_bwd_kernel<2><<<grid, 256, 0, ST>>>(d, x, gamma, stats, dres, dx, xb, bf16_row_scale, rows_per_scale, partials, rows); break;
    case 128: ln_bwd_kernel<1><<<grid, 256, 0, ST>>>(d, x, gamma, stats, dres, dx, xb, bf16_row_scale, rows_per_scale, partials, rows); break;
    default: return set_error("saicv_layernorm_bwd: unsupported width %d", c);
  }
  if (int e = check_launch("ln_bwd_kernel")) return e;
  ln_fold_kernel<<<(2 * c + 127) / 128, 128, 0, ST>>>(partials, dgamma, dbeta, grid, c, accumulate);
  return check_launch("ln_fold_kernel");
}

int saicv_gelu_fwd(const void* u, void* h, long long n, void* stream) {
  if (n % 8) return set_error("saicv_gelu_fwd: n %% 8 != 0");
  gelu_fwd_kernel<<<grid_1d(n / 8), 256, 0, ST>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<uint4*>(h), n / 8);
  return check_launch("gelu_fwd_kernel");
}

int saicv_gelu_bwd(const void* dh, const void* u, void* du, long long n, void* stream) {
  if (n % 8) return set_error("saicv_gelu_bwd: n %% 8 != 0");
  gelu_bwd_kernel<<<grid_1d(n / 8), 256, 0, ST>>>(reinterpret_cast<const uint4*>(dh), reinterpret_cast<const uint4*>(u),
                                                  reinterpret_cast<uint4*>(du), n / 8);
  return check_launch("gelu_bwd_kernel");
}

int saicv_vit_assemble_tokens(const float* patch, const float* cls, const float* pos, float* x, int b, int np, int c,
                              void* stream) {
  if (c % 4) return set_error("saicv_vit_assemble_tokens: C %% 4 != 0");
  assemble_tokens_kernel<<<grid_1d((long long)b * (np + 1) * (c / 4)), 256, 0, ST>>>(patch, cls, pos, x, b, np, c);
  return check_launch("assemble_tokens_kernel");
}

int saicv_vit_assemble_tokens_bwd(const float* dx, float* dpos, float* dcls, void* dpatch, int b, int np, int c,
                                  int accumulate, void* stream) {
  assemble_tokens_bwd_kernel<<<grid_1d((long long)(np + 1) * c), 256, 0, ST>>>(
      dx, dpos, dcls, reinterpret_cast<__nv_bfloat16*>(dpatch), b, np, c, accumulate);
  return check_launch("assemble_tokens_bwd_kernel");
}

int saicv_token_pool_fwd(const float* x, float* pooled, int b, int l, int c, int mean_pool, void* stream) {
  token_pool_fwd_kernel<<<grid_1d((long long)b * c), 256, 0, ST>>>(x, pooled, b, l, c, mean_pool);
  return check_launch("token_pool_fwd_kernel");
}

int saicv_token_pool_bwd(const float* dpooled, float* dx, void* dx_bf16, const float* bf16_row_scale, int b, int l,
                         int c, int mean_pool, void* stream) {
  token_pool_bwd_kernel<<<grid_1d((long long)b * l * c), 256, 0, ST>>>(dpooled, dx, reinterpret_cast<__nv_bfloat16*>(dx_bf16),
                                                                       bf16_row_scale, b, l, c, mean_pool);
  return check_launch("token_pool_bwd_kernel");
}

int saicv_attention_fwd(const void* qkv, void* out, float* lse, int b, int l, int h, int d, float scale, void* stream) {
  if (d != HD || l > 256 || l < 1) return set_error("saicv_attention_fwd: supports head_dim 64 and 1 <= L <= 256 (d=%d L=%d)", d, l);
  const int Lp = (l + 15) / 16 * 16, Lk = (l + 63) / 64 * 64;
  const size_t smem = (size_t)(Lp + 2 * Lk) * LDS * 2;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (256 + 2 * 256) * LDS * 2);
    attr = true;
  }
  attn_fwd_kernel<<<b * h, (Lp / 16) * 32, smem, ST>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                      reinterpret_cast<__nv_bfloat16*>(out), lse, b, l, h,
                                                      scale * 1.4426950408889634f, Lp, Lk);
  return check_launch("attn_fwd_kernel");
}

int saicv_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int b, int l,
                        int h, int d, float scale, void* stream) {
  if (d != HD || l > 256 || l < 1) return set_error("saicv_attention_bwd: supports head_dim 64 and 1 <= L <= 256 (d=%d L=%d)", d, l);
  const int Lp = (l + 15) / 16 * 16;
  const int Lq = (l + 31) / 32 * 32;
  const size_t smem = (size_t)2 * Lq * LDS * 2 + (size_t)2 * Lp * 4;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * LDS * 2 + 2 * 256 * 4);
    attr = true;
  }
  attn_bwd_kernel<<<dim3(b * h, 2), (Lp / 16) * 32, smem, ST>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(out),
      reinterpret_cast<const __nv_bfloat16*>(dout), lse, reinterpret_cast<__nv_bfloat16*>(dqkv), b, l, h, scale,
      scale * 1.4426950408889634f, Lp, Lq);
  return check_launch("attn_bwd_kernel");
}

}  // extern "C"
