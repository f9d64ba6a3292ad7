// HBM-bound kernels of the conv-net hot path: BatchNorm (training) statistics / apply /
// backward, ReLU + residual fusion, pooling, layout and weight preparation.  All activations
// are NHWC bf16 viewed as [rows][C]; every thread moves 16-byte vectors (8 x bf16), channel is the
// fastest dimension so warps read/write fully coalesced 512 B segments.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/saicv_b200.h"
#include "host_util.h"

namespace saicv {

static thread_local char g_err[512] = "";
int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e));
  count_launch(1);
  return 0;
}
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
const char* last_error() { return g_err; }
long long launch_count();

namespace {

constexpr int kThreads = 256;

struct alignas(16) V8 {
  __nv_bfloat162 h[4];
};
__device__ __forceinline__ void unpack8(const V8& v, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(v.h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ V8 pack8(const float (&f)[8]) {
  V8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ V8 ldg8(const void* p, long long vec_idx) {
  return *(reinterpret_cast<const V8*>(p) + vec_idx);
}
__device__ __forceinline__ void stg8(void* p, long long vec_idx, const V8& v) {
  *(reinterpret_cast<V8*>(p) + vec_idx) = v;
}
__device__ __forceinline__ float act_apply(float v, int act) {
  return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.1f * v) : v);
}
__device__ __forceinline__ float act_grad(float out, int act) {
  // derivative expressed through the activated output (sign is preserved by ReLU / LeakyReLU)
  return act == 1 ? (out > 0.f ? 1.f : 0.f) : (act == 2 ? (out > 0.f ? 1.f : 0.1f) : 1.f);
}

int grid_for(long long work_items, int per_block = kThreads, int max_blocks = 148 * 16) {
  long long b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}
int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ----------------------------------------------------------------------------- column sums
// MODE 0: sum x, sum x^2 (BN forward statistics)
// MODE 1: g = dout*act'(out); sum g, sum g*xhat (BN backward reductions)
// MODE 2: sum x only (bias gradients)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
colreduce_kernel(const void* __restrict__ a, const void* __restrict__ b, const void* __restrict__ y,
                 const float* __restrict__ saved, float* __restrict__ out, long long rows, int C,
                 int tx_count, int act, long long rows_per_block) {
  __shared__ float red[kThreads][17];
  const int vpr = C >> 3;
  const int tx = threadIdx.x % tx_count;
  const int ty = threadIdx.x / tx_count;
  const int ty_count = kThreads / tx_count;
  float s0[8], s1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  if (tx < vpr) {
    float mean[8], rstd[8];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mean[i] = saved[tx * 8 + i];
        rstd[i] = saved[C + tx * 8 + i];
      }
    }
    for (long long r = r0 + ty; r < r1; r += ty_count) {
      const long long vi = r * vpr + tx;
      float fa[8];
      unpack8(ldg8(a, vi), fa);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] += fa[i];
          s1[i] += fa[i] * fa[i];
        }
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s0[i] += fa[i];
      } else {
        float fy[8];
        unpack8(ldg8(y, vi), fy);
        if (act != 0) {
          float fo[8];
          unpack8(ldg8(b, vi), fo);
#pragma unroll
          for (int i = 0; i < 8; ++i) fa[i] *= act_grad(fo[i], act);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] += fa[i];
          s1[i] += fa[i] * (fy[i] - mean[i]) * rstd[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x][i] = s0[i];
    red[threadIdx.x][8 + i] = s1[i];
  }
  __syncthreads();
  // threads of row ty == 0 fold the partials of the other rows sharing their channel vector
  if (ty == 0 && tx < vpr) {
    for (int t = 1; t < ty_count; ++t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s0[i] += red[t * tx_count + tx][i];
        s1[i] += red[t * tx_count + tx][8 + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(out + tx * 8 + i, s0[i]);
      if (MODE != 2) atomicAdd(out + C + tx * 8 + i, s1[i]);
    }
  }
}

template <int MODE>
int launch_colreduce(const void* a, const void* b, const void* y, const float* saved, float* out,
                     long long rows, int C, int act, cudaStream_t st) {
  if (C % 8 || C > 2048) return set_error("column reduction needs C %% 8 == 0 and C <= 2048 (C=%d)", C);
  const int tx_count = next_pow2(C / 8) > kThreads ? kThreads : next_pow2(C / 8);
  const int ty_count = kThreads / tx_count;
  long long blocks = (rows + ty_count * 8 - 1) / (ty_count * 8);  // >= 8 rows per thread
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const long long rpb = (rows + blocks - 1) / blocks;
  blocks = (rows + rpb - 1) / rpb;
  colreduce_kernel<MODE><<<(int)blocks, kThreads, 0, st>>>(a, b, y, saved, out, rows, C, tx_count, act, rpb);
  return check_launch("colreduce_kernel");
}

// fp32 column sums (fc bias gradient from fp32 dlogits): small, one thread per column
__global__ void colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows,
                                  int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (long long r = 0; r < rows; ++r) s += x[r * C + c];
  out[c] = accumulate ? out[c] + s : s;
}

// ----------------------------------------------------------------------------- BN finalize
__global__ void bn_finalize_kernel(float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, float* __restrict__ ss,
                                   float* __restrict__ saved, long long rows, int C, float eps,
                                   float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv_n = 1.0f / (float)rows;
  const float mean = stats[c] * inv_n;
  float var = stats[C + c] * inv_n - mean * mean;
  var = fmaxf(var, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  ss[c] = sc;
  ss[C + c] = beta[c] - mean * sc;
  saved[c] = mean;
  saved[C + c] = rstd;
  if (rmean) {
    const float unbiased = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * unbiased;
  }
  stats[c] = 0.f;
  stats[C + c] = 0.f;
}

// ----------------------------------------------------------------------------- BN apply
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const void* __restrict__ y, const float* __restrict__ ss, const void* __restrict__ res,
                const float* __restrict__ rss, void* __restrict__ out, long long nvec, int C, int act) {
  const int vpr = C >> 3;
  for (long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec;
       vi += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(vi % vpr) * 8;
    float f[8];
    unpack8(ldg8(y, vi), f);
    const float4 sa = *reinterpret_cast<const float4*>(ss + c0), sb = *reinterpret_cast<const float4*>(ss + c0 + 4);
    const float4 ha = *reinterpret_cast<const float4*>(ss + C + c0), hb = *reinterpret_cast<const float4*>(ss + C + c0 + 4);
    const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
    const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f[i] * sc[i] + sh[i];
    if (res) {
      float r[8];
      unpack8(ldg8(res, vi), r);
      if (rss) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = r[i] * rss[c0 + i] + rss[C + c0 + i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += r[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = act_apply(f[i], act);
    stg8(out, vi, pack8(f));
  }
}

// ----------------------------------------------------------------------------- BN backward apply
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const void* __restrict__ dout, const void* __restrict__ out, const void* __restrict__ y,
                    const float* __restrict__ saved, const float* __restrict__ gamma,
                    const float* __restrict__ sums, void* __restrict__ dy, void* __restrict__ dres,
                    long long nvec, int C, int act, float inv_rows) {
  const int vpr = C >> 3;
  for (long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec;
       vi += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(vi % vpr) * 8;
    float g[8], fy[8];
    unpack8(ldg8(dout, vi), g);
    unpack8(ldg8(y, vi), fy);
    if (act != 0) {
      float fo[8];
      unpack8(ldg8(out, vi), fo);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] *= act_grad(fo[i], act);
    }
    if (dres) stg8(dres, vi, pack8(g));
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      const float mean = saved[c], rstd = saved[C + c];
      const float xhat = (fy[i] - mean) * rstd;
      d[i] = gamma[c] * rstd * (g[i] - sums[c] * inv_rows - xhat * sums[C + c] * inv_rows);
    }
    stg8(dy, vi, pack8(d));
  }
}

__global__ void bn_param_grad_kernel(const float* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (accumulate) {
    dbeta[c] += sums[c];
    dgamma[c] += sums[C + c];
  } else {
    dbeta[c] = sums[c];
    dgamma[c] = sums[C + c];
  }
}

// ----------------------------------------------------------------------------- small elementwise
__global__ void add_bf16_kernel(void* __restrict__ a, const void* __restrict__ b, long long nvec) {
  for (long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec;
       vi += (long long)gridDim.x * blockDim.x) {
    float fa[8], fb[8];
    unpack8(ldg8(a, vi), fa);
    unpack8(ldg8(b, vi), fb);
#pragma unroll
    for (int i = 0; i < 8; ++i) fa[i] += fb[i];
    stg8(a, vi, pack8(fa));
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits,
                                       long long n, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float s = accumulate ? out[i] : 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(long long)k * n + i];
    out[i] = s;
  }
}

__global__ void prep_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int K, int C,
                                        int R, int S, int kpad) {
  const long long total = (long long)K * kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i / kpad), j = (int)(i % kpad);
    float v = 0.f;
    if (j < R * S * C) {
      const int tap = j / C, c = j % C;
      v = w[((long long)k * C + c) * (R * S) + tap];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}

__global__ void finish_conv_wgrad_kernel(const float* __restrict__ partial, float* __restrict__ grad, int splits,
                                         int K, int C, int R, int S, int kpad, int accumulate) {
  const long long total = (long long)K * C * R * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % (R * S));
    const long long kc = i / (R * S);
    const int c = (int)(kc % C), k = (int)(kc / C);
    const long long src = (long long)k * kpad + (long long)tap * C + c;
    float s = accumulate ? grad[i] : 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[(long long)sp * K * kpad + src];
    grad[i] = s;
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int H,
                                    int W) {
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % ((long long)H * W);
    const long long n = i / ((long long)H * W);
    for (int c = 0; c < C; ++c) y[i * C + c] = __float2bfloat16_rn(x[(n * C + c) * (long long)H * W + hw]);
  }
}

// one thread per (output pixel, 8-column vector) of the im2col matrix
__global__ void stem_im2col_kernel(const float* __restrict__ x, void* __restrict__ cols, int N, int C, int H, int W,
                                   int R, int S, int stride, int pad, int P, int Q, int kpad) {
  const int vpr = kpad >> 3;
  const long long total = (long long)N * P * Q * vpr;
  const int RSC = R * S * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long m = i / vpr;
    const int q = (int)(m % Q);
    const int p = (int)((m / Q) % P);
    const int n = (int)(m / ((long long)P * Q));
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = v * 8 + e;
      float val = 0.f;
      if (j < RSC) {
        const int tap = j / C, c = j - tap * C;
        const int r = tap / S, s = tap - r * S;
        const int h = p * stride - pad + r, w = q * stride - pad + s;
        if (h >= 0 && h < H && w >= 0 && w < W) val = __ldg(x + (((long long)n * C + c) * H + h) * W + w);
      }
      f[e] = val;
    }
    stg8(cols, i, pack8(f));
  }
}

__global__ void zero_upsample2_kernel(const void* __restrict__ dy, void* __restrict__ u, int N, int P, int Q, int H,
                                      int W, int C) {
  const int vpr = C >> 3;
  const long long total = (long long)N * H * W * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long pix = i / vpr;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    V8 val;
#pragma unroll
    for (int k = 0; k < 4; ++k) val.h[k] = __floats2bfloat162_rn(0.f, 0.f);
    if (!(h & 1) && !(w & 1) && (h >> 1) < P && (w >> 1) < Q)
      val = ldg8(dy, ((n * P + (h >> 1)) * Q + (w >> 1)) * vpr + v);
    stg8(u, i, val);
  }
}

__global__ void add_strided2_kernel(void* __restrict__ dx, const void* __restrict__ dd, int N, int P, int Q, int H,
                                    int W, int C) {
  const int vpr = C >> 3;
  const long long total = (long long)N * P * Q * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long pix = i / vpr;
    const int q = (int)(pix % Q);
    const int p = (int)((pix / Q) % P);
    const long long n = pix / ((long long)Q * P);
    const long long di = ((n * H + 2 * p) * W + 2 * q) * vpr + v;
    float a[8], b[8];
    unpack8(ldg8(dx, di), a);
    unpack8(ldg8(dd, i), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    stg8(dx, di, pack8(a));
  }
}

// ----------------------------------------------------------------------------- pooling
__global__ void maxpool_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, uint8_t* __restrict__ amax,
                                   int N, int H, int W, int C, int P, int Q) {
  const int vpr = C >> 3;
  const long long total = (long long)N * P * Q * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long pix = i / vpr;
    const int q = (int)(pix % Q);
    const int p = (int)((pix / Q) % P);
    const long long n = pix / ((long long)Q * P);
    float best[8];
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
    bool first = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * p - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * q - 1 + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(ldg8(x, ((n * H + h) * W + w) * vpr + v), f);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (first || f[k] > best[k]) { best[k] = f[k]; arg[k] = r * 3 + s; }
        first = false;
      }
    }
    stg8(y, i, pack8(best));
    uint64_t packed = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) packed |= (uint64_t)(arg[k] & 0xff) << (8 * k);
    reinterpret_cast<uint64_t*>(amax)[i] = packed;
  }
}

__global__ void maxpool_bwd_kernel(const void* __restrict__ dy, const uint8_t* __restrict__ amax,
                                   void* __restrict__ dx, int N, int H, int W, int C, int P, int Q) {
  const int vpr = C >> 3;
  const long long total = (long long)N * H * W * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long pix = i / vpr;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // output windows (p, q) that contain (h, w): 2p-1 <= h <= 2p+1
    const int p_lo = h >> 1, p_hi = (h + 1) >> 1;
    const int q_lo = w >> 1, q_hi = (w + 1) >> 1;
    for (int p = p_lo; p <= p_hi; ++p) {
      if (p >= P) continue;
      const int r = h - (2 * p - 1);
      for (int q = q_lo; q <= q_hi; ++q) {
        if (q >= Q) continue;
        const int s = w - (2 * q - 1);
        const long long oi = ((n * P + p) * Q + q) * vpr + v;
        const uint64_t packed = reinterpret_cast<const uint64_t*>(amax)[oi];
        float g[8];
        unpack8(ldg8(dy, oi), g);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if ((int)((packed >> (8 * k)) & 0xff) == r * 3 + s) acc[k] += g[k];
      }
    }
    stg8(dx, i, pack8(acc));
  }
}

__global__ void avgpool_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, int N, int HW, int C) {
  const int vpr = C >> 3;
  const long long total = (long long)N * vpr;
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long n = i / vpr;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int t = 0; t < HW; ++t) {
      float f[8];
      unpack8(ldg8(x, (n * HW + t) * vpr + v), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv;
    stg8(y, i, pack8(acc));
  }
}

__global__ void avgpool_bwd_kernel(const void* __restrict__ dy, void* __restrict__ dx, int N, int HW, int C) {
  const int vpr = C >> 3;
  const long long total = (long long)N * HW * vpr;
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long n = i / ((long long)HW * vpr);
    float f[8];
    unpack8(ldg8(dy, n * vpr + v), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= inv;
    stg8(dx, i, pack8(f));
  }
}

}  // namespace
}  // namespace saicv

using namespace saicv;

#define ST ((cudaStream_t)stream)

extern "C" {

int saicv_version(void) { return 100; }
long long saicv_launch_count(void) { return saicv::launch_count(); }
const char* saicv_last_error(void) { return saicv::last_error(); }

int saicv_bn_stats(const void* y, float* stats, long long rows, int c, void* stream) {
  cudaMemsetAsync(stats, 0, sizeof(float) * 2 * c, ST);
  return launch_colreduce<0>(y, nullptr, nullptr, nullptr, stats, rows, c, 0, ST);
}

int saicv_bn_finalize(float* stats, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float* scale_shift, float* saved, long long rows, int c,
                      float eps, float momentum, void* stream) {
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, ST>>>(stats, gamma, beta, running_mean, running_var, scale_shift,
                                                       saved, rows, c, eps, momentum);
  return check_launch("bn_finalize_kernel");
}

int saicv_bn_apply(const void* y, const float* scale_shift, const void* res, const float* res_scale_shift,
                   void* out, long long rows, int c, int act, void* stream) {
  if (c % 8) return set_error("saicv_bn_apply: C %% 8 != 0");
  const long long nvec = rows * (c / 8);
  bn_apply_kernel<<<grid_for(nvec), kThreads, 0, ST>>>(y, scale_shift, res, res_scale_shift, out, nvec, c, act);
  return check_launch("bn_apply_kernel");
}

int saicv_bn_bwd_reduce(const void* dout, const void* out, const void* y, const float* saved, float* sums,
                        long long rows, int c, int act, void* stream) {
  if (act != 0 && out == nullptr) return set_error("saicv_bn_bwd_reduce: activated output required when act != 0");
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, ST);
  return launch_colreduce<1>(dout, out, y, saved, sums, rows, c, act, ST);
}

int saicv_bn_bwd_apply(const void* dout, const void* out, const void* y, const float* saved, const float* gamma,
                       float* sums, void* dy, void* dres, float* dgamma, float* dbeta, long long rows, int c,
                       int act, int accumulate, void* stream) {
  if (c % 8) return set_error("saicv_bn_bwd_apply: C %% 8 != 0");
  const long long nvec = rows * (c / 8);
  bn_bwd_apply_kernel<<<grid_for(nvec), kThreads, 0, ST>>>(dout, out, y, saved, gamma, sums, dy, dres, nvec, c, act,
                                                           1.0f / (float)rows);
  if (int e = check_launch("bn_bwd_apply_kernel")) return e;
  bn_param_grad_kernel<<<(c + 127) / 128, 128, 0, ST>>>(sums, dgamma, dbeta, c, accumulate);
  return check_launch("bn_param_grad_kernel");
}

int saicv_add_bf16(void* a, const void* b, long long n, void* stream) {
  if (n % 8) return set_error("saicv_add_bf16: n %% 8 != 0");
  add_bf16_kernel<<<grid_for(n / 8), kThreads, 0, ST>>>(a, b, n / 8);
  return check_launch("add_bf16_kernel");
}

int saicv_cast_bf16(const float* src, void* dst, long long n, void* stream) {
  cast_bf16_kernel<<<grid_for(n), kThreads, 0, ST>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  return check_launch("cast_bf16_kernel");
}

int saicv_reduce_partials(const float* partial, float* out, int splits, long long n, int accumulate,
                          void* stream) {
  reduce_partials_kernel<<<grid_for(n), kThreads, 0, ST>>>(partial, out, splits, n, accumulate);
  return check_launch("reduce_partials_kernel");
}

int saicv_prep_conv_weight(const float* w, void* w_bf16, int k, int c, int r, int s, int kpad, void* stream) {
  if (kpad % 8 || kpad < r * s * c) return set_error("saicv_prep_conv_weight: bad kpad %d", kpad);
  prep_conv_weight_kernel<<<grid_for((long long)k * kpad), kThreads, 0, ST>>>(
      w, reinterpret_cast<__nv_bfloat16*>(w_bf16), k, c, r, s, kpad);
  return check_launch("prep_conv_weight_kernel");
}

int saicv_finish_conv_wgrad(const float* partial, float* grad, int splits, int k, int c, int r, int s, int kpad,
                            int accumulate, void* stream) {
  finish_conv_wgrad_kernel<<<grid_for((long long)k * c * r * s), kThreads, 0, ST>>>(partial, grad, splits, k, c, r,
                                                                                     s, kpad, accumulate);
  return check_launch("finish_conv_wgrad_kernel");
}

int saicv_nchw_to_nhwc_bf16(const float* x, void* y, int n, int c, int h, int w, void* stream) {
  nchw_to_nhwc_kernel<<<grid_for((long long)n * h * w), kThreads, 0, ST>>>(x, reinterpret_cast<__nv_bfloat16*>(y),
                                                                           n, c, h, w);
  return check_launch("nchw_to_nhwc_kernel");
}

int saicv_stem_im2col(const float* x, void* cols, int n, int c, int h, int w, int r, int s, int stride, int pad,
                      int kpad, void* stream) {
  if (kpad % 8 || kpad < r * s * c) return set_error("saicv_stem_im2col: bad kpad %d", kpad);
  const int P = (h + 2 * pad - r) / stride + 1, Q = (w + 2 * pad - s) / stride + 1;
  stem_im2col_kernel<<<grid_for((long long)n * P * Q * (kpad / 8)), kThreads, 0, ST>>>(x, cols, n, c, h, w, r, s,
                                                                                      stride, pad, P, Q, kpad);
  return check_launch("stem_im2col_kernel");
}

int saicv_zero_upsample2(const void* dy, void* u, int n, int p, int q, int h, int w, int c, void* stream) {
  if (c % 8) return set_error("saicv_zero_upsample2: C %% 8 != 0");
  zero_upsample2_kernel<<<grid_for((long long)n * h * w * (c / 8)), kThreads, 0, ST>>>(dy, u, n, p, q, h, w, c);
  return check_launch("zero_upsample2_kernel");
}

int saicv_add_strided2(void* dx, const void* dd, int n, int p, int q, int h, int w, int c, void* stream) {
  if (c % 8) return set_error("saicv_add_strided2: C %% 8 != 0");
  add_strided2_kernel<<<grid_for((long long)n * p * q * (c / 8)), kThreads, 0, ST>>>(dx, dd, n, p, q, h, w, c);
  return check_launch("add_strided2_kernel");
}

int saicv_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* argmax, int n, int h, int w, int c, void* stream) {
  if (c % 8) return set_error("saicv_maxpool3x3s2_fwd: C %% 8 != 0");
  const int P = (h + 2 - 3) / 2 + 1, Q = (w + 2 - 3) / 2 + 1;
  maxpool_fwd_kernel<<<grid_for((long long)n * P * Q * (c / 8)), kThreads, 0, ST>>>(x, y, argmax, n, h, w, c, P, Q);
  return check_launch("maxpool_fwd_kernel");
}

int saicv_maxpool3x3s2_bwd(const void* dy, const uint8_t* argmax, void* dx, int n, int h, int w, int c,
                           void* stream) {
  if (c % 8) return set_error("saicv_maxpool3x3s2_bwd: C %% 8 != 0");
  const int P = (h + 2 - 3) / 2 + 1, Q = (w + 2 - 3) / 2 + 1;
  maxpool_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8)), kThreads, 0, ST>>>(dy, argmax, dx, n, h, w, c, P, Q);
  return check_launch("maxpool_bwd_kernel");
}

int saicv_avgpool_fwd(const void* x, void* y, int n, int hw, int c, void* stream) {
  if (c % 8) return set_error("saicv_avgpool_fwd: C %% 8 != 0");
  avgpool_fwd_kernel<<<grid_for((long long)n * (c / 8), 128), 128, 0, ST>>>(x, y, n, hw, c);
  return check_launch("avgpool_fwd_kernel");
}

int saicv_avgpool_bwd(const void* dy, void* dx, int n, int hw, int c, void* stream) {
  if (c % 8) return set_error("saicv_avgpool_bwd: C %% 8 != 0");
  avgpool_bwd_kernel<<<grid_for((long long)n * hw * (c / 8)), kThreads, 0, ST>>>(dy, dx, n, hw, c);
  return check_launch("avgpool_bwd_kernel");
}

int saicv_colsum(const void* x, float* out, long long rows, int c, int accumulate, int is_f32, void* stream) {
  if (is_f32) {
    colsum_f32_kernel<<<(c + 127) / 128, 128, 0, ST>>>(reinterpret_cast<const float*>(x), out, rows, c, accumulate);
    return check_launch("colsum_f32_kernel");
  }
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * c, ST);
  return launch_colreduce<2>(x, nullptr, nullptr, nullptr, out, rows, c, 0, ST);
}

}  // extern "C"
