// Fused multi-head attention on tcgen05 / TMEM / TMA for sm_100a: forward and the two backward phases.
//
//   S = Q K^T * scale (+ key-padding mask);  P = softmax(S);  O = P V          (never materialised in HBM)
//
// One kernel family, templated on the score width DQK (q/k row length, multiple of 16, <= 256) and the
// value width DV (multiple of 16, <= 128); sequence lengths are runtime and tiled (128-row tiles, key blocks
// of 128 in the forward / 64 in the backward), so the same code serves
//   ViT            hd 64, L = 197                     (classification/backbones/vit.py:62-80)
//   SAM encoder    hd 64/80, 14x14 windows (L = 196) and global attention (L = 4096); the decomposed
//                  rel-pos bias enters as extra score columns  [q | rel_h | rel_w] . [k | onehot_h | onehot_w]
//                  so DQK = hd + bias columns            (segment_anything/image_encoder.py:82-144,167-184)
//   DETR           hd 32, self attention with key-padding mask and cross attention Lq != Lk
//                                                        (detection/models/detr.py:54-56,103-109)
//
// CTA = 192 threads, persistent over (batch*head, row tile) work items, up to two CTAs per SM so that the
// softmax of one overlaps the MMAs of the other:
//   warp 0      TMA producer (one lane): Q tile once per item, K / V blocks through a 2-stage ring
//   warp 1      MMA issuer (one lane) + TMEM allocator: S = Q K^T, O_blk = P V via tcgen05.mma
//   warps 2..5  softmax: thread = one query row (TMEM lane); tcgen05.ld S -> exp2 -> bf16 P into 128B-swizzled
//               shared memory (the A operand of P V); O accumulated in registers with the online-softmax rescale
#pragma once
#include "dropout_hash.cuh"
#include "ptx.cuh"

namespace saicv {

constexpr int kAttnThreads = 192;

struct AttnDrop {            // attention-probability dropout (thresh == 0: off), see dropout_hash.cuh
  uint32_t thresh, seed_lo, seed_hi;
  float scale;                // 1 / (1 - p)
  const unsigned long long* seed_base;   // device word added to the seed (nullable): fresh masks per CUDA-graph replay
};
// the effective seed words of this launch (one uniform global load when a device base is given)
__device__ __forceinline__ AttnDrop resolve_drop(const AttnDrop& d) {
  AttnDrop r = d;
  if (d.thresh != 0 && d.seed_base != nullptr) {
    const unsigned long long s = (((unsigned long long)d.seed_hi << 32) | d.seed_lo) + *d.seed_base;
    r.seed_lo = (uint32_t)s;
    r.seed_hi = (uint32_t)(s >> 32);
  }
  return r;
}

struct AttnParams {
  AttnDrop drop;
  int B, H, Lq, Lk;
  int num_q_tiles;          // ceil(Lq / 128)
  float scale_log2;         // softmax scale * log2(e)
  float scale;
  __nv_bfloat16* out;       // [b][row][h][DV] through strides (elements)
  long long o_sb, o_sh, o_sl;
  float* lse;               // [B][H][Lq]  log2 domain: m + log2(sum)
  const uint32_t* mask_bits;  // [B][mask_words] bit k%32 of word k/32 set = key k masked out; may be null
  int mask_words;
  int* error_flag;
};

__device__ __forceinline__ uint32_t idesc_with_n(uint32_t idesc_base, int n) {
  return (idesc_base & ~(0x3Fu << 17)) | (static_cast<uint32_t>(n >> 3) << 17);
}

// issue the K-loop of one score GEMM D[tmem] = A[128 x D] * B[n x D]^T, both operands K-major in 64-column chunks
template <int D>
__device__ __forceinline__ void issue_scores(uint32_t d_tmem, uint32_t a_addr, uint32_t a_chunk_bytes, uint32_t b_addr,
                                             uint32_t b_chunk_bytes, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < D / 16; ++k) {
    const uint64_t da = make_smem_desc(a_addr + (k >> 2) * a_chunk_bytes + (k & 3) * 32, 16, 1024);
    const uint64_t db = make_smem_desc(b_addr + (k >> 2) * b_chunk_bytes + (k & 3) * 32, 16, 1024);
    umma_bf16(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
  }
}
// D[tmem] (+)= A[128 x nk] * B[nk x N]: A K-major in 64-column chunks (written by the softmax threads),
// B MN-major (rows = reduction index, N contiguous) in 64-column chunks `b_chunk_bytes` apart
__device__ __forceinline__ void issue_pv(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t b_chunk_bytes, int nk,
                                         uint32_t idesc, bool accumulate) {
  for (int k = 0; k < nk / 16; ++k) {
    const uint64_t da = make_smem_desc(a_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
    const uint64_t db = make_smem_desc(b_addr + k * 2048, b_chunk_bytes, 1024);
    umma_bf16(d_tmem, da, db, idesc, (accumulate || k > 0) ? 1u : 0u);
  }
}

// running maximum over one 32-column chunk of raw scores; `dead` bit i set = column i is padding / masked
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], uint32_t dead, float mx) {
  if (dead == 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, ((dead >> i) & 1u) ? -INFINITY : __uint_as_float(v[i]));
  }
  return mx;
}
// e = exp2(s * sl - msl) for one 32-column chunk (0 where dead), packed to bf16 into the 128B-swizzled A tile:
// 16-byte pieces piece0 .. piece0+3 of this thread's 128-byte row (row & 7 == sw).  Returns rs + sum(e).
template <bool DROP>
__device__ __forceinline__ float chunk_exp_store(const uint32_t (&v)[32], uint32_t dead, float sl, float msl, float rs,
                                                 uint8_t* row_base, int piece0, int sw, const AttnDrop& dr, uint32_t drow,
                                                 uint32_t dcol0) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {   // 8 columns -> one 16-byte piece
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(v[8 * t + i]), sl, -msl));
    if (dead != 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = ((dead >> (8 * t + i)) & 1u) ? 0.f : e[i];
    }
    rs += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
    if (DROP) {   // the row sum is that of the undropped probabilities; P V uses the dropped ones
#pragma unroll
      for (int i = 0; i < 8; ++i)
        e[i] = dropout_hash(dr.seed_lo, dr.seed_hi, drow, dcol0 + 8 * t + i) >= dr.thresh ? e[i] * dr.scale : 0.f;
    }
    *reinterpret_cast<uint4*>(row_base + (((piece0 + t) ^ sw) << 4)) =
        make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
  }
  return rs;
}

template <int DQK, int DV, int MINB>
struct AttnFwdCfg {
  static constexpr int BN = 128;
  static constexpr int NCQ = (DQK + 63) / 64, NCV = (DV + 63) / 64;
  static constexpr int QBYTES = NCQ * 16384, KBYTES = NCQ * BN * 128, VBYTES = NCV * BN * 128, PBYTES = (BN / 64) * 16384;
  static constexpr int kStages = (QBYTES + 2 * (KBYTES + VBYTES) + PBYTES + 256 <= 232448) ? 2 : 1;   // K / V ring depth
  static constexpr int kSmemBytes = QBYTES + kStages * (KBYTES + VBYTES) + PBYTES + 256;
  static constexpr int OCOLS = ((DV + 31) & ~31);   // O accumulator columns
  static constexpr int kTmemCols = 256;             // S: columns 0..127, O: 128..128+OCOLS
};

// Forward.  Numerics: P = exp2((s - m_ref) * scale_log2) with a per-row reference m_ref that is only raised when a
// block's scores exceed it by more than 2^8 (lazy rescaling): O and the row sums then stay in TMEM across key
// blocks (accumulating MMAs) and are read once per work item; a raise multiplies both by 2^((m_old - m_new) * scale_log2)
// in TMEM (rare); the row sums live in registers and follow the same rescaling.
template <int DQK, int DV, int MINB, bool DROP>
__global__ void __launch_bounds__(kAttnThreads, MINB)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Cfg = AttnFwdCfg<DQK, DV, MINB>;
  constexpr int BN = Cfg::BN, NCQ = Cfg::NCQ, NCV = Cfg::NCV, kStages = Cfg::kStages, OCOLS = Cfg::OCOLS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::QBYTES;
  uint8_t* sV = sK + kStages * Cfg::KBYTES;
  uint8_t* sP = sV + kStages * Cfg::VBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::PBYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;              // [kStages]
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;     // S_j in TMEM
  uint64_t* s_empty = s_full + 1;           // S_j consumed by the softmax warps
  uint64_t* p_full = s_full + 2;            // P_j in shared memory
  uint64_t* pv_done = s_full + 3;           // P_j V_j has completed (once per key block)
  uint64_t* o_empty = s_full + 4;           // O / l of the work item read out (once per work item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) {  // 128B-swizzle atoms need a 1024-byte aligned base
    if (threadIdx.x == 0 && p.error_flag) *p.error_flag = 1;
    return;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  const int nkb = (p.Lk + BN - 1) / BN;
  const int total = p.B * p.H * p.num_q_tiles;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      uint32_t it = 0, g = 0;  // work items done, key blocks issued (ring position)
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int qt = w % p.num_q_tiles, bh = w / p.num_q_tiles, b = bh / p.H, h = bh % p.H;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, Cfg::QBYTES);
#pragma unroll
        for (int c = 0; c < NCQ; ++c) tma_load_4d(&tmQ, q_full, sQ + c * 16384, c * 64, qt * 128, h, b);
        for (int j = 0; j < nkb; ++j, ++g) {
          const int st = g % kStages;
          const uint32_t ph = (g / kStages) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], Cfg::KBYTES);
#pragma unroll
          for (int c = 0; c < NCQ; ++c)
            tma_load_4d(&tmK, &k_full[st], sK + st * Cfg::KBYTES + c * (BN * 128), c * 64, j * BN, h, b);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], Cfg::VBYTES);
#pragma unroll
          for (int c = 0; c < NCV; ++c)
            tma_load_4d(&tmV, &v_full[st], sV + st * Cfg::VBYTES + c * (BN * 128), c * 64, j * BN, h, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, BN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, DV, 0, 1);
      uint32_t it = 0, g = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        for (int j = 0; j < nkb; ++j, ++g) {
          const int st = g % kStages;
          const uint32_t ph = (g / kStages) & 1;
          const int nvalid = min(BN, p.Lk - j * BN);
          const int nj = (nvalid + 15) & ~15;
          // ---- S_j = Q K_j^T
          mbar_wait(&k_full[st], ph);
          mbar_wait(s_empty, (g & 1) ^ 1);
          tc_fence_after();
          issue_scores<DQK>(tmem_S, smem_u32(sQ), 16384, smem_u32(sK + st * Cfg::KBYTES), BN * 128, idesc_with_n(idesc_s, nj));
          umma_commit(&k_empty[st]);
          umma_commit(s_full);
          if (j == nkb - 1) umma_commit(q_empty);
          // ---- O (+)= P_j V_j
          mbar_wait(&v_full[st], ph);
          if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);   // the previous item's O has been read out
          mbar_wait(p_full, g & 1);
          tc_fence_after();
          issue_pv(tmem_O, smem_u32(sP), smem_u32(sV + st * Cfg::VBYTES), BN * 128, nj, idesc_o, j > 0);
          umma_commit(&v_empty[st]);
          umma_commit(pv_done);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================ softmax / output
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* const prow = sP + row * 128;
    const int sw = row & 7;
    const float thr_raw = 8.0f / p.scale_log2;        // raise the reference only past a factor 2^8
    const AttnDrop drop = DROP ? resolve_drop(p.drop) : p.drop;
    uint32_t g = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int qt = w % p.num_q_tiles, bh = w / p.num_q_tiles, b = bh / p.H, h = bh % p.H;
      float m_ref = -INFINITY, l = 0.f;               // reference in raw-score units; row sum relative to it
      for (int j = 0; j < nkb; ++j, ++g) {
        const int nvalid = min(BN, p.Lk - j * BN);
        const int nchunks = (nvalid + 31) >> 5;
        auto dead_bits = [&](int c) -> uint32_t {
          uint32_t dead = 0;
          const int rem = nvalid - c * 32;
          if (rem < 32) dead = 0xffffffffu << rem;
          if (p.mask_bits) dead |= __ldg(p.mask_bits + (long long)b * p.mask_words + ((j * BN) >> 5) + c);
          return dead;
        };
        mbar_wait(s_full, g & 1);
        if (g > 0) mbar_wait(pv_done, (g - 1) & 1);   // P_{g-1} V_{g-1} no longer reads the P tile (and O is quiescent)
        tc_fence_after();
        // Single pass over S_j.  The block is processed optimistically with the current reference; if some chunk
        // exceeds it by more than 2^8 the reference is raised (O in TMEM is rescaled, warp-collectively) and, unless
        // nothing of this block was written yet, the block is started over with the new reference.
        // (Measured alternatives on B200, ViT-B shape: all of S in registers + S_{j+1} issued ahead of P_j V_j: 226 us;
        //  8 softmax warps with a half-row max exchange: 228 us; this form: 179 us.)
        bool restart;
        float rs;
        do {
          restart = false;
          rs = 0.f;
          for (int c = 0; c < nchunks; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_S + lane_off + c * 32, v);
            tmem_ld_wait();
            const uint32_t dead = dead_bits(c);
            const float cm = chunk_max(v, dead, -INFINITY);
            const bool raise = cm > m_ref + thr_raw;           // (-inf + thr == -inf: the first valid score always raises)
            if (__any_sync(0xffffffffu, raise)) {
              float bm = cm;                                   // maximum of chunks c .. end of this block
              for (int c2 = c + 1; c2 < nchunks; ++c2) {
                uint32_t t[32];
                tmem_ld_32x32(tmem_S + lane_off + c2 * 32, t);
                tmem_ld_wait();
                bm = chunk_max(t, dead_bits(c2), bm);
              }
              const bool had_ref = m_ref != -INFINITY;
              const float m_new = raise ? bm : m_ref;
              const float corr = (raise && had_ref) ? ex2_approx((m_ref - m_new) * p.scale_log2) : 1.f;
              l *= corr;
              if (j > 0 && __any_sync(0xffffffffu, raise && had_ref)) {   // blocks < j sit in TMEM relative to the old reference
#pragma unroll 1
                for (int c2 = 0; c2 < OCOLS / 32; ++c2) {
                  uint32_t t[32];
                  tmem_ld_32x32(tmem_O + lane_off + c2 * 32, t);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * corr);
                  tmem_st_32x32(tmem_O + lane_off + c2 * 32, t);
                }
                tmem_st_wait();
              }
              m_ref = m_new;
              if (c > 0) {   // chunks 0 .. c-1 were written relative to the old reference
                restart = true;
                break;
              }
            }
            const float msl = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;
            rs = chunk_exp_store<DROP>(v, dead, p.scale_log2, msl, rs, prow + (c >> 1) * 16384, (c & 1) * 4, sw, drop,
                                 (uint32_t)(bh * p.Lq + qt * 128 + row), (uint32_t)(j * BN + c * 32));
          }
        } while (restart);
        l += rs;
        // S fully consumed: the MMA warp may overwrite it with the next block's scores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- epilogue: wait for the last P V, read O and l once, normalise, store the row and its log-sum-exp
      mbar_wait(pv_done, (g - 1) & 1);
      tc_fence_after();
      const float inv = l > 0.f ? __fdividef(1.f, l) : 0.f;
      const int qrow = qt * 128 + row;
      __nv_bfloat16* orow = p.out + (long long)b * p.o_sb + (long long)h * p.o_sh + (long long)qrow * p.o_sl;
#pragma unroll
      for (int c = 0; c < (DV + 31) / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_O + lane_off + c * 32, v);
        tmem_ld_wait();
        if (qrow < p.Lq) {
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            if (c * 32 + i < DV)
              *reinterpret_cast<uint4*>(orow + c * 32 + i) =
                  make_uint4(pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv),
                             pack_bf16x2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv),
                             pack_bf16x2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv),
                             pack_bf16x2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (qrow < p.Lq) p.lse[((long long)b * p.H + h) * p.Lq + qrow] = l > 0.f ? m_ref * p.scale_log2 + log2f(l) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ backward
struct AttnBwdParams {
  AttnDrop drop;
  int B, H, Lq, Lk;
  int num_tiles;            // row tiles per (b, h): ceil(Lq / 128) in the dQ phase, ceil(Lk / 128) in the dK/dV phase
  float scale_log2, scale;
  const float* lse;         // [B][H][Lq]
  const float* delta;       // [B][H][Lq]  D = rowsum(dO * O)
  const uint32_t* mask_bits;
  int mask_words;
  __nv_bfloat16* d_out0;    // dQ (phase A) or dK (phase B)
  long long s0b, s0h, s0l;
  __nv_bfloat16* d_out1;    // dV (phase B)
  long long s1b, s1h, s1l;
  int dk_cols;              // phase B: leading columns of dK that are computed / stored (<= DQK)
  int* error_flag;
};

// Phase A (ROWS_ARE_KEYS = false): row tile = 128 queries, column blocks = 64 keys:
//     S = Q K^T, dP = dO V^T, dS = P (dP - D) scale, dQ += dS K              (dQ accumulates in TMEM)
// Phase B (ROWS_ARE_KEYS = true): row tile = 128 keys, column blocks = 64 queries:
//     S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q                  (dK, dV accumulate in TMEM)
// tmR0 / tmR1: the row-tile operands (A: Q, dO; B: K, V); tmC0 / tmC1: the column-block operands (A: K, V; B: Q, dO).
template <int DQK, int DV, bool ROWS_ARE_KEYS, int TMEM_COLS, int MINB>
struct AttnBwdCfg {
  static constexpr int BN = 64;
  static constexpr int NCQ = (DQK + 63) / 64, NCV = (DV + 63) / 64;
  static constexpr int kStages = 2;
  static constexpr int R0BYTES = NCQ * 16384, R1BYTES = NCV * 16384;       // row tiles (128 rows)
  static constexpr int C0BYTES = NCQ * 8192, C1BYTES = NCV * 8192;         // column blocks (64 rows)
  static constexpr int ABYTES = 16384;                                       // one [128 x 64] bf16 A tile
  static constexpr int NA = ROWS_ARE_KEYS ? 2 : 1;
  static constexpr int kSmemBytes = R0BYTES + R1BYTES + kStages * (C0BYTES + C1BYTES) + NA * ABYTES + 1024 /*lse, delta*/ + 256;
};

template <int DQK, int DV, bool ROWS_ARE_KEYS, int TMEM_COLS, int MINB, bool DROP>
__global__ void __launch_bounds__(kAttnThreads, MINB)
attn_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                      const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1,
                      const AttnBwdParams p) {
  using Cfg = AttnBwdCfg<DQK, DV, ROWS_ARE_KEYS, TMEM_COLS, MINB>;
  constexpr int BN = Cfg::BN, NCQ = Cfg::NCQ, NCV = Cfg::NCV, kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sR0 = smem;
  uint8_t* sR1 = sR0 + Cfg::R0BYTES;
  uint8_t* sC0 = sR1 + Cfg::R1BYTES;
  uint8_t* sC1 = sC0 + kStages * Cfg::C0BYTES;
  uint8_t* sA0 = sC1 + kStages * Cfg::C1BYTES;      // dS (phase A) / P^T (phase B)
  uint8_t* sA1 = sA0 + Cfg::ABYTES;                 // dS^T (phase B)
  float* sLse = reinterpret_cast<float*>(sA0 + Cfg::NA * Cfg::ABYTES);   // [2][64] phase B: lse / delta of the query block
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLse) + 1024);
  uint64_t* r_full = bars + 0;
  uint64_t* r_empty = bars + 1;
  uint64_t* c_full = bars + 2;            // [kStages]
  uint64_t* c_empty = c_full + kStages;
  uint64_t* sp_full = c_empty + kStages;  // S and dP ready in TMEM
  uint64_t* sp_empty = sp_full + 1;       // S / dP consumed
  uint64_t* a_full = sp_full + 2;         // dS (/ P^T, dS^T) written to shared memory
  uint64_t* a_empty = sp_full + 3;        // accumulate MMAs that read the A tiles have completed
  uint64_t* acc_full = sp_full + 4;       // all accumulate MMAs of the work item have completed
  uint64_t* acc_empty = sp_full + 5;      // accumulators read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sp_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0 && p.error_flag) *p.error_flag = 1;
    return;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmR0);
    tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmC0);
    tma_prefetch_desc(&tmC1);
    mbar_init(r_full, 1);
    mbar_init(r_empty, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(sp_empty, 4);
    mbar_init(a_full, 4);
    mbar_init(a_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64;
  const uint32_t tmem_acc0 = tmem_base + 128;                            // dQ (A) / dV (B)
  const uint32_t tmem_acc1 = tmem_base + 128 + ((DV + 31) & ~31);        // dK (B)

  const int Lrow = ROWS_ARE_KEYS ? p.Lk : p.Lq;     // length along the row tiles
  const int Lcol = ROWS_ARE_KEYS ? p.Lq : p.Lk;     // length along the column blocks
  const int ncb = (Lcol + BN - 1) / BN;
  const int total = p.B * p.H * p.num_tiles;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      uint32_t it = 0, g = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int rt = w % p.num_tiles, bh = w / p.num_tiles, b = bh / p.H, h = bh % p.H;
        mbar_wait(r_empty, (it & 1) ^ 1);
        mbar_expect_tx(r_full, Cfg::R0BYTES + Cfg::R1BYTES);
#pragma unroll
        for (int c = 0; c < NCQ; ++c) tma_load_4d(&tmR0, r_full, sR0 + c * 16384, c * 64, rt * 128, h, b);
#pragma unroll
        for (int c = 0; c < NCV; ++c) tma_load_4d(&tmR1, r_full, sR1 + c * 16384, c * 64, rt * 128, h, b);
        for (int j = 0; j < ncb; ++j, ++g) {
          const int st = g % kStages;
          const uint32_t ph = (g / kStages) & 1;
          mbar_wait(&c_empty[st], ph ^ 1);
          mbar_expect_tx(&c_full[st], Cfg::C0BYTES + Cfg::C1BYTES);
#pragma unroll
          for (int c = 0; c < NCQ; ++c)
            tma_load_4d(&tmC0, &c_full[st], sC0 + st * Cfg::C0BYTES + c * 8192, c * 64, j * BN, h, b);
#pragma unroll
          for (int c = 0; c < NCV; ++c)
            tma_load_4d(&tmC1, &c_full[st], sC1 + st * Cfg::C1BYTES + c * 8192, c * 64, j * BN, h, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, BN, 0, 0);
      const uint32_t idesc_dq = make_idesc_bf16(128, DQK, 0, 1);    // phase A: dQ += dS K     (N = DQK)
      const uint32_t idesc_dv = make_idesc_bf16(128, DV, 0, 1);     // phase B: dV += P^T dO   (N = DV)
      uint32_t it = 0, gs = 0, ga = 0;   // work items, score blocks issued, accumulate blocks issued
      const uint32_t idesc_dk = make_idesc_bf16(128, p.dk_cols, 0, 1);   // phase B: dK += dS^T Q (N = dk_cols)
      // scores + dP of column block j: issued as soon as the elementwise warps have copied block j-1 out of TMEM, i.e.
      // ahead of the accumulate GEMMs of block j-1, so the tensor core works while the exponentials are evaluated
      auto issue_sp = [&](int j) {
        const int st = gs % kStages;
        const uint32_t ph = (gs / kStages) & 1;
        const int nj = (min(BN, Lcol - j * BN) + 15) & ~15;
        mbar_wait(&c_full[st], ph);
        mbar_wait(sp_empty, (gs & 1) ^ 1);
        tc_fence_after();
        const uint32_t c0 = smem_u32(sC0 + st * Cfg::C0BYTES), c1 = smem_u32(sC1 + st * Cfg::C1BYTES);
        issue_scores<DQK>(tmem_S, smem_u32(sR0), 16384, c0, 8192, idesc_with_n(idesc_s, nj));
        issue_scores<DV>(tmem_dP, smem_u32(sR1), 16384, c1, 8192, idesc_with_n(idesc_s, nj));
        umma_commit(sp_full);
        ++gs;
      };
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        mbar_wait(r_full, it & 1);
        issue_sp(0);
        for (int j = 0; j < ncb; ++j, ++ga) {
          if (j + 1 < ncb) issue_sp(j + 1);
          const int st = ga % kStages;
          const int nj = (min(BN, Lcol - j * BN) + 15) & ~15;
          const uint32_t c0 = smem_u32(sC0 + st * Cfg::C0BYTES), c1 = smem_u32(sC1 + st * Cfg::C1BYTES);
          if (j == 0) mbar_wait(acc_empty, (it & 1) ^ 1);    // the previous item's accumulators have been read out
          mbar_wait(a_full, ga & 1);
          tc_fence_after();
          if (!ROWS_ARE_KEYS) {
            issue_pv(tmem_acc0, smem_u32(sA0), c0, 8192, nj, idesc_dq, j > 0);             // dQ += dS K_j
          } else {
            issue_pv(tmem_acc0, smem_u32(sA0), c1, 8192, nj, idesc_dv, j > 0);             // dV += P^T dO_i
            issue_pv(tmem_acc1, smem_u32(sA1), c0, 8192, nj, idesc_dk, j > 0);             // dK += dS^T Q_i
          }
          umma_commit(&c_empty[st]);
          umma_commit(a_empty);
        }
        umma_commit(r_empty);
        umma_commit(acc_full);
      }
    }
    __syncwarp();
  } else {
    // ================================================================ elementwise (thread = one row of the tile)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int sw = row & 7;
    const AttnDrop drop = DROP ? resolve_drop(p.drop) : p.drop;
    const int tid = threadIdx.x - 64;
    uint32_t it = 0, g = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int rt = w % p.num_tiles, bh = w / p.num_tiles, b = bh / p.H, h = bh % p.H;
      const int grow = rt * 128 + row;                 // global row index (query in phase A, key in phase B)
      const bool row_ok = grow < Lrow;
      float lse_r = 0.f, dl_r = 0.f;
      bool row_masked = false;
      if (!ROWS_ARE_KEYS) {
        if (row_ok) {
          lse_r = p.lse[(long long)bh * p.Lq + grow];
          dl_r = p.delta[(long long)bh * p.Lq + grow];
        }
      } else if (p.mask_bits && row_ok) {
        row_masked = (__ldg(p.mask_bits + (long long)b * p.mask_words + (grow >> 5)) >> (grow & 31)) & 1u;
      }
      for (int j = 0; j < ncb; ++j, ++g) {
        const int nvalid = min(BN, Lcol - j * BN);
        if (ROWS_ARE_KEYS) {
          // lse / delta of the 64 queries of this block -> shared memory (broadcast reads below)
          named_bar_sync(1, 128);   // previous block's readers are done
          if (tid < 64) {
            const int q = j * BN + tid;
            sLse[tid] = q < p.Lq ? p.lse[(long long)bh * p.Lq + q] : 0.f;
          } else {
            const int q = j * BN + tid - 64;
            sLse[tid] = q < p.Lq ? p.delta[(long long)bh * p.Lq + q] : 0.f;
          }
          named_bar_sync(1, 128);
        }
        mbar_wait(sp_full, g & 1);
        tc_fence_after();
        // S and dP of the block go to registers with one wait; TMEM is handed back immediately (the MMA warp then
        // computes the next block's scores while this block's exponentials run)
        const int nchunks = (nvalid + 31) >> 5;
        uint32_t sv[2][32], dv[2][32];
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nchunks) {
            tmem_ld_32x32(tmem_S + lane_off + c * 32, sv[c]);
            tmem_ld_32x32(tmem_dP + lane_off + c * 32, dv[c]);
          }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sp_empty);
        mbar_wait(a_empty, (g & 1) ^ 1);     // the accumulate MMAs of the previous block have read the A tiles
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c >= nchunks) continue;          // (nj <= 32: the accumulate GEMMs do not read these columns)
          uint32_t dead = 0;
          const int rem = nvalid - c * 32;
          if (rem < 32) dead = 0xffffffffu << rem;
          if (!ROWS_ARE_KEYS && p.mask_bits)
            dead |= __ldg(p.mask_bits + (long long)b * p.mask_words + ((j * BN) >> 5) + c);
          if (!row_ok || row_masked) dead = 0xffffffffu;
#pragma unroll
          for (int t = 0; t < 4; ++t) {        // 8 columns -> one 16-byte piece of P (/ P^T) and dS (/ dS^T)
            float pp[8], ds[8];
            if (ROWS_ARE_KEYS) {
              const float4 L0 = *reinterpret_cast<const float4*>(sLse + c * 32 + 8 * t);
              const float4 L1 = *reinterpret_cast<const float4*>(sLse + c * 32 + 8 * t + 4);
              const float4 D0 = *reinterpret_cast<const float4*>(sLse + 64 + c * 32 + 8 * t);
              const float4 D1 = *reinterpret_cast<const float4*>(sLse + 64 + c * 32 + 8 * t + 4);
              const float ls[8] = {L0.x, L0.y, L0.z, L0.w, L1.x, L1.y, L1.z, L1.w};
              const float dl[8] = {D0.x, D0.y, D0.z, D0.w, D1.x, D1.y, D1.z, D1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                pp[i] = ex2_approx(fmaf(__uint_as_float(sv[c][8 * t + i]), p.scale_log2, -ls[i]));
                ds[i] = pp[i] * (__uint_as_float(dv[c][8 * t + i]) - dl[i]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                pp[i] = ex2_approx(fmaf(__uint_as_float(sv[c][8 * t + i]), p.scale_log2, -lse_r));
                ds[i] = pp[i] * (__uint_as_float(dv[c][8 * t + i]) - dl_r);
              }
            }
            if (DROP) {
              // A~ = M A / (1 - p):  dS = A (M dP / (1 - p) - D);  dV uses A~ (phase B's pp tile)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int col = j * BN + c * 32 + 8 * t + i;
                const uint32_t qi = ROWS_ARE_KEYS ? (uint32_t)col : (uint32_t)grow;
                const uint32_t ki = ROWS_ARE_KEYS ? (uint32_t)grow : (uint32_t)col;
                const bool keep = dropout_hash(drop.seed_lo, drop.seed_hi, (uint32_t)bh * (uint32_t)p.Lq + qi, ki) >= drop.thresh;
                const float dpv = keep ? __uint_as_float(dv[c][8 * t + i]) * drop.scale : 0.f;
                float dlv;
                if (ROWS_ARE_KEYS) dlv = sLse[64 + c * 32 + 8 * t + i];
                else dlv = dl_r;
                ds[i] = pp[i] * (dpv - dlv);
                pp[i] = keep ? pp[i] * drop.scale : 0.f;
              }
            }
            if (dead != 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const bool dd = (dead >> (8 * t + i)) & 1u;
                pp[i] = dd ? 0.f : pp[i];
                ds[i] = dd ? 0.f : ds[i];
              }
            }
            const int piece = c * 4 + t;
            const uint4 dsv = make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]),
                                         pack_bf16x2(ds[6], ds[7]));
            if (ROWS_ARE_KEYS) {
              *reinterpret_cast<uint4*>(sA0 + row * 128 + ((piece ^ sw) << 4)) =
                  make_uint4(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]), pack_bf16x2(pp[4], pp[5]), pack_bf16x2(pp[6], pp[7]));
              *reinterpret_cast<uint4*>(sA1 + row * 128 + ((piece ^ sw) << 4)) = dsv;
            } else {
              *reinterpret_cast<uint4*>(sA0 + row * 128 + ((piece ^ sw) << 4)) = dsv;
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }
      // ---- read the accumulators out
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      if (!ROWS_ARE_KEYS) {
        __nv_bfloat16* orow = p.d_out0 + (long long)b * p.s0b + (long long)h * p.s0h + (long long)grow * p.s0l;
#pragma unroll
        for (int c = 0; c < (DQK + 31) / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_acc0 + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * p.scale);
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 8)
              if (c * 32 + i < DQK)
                *reinterpret_cast<uint4*>(orow + c * 32 + i) =
                    make_uint4(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                               pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                               pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                               pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
          }
        }
      } else {
        __nv_bfloat16* vrow = p.d_out1 + (long long)b * p.s1b + (long long)h * p.s1h + (long long)grow * p.s1l;
        __nv_bfloat16* krow = p.d_out0 + (long long)b * p.s0b + (long long)h * p.s0h + (long long)grow * p.s0l;
#pragma unroll
        for (int c = 0; c < (DV + 31) / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_acc0 + lane_off + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 8)
              if (c * 32 + i < DV)
                *reinterpret_cast<uint4*>(vrow + c * 32 + i) =
                    make_uint4(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                               pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                               pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                               pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
          }
        }
#pragma unroll
        for (int c = 0; c < (DQK + 31) / 32; ++c) {
          if (c * 32 < p.dk_cols) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_acc1 + lane_off + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * p.scale);
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < 32; i += 8)
                if (c * 32 + i < p.dk_cols)
                  *reinterpret_cast<uint4*>(krow + c * 32 + i) =
                      make_uint4(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                 pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                                 pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                                 pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace saicv
