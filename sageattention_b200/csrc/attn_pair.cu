// Fused INT8-QK / FP8-PV attention for sm_100a — PAIRED-TILE kernel: one CTA per SM works on TWO adjacent 128-row
// Q tiles of one (batch, head) against the same K / V^T stream.
//
// Why: the exponentials (MUFU.EX2, 16 lanes/clk/SM) bound 8-bit attention, not the tensor core.  With two independent
// CTAs per SM (attn.cu) the two softmax streams drift into phase and leave the MUFU idle ~40 % of the time.  Here the
// two exp warpgroups hand an "exp token" back and forth through named barriers (exp A0, B0, A1, B1, ...): while one
// tile runs its 512-cycle exp phase at full MUFU rate, the other does everything that is not exp — mbarrier waits,
// TMEM loads, row max -> alpha, P store — and K / V^T tiles are loaded into shared memory once for both Q tiles.
//
// 512 threads, four warpgroups (registers re-split with setmaxnreg: 176 / 176 / 80 / 48):
//   warps 0-3 / 4-7  : exp warpgroup of tile 0 / tile 1 — one thread per Q row (TMEM lane == row)
//   warps 8-11       : correction — rescales row r of O0 and O1 (in TMEM) when that row's running max moved
//   warp 12          : TMA producer (Q0, Q1 once; K and V^T in 128-key stages through an NS-deep ring)
//   warp 13          : tcgen05.mma issuer (single thread) + TMEM allocator;  warps 14-15 idle
// Tensor memory (512 columns): tile t uses columns [256t, 256t+256): S double-buffered at +0 / +64 (P aliases the first
// 16 columns of its S buffer), O fp32 at +128.  Numerics are identical to attn.cu (reference-exact P, m, d).
#include "attn_common.cuh"

namespace sab {

constexpr int kPairThreads = 512;
constexpr uint32_t kPairTmemCols = 512;

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void setmaxnreg_inc_176() { asm volatile("setmaxnreg.inc.sync.aligned.u32 176;"); }
__device__ __forceinline__ void setmaxnreg_dec_80() { asm volatile("setmaxnreg.dec.sync.aligned.u32 80;"); }

template <int D, bool kKT, typename OutT>
__global__ void __launch_bounds__(kPairThreads, 1)
sage_attn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int NS = (D == 128) ? 5 : 10;     // K/V ring depth (128-key stages)
  constexpr int SWQK = (D == 128) ? 128 : 64; // swizzle span of the Q/K tiles (= row bytes)
  constexpr uint32_t Q_BYTES = BM * D, K_BYTES = LK * D, V_BYTES = D * LK;
  constexpr uint64_t K_HALF = (uint64_t(BN) * D) >> 4;  // descriptor delta: keys 64..127 of a K stage
  constexpr uint64_t V_HALF = uint64_t(BN) >> 4;        // descriptor delta: byte column 64 of a V^T stage
  constexpr int NG = kKT ? 4 : 1;                       // dequant-scale groups per 64-key tile

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                       // [2] Q tiles
  uint8_t* sK = sQ + 2 * Q_BYTES;
  uint8_t* sV = sK + NS * K_BYTES;
  float* s_alpha = reinterpret_cast<float*>(sV + NS * V_BYTES);   // [2 tiles][2 buffers][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_alpha + 2 * 2 * BM);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;    // [t][2] step_t(j) retired: S_t(j+2) ready in buffer j&1 AND PV_t(j) accumulated
  uint64_t* p_full = bars + 5;    // [t][2] 256 arrivals: P_t(j) stored (exp) + O_t rescaled (correction)
  uint64_t* a_full = bars + 9;    // [t][2] 128 arrivals: alpha_t(j) published
  uint64_t* kv_full = bars + 13;  // [NS]
  uint64_t* kv_empty = kv_full + NS;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(kv_empty + NS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---------------- work decode (uniform across the CTA)
  const int n_pairs = (p.n_q_tiles + 1) / 2;
  int pair = blockIdx.x;
  if (p.causal) pair = n_pairs - 1 - pair;  // heaviest first
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const bool varlen = p.cu_q != nullptr;
  int q_len = p.Sq, kv_len = p.Sk, q_off = 0, k_off = 0, v_off = 0, tb = b;
  int q_blk0 = 0, k_blk0 = 0;
  if (varlen) {
    q_off = p.cu_q[b];
    q_len = p.cu_q[b + 1] - q_off;
    k_off = p.cu_k[b];
    kv_len = p.cu_k[b + 1] - k_off;
    v_off = p.cu_v[b];
    q_blk0 = p.cu_qs[b];
    k_blk0 = p.cu_ks[b];
    tb = 0;
    if (pair * 2 * BM >= q_len) return;
  }
  int n_kv_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int qt = 2 * pair + t;
    int n = (kv_len + BN - 1) / BN;
    if (p.causal) n = min(n, (p.causal_q_offset + (qt + 1) * BM + BN - 1) / BN);
    n_kv_t[t] = (qt * BM < q_len) ? n : 0;
  }
  const int n_max = max(n_kv_t[0], n_kv_t[1]);
  const int n_st = (n_max + 1) / 2;

  // ---------------- one-time setup
  if (warp == 12 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(p_full + i, 256);
      mbar_init(a_full + i, 128);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 13) tmem_alloc<kPairTmemCols>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  auto s_parity = [](int t) { return uint32_t(t >> 1) & 1u; };

  if (warp >= 12) {
    setmaxnreg_dec_48();
    if (warp == 12) {
      // =============================== TMA producer ===============================
      if (lane == 0 && n_max > 0) {
        const int nq = (n_kv_t[0] > 0) + (n_kv_t[1] > 0);
        mbar_expect_tx(q_full, Q_BYTES * nq);
        if (n_kv_t[0] > 0) tma_load_4d(sQ, &tmQ, q_full, 0, q_off + (2 * pair) * BM, h, tb);
        if (n_kv_t[1] > 0) tma_load_4d(sQ + Q_BYTES, &tmQ, q_full, 0, q_off + (2 * pair + 1) * BM, h, tb);
        for (int jj = 0; jj < n_st; ++jj) {
          const int s = jj % NS;
          const uint32_t ph = (jj / NS) & 1;
          int kc = k_off + jj * LK, vc = v_off + jj * LK, kb = tb;
          if (p.kv_seg_len > 0) {
            const int seg = (jj * LK) / p.kv_seg_len;
            kc = vc = jj * LK - seg * p.kv_seg_len;
            kb = seg * p.B + b;
          }
          mbar_wait_wd(kv_empty + s, ph ^ 1);
          mbar_expect_tx(kv_full + s, K_BYTES + V_BYTES);
          tma_load_4d(sK + s * K_BYTES, &tmK, kv_full + s, 0, kc, hk, kb);
          tma_load_4d(sV + s * V_BYTES, &tmV, kv_full + s, vc, 0, hk, kb);
        }
      }
    } else if (warp == 13) {
      // =============================== MMA issuer ===============================
      // per tile t: step_t(j) = PV_t(j) ; QK_t(j+2) ; commit -> s_full[t][j&1].  Tiles alternate: 0,1,0,1,...
      if (n_max > 0) {   // whole warp runs the loop (uniform control flow); one elected lane issues
        constexpr uint32_t idesc_qk = make_idesc(2, 1, 1, BM, BN);
        constexpr uint32_t idesc_pv = make_idesc(1, 0, 0, BM, D);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dQ0 = make_smem_desc<SWQK>(smem_u32(sQ));
        const uint64_t dK0 = make_smem_desc<SWQK>(smem_u32(sK));
        const uint64_t dV0 = make_smem_desc<128>(smem_u32(sV));
        auto issue_qk = [&](int t, int j) {
          const int st = (j >> 1) % NS;
          const uint64_t dQ = dQ0 + uint64_t(t) * (Q_BYTES >> 4);
          const uint64_t dK = dK0 + uint64_t(st) * (K_BYTES >> 4) + uint64_t(j & 1) * K_HALF;
          const uint32_t tS = tmem_u + t * 256 + (j & 1) * BN;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 32; ++k) umma_i8_ss(tS, dQ + 2 * k, dK + 2 * k, idesc_qk, k > 0);
          }
        };
        mbar_wait_wd(q_full, 0);
        mbar_wait_wd(kv_full + 0, 0);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (n_kv_t[t] > 0) { issue_qk(t, 0); if (elect_one()) tc_commit(s_full + t * 2 + 0); }
          if (n_kv_t[t] > 1) { issue_qk(t, 1); if (elect_one()) tc_commit(s_full + t * 2 + 1); }
        }
#ifdef SAB_TIMELINE
        const bool tl_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
        long long* tl = reinterpret_cast<long long*>(p.dbg) + 4096;
#endif
        for (int j = 0; j < n_max; ++j) {
          SAB_TL(8);
          if ((j & 1) == 0 && j + 2 < n_max) mbar_wait_wd(kv_full + ((j + 2) >> 1) % NS, (((j + 2) >> 1) / NS) & 1);
          const int st = (j >> 1) % NS;
          const uint64_t dV = dV0 + uint64_t(st) * (V_BYTES >> 4) + uint64_t(j & 1) * V_HALF;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j < n_kv_t[t]) {
              mbar_wait_wd(p_full + t * 2 + (j & 1), (j >> 1) & 1);
              tc_fence_after();
              if (t == 0) SAB_TL(9);
              const uint32_t tP = tmem_u + t * 256 + (j & 1) * BN;
              const uint32_t tO = tmem_u + t * 256 + 128;
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BN / 32; ++k) umma_f8_ts(tO, tP + 8 * k, dV + 2 * k, idesc_pv, (j > 0 || k > 0));
              }
              if (j + 2 < n_kv_t[t]) issue_qk(t, j + 2);
              if (elect_one()) tc_commit(s_full + t * 2 + (j & 1));
              if (t == 0) SAB_TL(10);
            }
          }
          if (((j & 1) == 1 || j == n_max - 1) && elect_one()) tc_commit(kv_empty + st);
          SAB_TL(11);
        }
      }
    }
  } else if (warp >= 8) {
    // =============================== correction ===============================
    setmaxnreg_dec_80();
    const int row = (warp - 8) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp - 8) * 32) << 16;
    for (int j = 0; j < n_max; ++j) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (j < n_kv_t[t]) {
          mbar_wait_wd(a_full + t * 2 + (j & 1), (j >> 1) & 1);
          const float alpha = s_alpha[(t * 2 + (j & 1)) * BM + row];
          if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
            mbar_wait_wd(s_full + t * 2 + ((j + 1) & 1), s_parity(j + 1));   // step_t(j-1) retired: PV_t(j-1) is in O_t
            tc_fence_after();
            const uint32_t tO = tmem_base + t * 256 + 128 + lane_off;
            const uint64_t alpha2 = pack_f2(alpha, alpha);
#pragma unroll
            for (int ch = 0; ch < D / 32; ++ch) {
              uint32_t r[32];
              tmem_ld32(tO + ch * 32, r);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float lo, hi;
                unpack_f2(fmul2(pack_f2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), lo, hi);
                r[i] = __float_as_uint(lo);
                r[i + 1] = __float_as_uint(hi);
              }
              tmem_st32(tO + ch * 32, r);
            }
            tc_wait_st();
          }
          tc_fence_before();
          mbar_arrive(p_full + t * 2 + (j & 1));
        }
      }
    }
  } else {
    // =============================== exp warpgroups (tile t = warp / 4) + epilogue ===============================
    setmaxnreg_inc_176();
    const int t = warp >> 2;
    const int wq = warp & 3;
    const int qt = 2 * pair + t;
    const int n_kv = n_kv_t[t];
    const int row = wq * 32 + lane;  // TMEM lane == Q row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t t_base = tmem_base + t * 256 + lane_off;
    const int q_row = qt * BM + row;
    const bool dump = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t == 0;

    int q_idx = (q_blk0 + qt) * p.q_mult;
    if (p.q_gran == SAB_GRAN_PER_WARP) q_idx += row >> 5;
    if (p.q_gran == SAB_GRAN_PER_THREAD) q_idx += (row >> 5) * 8 + (row & 7);
    const float* qs_base = p.q_scale + (varlen ? int64_t(h) : (int64_t(b) * p.Hq + h) * p.qs_stride_bh);
    const float* ks_base = p.k_scale + (varlen ? int64_t(hk) : (int64_t(b) * p.Hkv + hk) * p.ks_stride_bh);
    const float qss = n_kv > 0 ? qs_base[int64_t(q_idx) * p.qs_stride_idx] * p.sm_scale_log2 : 0.f;

    float m = kMaskValue;
    float d = 0.f;
#ifdef SAB_TIMELINE
    const bool tl_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    long long* tl = reinterpret_cast<long long*>(p.dbg) + 4096;
#endif
    // exp token: barrier 1 = "tile 0 may run its exp phase", barrier 2 = "tile 1 may".  Tile 1 grants the first turn.
    if (t == 1) named_bar_arrive(1, 256);

    for (int j = 0; j < n_max; ++j) {
      const bool mine = j < n_kv;
      const uint32_t tS = t_base + (j & 1) * BN;
      uint32_t s[BN];
      float coef[NG];
      float m_new = m;
      int limit = BN;
      bool masked_tile = false;
      if (mine) {
#pragma unroll
        for (int g = 0; g < NG; ++g) coef[g] = ks_base[int64_t((k_blk0 + j) * NG + g) * p.ks_stride_idx] * qss;
        limit = kv_len - j * BN;
        if (p.causal) limit = min(limit, p.causal_q_offset + q_row - j * BN + 1);
        masked_tile = (kv_len - j * BN < BN) || (p.causal && (j + 1) * BN > p.causal_q_offset + qt * BM + 1);
        SAB_TL(0);
        mbar_wait_wd(s_full + t * 2 + (j & 1), s_parity(j));
        tc_fence_after();
        SAB_TL(1);
        {
          uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
          uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
          tmem_ld32(tS, lo);
          tmem_ld32(tS + 32, hi);
          tc_wait_ld();
        }
        SAB_TL(2);
        if (dump && j == 0) {
#pragma unroll
          for (int i = 0; i < BN; ++i) p.dbg[row * BN + i] = int(s[i]);
        }
        // ---- row max (integer max per dequant-scale group, scales are positive) -> m, alpha (update_mdo)
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= limit) s[i] = uint32_t(kIntSentinel);
        }
        float mx = kMaskValue;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          int v;
          if constexpr (kKT) {
            v = kIntSentinel;
#pragma unroll
            for (int i8 = 0; i8 < BN; i8 += 8) v = __vimax3_s32(v, int(s[i8 + 2 * g]), int(s[i8 + 2 * g + 1]));
          } else {
            int v0 = kIntSentinel, v1 = kIntSentinel, v2 = kIntSentinel, v3 = kIntSentinel;
#pragma unroll
            for (int i = 0; i < BN; i += 8) {
              v0 = __vimax3_s32(v0, int(s[i]), int(s[i + 1]));
              v1 = __vimax3_s32(v1, int(s[i + 2]), int(s[i + 3]));
              v2 = __vimax3_s32(v2, int(s[i + 4]), int(s[i + 5]));
              v3 = __vimax3_s32(v3, int(s[i + 6]), int(s[i + 7]));
            }
            v = max(max(v0, v1), max(v2, v3));
          }
          const float c = (v == kIntSentinel) ? kMaskValue : float(v) * coef[g];
          mx = fmaxf(mx, c);
        }
        m_new = fmaxf(m, mx - kFp8Offset);
        const float alpha = ex2_approx(m - m_new);
        d *= alpha;
        m = m_new;
        s_alpha[(t * 2 + (j & 1)) * BM + row] = alpha;   // the correction warpgroup rescales O_t concurrently
        mbar_arrive(a_full + t * 2 + (j & 1));
        SAB_TL(3);
      }

      named_bar_sync(1 + t, 256);          // ---- my turn on the MUFU
      if (mine) {
       auto exp_part = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        const float nm = -m_new;
        uint64_t coef2[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) coef2[g] = pack_f2(coef[g], coef[g]);
        const uint64_t nm2 = pack_f2(nm, nm);
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
        uint32_t pk[BN / 4];
#pragma unroll
        for (int w = 0; w < BN / 4; ++w) {
          float e[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const int i = 4 * w + u;
            const int g = kKT ? ((i & 7) >> 1) : 0;
            const uint64_t f2 = pack_f2(__int2float_rn(int(s[i])), __int2float_rn(int(s[i + 1])));
            float y0, y1;
            unpack_f2(ffma2(f2, coef2[g], nm2), y0, y1);   // fmaf(S, sm_scale', -m)  (attn_utils.cuh:450)
            e[u] = ex2_approx(y0);
            e[u + 1] = ex2_approx(y1);
            if constexpr (MASKED) {
              e[u] = (i < limit) ? e[u] : 0.f;
              e[u + 1] = (i + 1 < limit) ? e[u + 1] : 0.f;
            }
            acc[(w & 1) * 2 + (u >> 1)] = fadd2(acc[(w & 1) * 2 + (u >> 1)], pack_f2(e[u], e[u + 1]));
          }
          pk[w] = pack_e4m3x4(e[0], e[1], e[2], e[3]);
        }
        named_bar_arrive(1 + (t ^ 1), 256);   // ---- hand the MUFU to the other tile
        {
          float a0, a1, a2, a3;
          unpack_f2(fadd2(acc[0], acc[1]), a0, a1);
          unpack_f2(fadd2(acc[2], acc[3]), a2, a3);
          d += (a0 + a1) + (a2 + a3);
        }
        SAB_TL(4);
        tmem_st16(tS, pk);
        if (dump && j == 0) {
#pragma unroll
          for (int w = 0; w < BN / 4; ++w) p.dbg[128 * BN + row * 16 + w] = int(pk[w]);
        }
        SAB_TL(5);
        tc_wait_st();
        SAB_TL(6);
        tc_fence_before();
        mbar_arrive(p_full + t * 2 + (j & 1));
        SAB_TL(7);
       };
       if (masked_tile) exp_part(std::true_type{});
       else exp_part(std::false_type{});
      } else {
        named_bar_arrive(1 + (t ^ 1), 256);
      }
    }

    // ---- epilogue: O / d * v_scale (+ v_mean) -> fp16/bf16, 16-byte stores (…sm89.cuh:572-703)
    if (qt * BM < q_len || !varlen) {
      OutT* orow = reinterpret_cast<OutT*>(p.out) + (varlen ? 0 : int64_t(b) * p.o_stride_b) + int64_t(h) * p.o_stride_h +
                   int64_t(q_off + q_row) * p.o_stride_s;
      const bool row_ok = q_row < q_len;
      const float* vs = p.v_scale ? p.v_scale + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D : nullptr;
      const float* vm = p.v_mean ? p.v_mean + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D : nullptr;
      if (n_kv > 0) {
        mbar_wait_wd(s_full + t * 2 + ((n_kv + 1) & 1), s_parity(n_kv + 1));   // step_t(n_kv-1) retired: O_t is final
        tc_fence_after();
        const float inv = rcp_approx(d);
#pragma unroll
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(t_base + 128 + ch * 32, r);
          tc_wait_ld();
          if (dump) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.dbg[128 * BN + 128 * 16 + row * D + ch * 32 + i] = int(r[i]);
          }
          uint32_t o16[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a = __uint_as_float(r[i]) * inv, c = __uint_as_float(r[i + 1]) * inv;
            if (vs) {
              a *= vs[ch * 32 + i];
              c *= vs[ch * 32 + i + 1];
            }
            if (vm) {
              a += vm[ch * 32 + i];
              c += vm[ch * 32 + i + 1];
            }
            o16[i / 2] = pack2<OutT>(a, c);
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) dst[v4] = make_uint4(o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
          }
        }
        if (dump) {
          p.dbg[128 * BN + 128 * 16 + 128 * D + row] = __float_as_int(d);
          p.dbg[128 * BN + 128 * 16 + 128 * D + 128 + row] = __float_as_int(m);
        }
      } else if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int v4 = 0; v4 < D / 8; ++v4) dst[v4] = make_uint4(0, 0, 0, 0);
      }
      if (p.lse != nullptr && row_ok) {
        const int64_t li = varlen ? (int64_t(h) * p.Sq + q_off + q_row) : ((int64_t(b) * p.Hq + h) * p.Sq + q_row);
        p.lse[li] = n_kv > 0 ? lg2_approx(d) + m : -INFINITY;
      }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc<kPairTmemCols>(tmem_base);
}

template <int D, bool kKT, typename OutT>
int launch_attn_pair(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                     cudaStream_t stream) {
  constexpr int NS = (D == 128) ? 5 : 10;
  const size_t smem = size_t(2) * BM * D + size_t(NS) * 2 * LK * D + 2 * 2 * BM * sizeof(float) + 512;
  auto kern = sage_attn_pair_kernel<D, kKT, OutT>;
  static bool configured[64] = {};
  if (int st = ensure_dynamic_smem(kern, smem, configured)) return st;
  kern<<<grid, kPairThreads, smem, stream>>>(tq, tk, tv, p);
  SAB_CUDA_OK(cudaGetLastError());
  return SAB_OK;
}

// explicit instantiations used by attn.cu's dispatcher
#define SAB_INST(DD, KT, T) \
  template int launch_attn_pair<DD, KT, T>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const AttnParams&, dim3, cudaStream_t)
SAB_INST(128, true, __nv_bfloat16);
SAB_INST(128, true, __half);
SAB_INST(128, false, __nv_bfloat16);
SAB_INST(128, false, __half);
SAB_INST(64, true, __nv_bfloat16);
SAB_INST(64, true, __half);
SAB_INST(64, false, __nv_bfloat16);
SAB_INST(64, false, __half);
#undef SAB_INST

}  // namespace sab
