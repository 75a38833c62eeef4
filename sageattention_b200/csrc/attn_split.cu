// Fused INT8-QK / FP8-PV attention for sm_100a — "split-row" softmax: TWO threads per Q row.
//
// Same tile pipeline as attn.cu (128-row Q tile per CTA, 64-key tiles, S double-buffered in TMEM, two CTAs per SM,
// QK0 QK1 | PV0 QK2 | PV1 QK3 ... on the in-order tensor pipe), but the softmax of a tile is done by 256 threads:
// warpgroup 0 owns S columns [0,32) of every row, warpgroup 1 owns [32,64) (a warp reaches TMEM lane quadrant warp%4,
// so warps w and w+4 see the same 32 rows).  That puts four softmax warps on every SM sub-partition instead of two and
// halves each warp's serial chain per tile (TMEM load -> max -> exp -> pack -> TMEM store), which is what bounds
// attn.cu: its MUFU pipe idles ~40 % of the time behind those chains.  The only cross-thread step is the row max:
// the two halves exchange their partial max through shared memory and a 64-thread named barrier; float max is exact,
// so m, alpha and P are bit-identical to the one-thread-per-row kernel.  The row sum d stays split (both halves scale
// their partial sums by the same alpha) and is combined once in the epilogue.  There is no separate correction
// warpgroup: each half rescales its own half of the O columns when the running max moved.
#include "attn_common.cuh"

namespace sab {

constexpr int kSplitThreads = 384;   // warpgroups: 0 = softmax cols [0,32), 1 = softmax cols [32,64), 2 = TMA / MMA / 2 idle

// 64-thread named barrier of the two warps that share TMEM lane quadrant q (immediate ids: ptxas then reserves 5 barriers, not 16)
__device__ __forceinline__ void split_bar_sync(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
// 128 x (96 + 96 + 48) = 30720 = the CTA's register pool (80 x 384).  Values are kept multiples of 16: 104/104/32 has the
// same sum but never got its registers on B200 (the allocator appears to round a warp's request up to 16-register units).
__device__ __forceinline__ void setmaxnreg_inc_96s() { asm volatile("setmaxnreg.inc.sync.aligned.u32 96;"); }

template <int D, bool kKT, typename OutT>
__global__ void __launch_bounds__(kSplitThreads, 2)
sage_attn_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr uint32_t K_TILE = BN * D;
  constexpr uint32_t V_TILE = D * BN;
  constexpr int NS = (D == 128) ? 5 : 10;
  constexpr int SWQK = (D == 128) ? 128 : 64;
  constexpr uint32_t Q_BYTES = BM * D;
  constexpr int NG = kKT ? 4 : 1;
  constexpr int HC = BN / 2;        // S columns per thread
  constexpr int OC = D / 2;         // O columns per thread

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + NS * K_TILE;
  float* s_xch = reinterpret_cast<float*>(sV + NS * V_TILE);   // [2 parity][2 half][128] partial row max (and d at the end)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_xch + 4 * BM);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;    // [2] step(t) retired: S(t+2) ready in buffer t&1 AND PV(t) accumulated into O
  uint64_t* p_full = bars + 3;    // [2] 256 arrivals: both halves stored P(j) and rescaled their O columns
  uint64_t* kv_full = bars + 5;
  uint64_t* kv_empty = kv_full + NS;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(kv_empty + NS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int qt = blockIdx.x;
  if (p.causal) qt = p.n_q_tiles - 1 - qt;
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
    if (qt * BM >= q_len) return;
  }
  int n_kv = (kv_len + BN - 1) / BN;
  if (p.causal) n_kv = min(n_kv, (p.causal_q_offset + (qt + 1) * BM + BN - 1) / BN);

  if (warp == 8 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(p_full + i, 256);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<kTmemCols>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  auto s_parity = [](int t) { return uint32_t(t >> 1) & 1u; };

  if (warp >= 8) {
    setmaxnreg_dec_48();
    if (warp == 8) {
      // =============================== TMA producer ===============================
      if (lane == 0 && n_kv > 0) {
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_4d(sQ, &tmQ, q_full, 0, q_off + qt * BM, h, tb);
        for (int j = 0; j < n_kv; ++j) {
          const int s = j % NS;
          const uint32_t ph = (j / NS) & 1;
          int kc = k_off + j * BN, vc = v_off + j * BN, kb = tb;
          if (p.kv_seg_len > 0) {
            const int seg = (j * BN) / p.kv_seg_len;
            kc = vc = j * BN - seg * p.kv_seg_len;
            kb = seg * p.B + b;
          }
          mbar_wait_wd(kv_empty + s, ph ^ 1);
          mbar_expect_tx(kv_full + s, K_TILE + V_TILE);
          tma_load_4d(sK + s * K_TILE, &tmK, kv_full + s, 0, kc, hk, kb);
          tma_load_4d(sV + s * V_TILE, &tmV, kv_full + s, vc, 0, hk, kb);
        }
      }
    } else if (warp == 9) {
      // =============================== MMA issuer ===============================
      if (n_kv > 0) {
        constexpr uint32_t idesc_qk = make_idesc(2, 1, 1, BM, BN);
        constexpr uint32_t idesc_pv = make_idesc(1, 0, 0, BM, D);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dQ = make_smem_desc<SWQK>(smem_u32(sQ));
        const uint64_t dK0 = make_smem_desc<SWQK>(smem_u32(sK));
        const uint64_t dV0 = make_smem_desc<64>(smem_u32(sV));
        auto issue_qk = [&](int t, bool wait_kv) {
          const int st = t % NS;
          if (wait_kv) {
            mbar_wait_wd(kv_full + st, (t / NS) & 1);
            tc_fence_after();
          }
          const uint64_t dK = dK0 + uint64_t(st) * (K_TILE >> 4);
          const uint32_t tS = tmem_u + (t & 1) * BN;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 32; ++k) umma_i8_ss(tS, dQ + 2 * k, dK + 2 * k, idesc_qk, k > 0);
          }
        };
        mbar_wait_wd(q_full, 0);
        issue_qk(0, true);
        if (elect_one()) tc_commit(s_full + 0);
        if (n_kv > 1) {
          issue_qk(1, true);
          if (elect_one()) tc_commit(s_full + 1);
        }
        for (int j = 0; j < n_kv; ++j) {
          if (j + 2 < n_kv) mbar_wait_wd(kv_full + (j + 2) % NS, ((j + 2) / NS) & 1);
          mbar_wait_wd(p_full + (j & 1), (j >> 1) & 1);
          tc_fence_after();
          const int st = j % NS;
          const uint64_t dV = dV0 + uint64_t(st) * (V_TILE >> 4);
          const uint32_t tP = tmem_u + (j & 1) * BN;   // P half k (32 keys, 8 columns) sits at the start of S half k
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BN / 32; ++k) umma_f8_ts(tmem_u + 128, tP + HC * k, dV + 2 * k, idesc_pv, (j > 0 || k > 0));
          }
          if (j + 2 < n_kv) issue_qk(j + 2, false);
          if (elect_one()) {
            tc_commit(s_full + (j & 1));
            tc_commit(kv_empty + st);
          }
        }
      }
    }
  } else {
    // =============================== softmax (half of the columns) / correction / epilogue ===============================
    setmaxnreg_inc_96s();
    const int half = warp >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tOh = tmem_base + lane_off + 128 + half * OC;
    const int q_row = qt * BM + row;

    int q_idx = (q_blk0 + qt) * p.q_mult;
    if (p.q_gran == SAB_GRAN_PER_WARP) q_idx += row >> 5;
    if (p.q_gran == SAB_GRAN_PER_THREAD) q_idx += (row >> 5) * 8 + (row & 7);
    const float* qs_base = p.q_scale + (varlen ? int64_t(h) : (int64_t(b) * p.Hq + h) * p.qs_stride_bh);
    const float* ks_base = p.k_scale + (varlen ? int64_t(hk) : (int64_t(b) * p.Hkv + hk) * p.ks_stride_bh);
    const float qss = qs_base[int64_t(q_idx) * p.qs_stride_idx] * p.sm_scale_log2;

    float m = kMaskValue;
    float d = 0.f;   // this half's share of the row sum

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t tS = tmem_base + lane_off + (j & 1) * BN + half * HC;
      float coef[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) coef[g] = ks_base[int64_t((k_blk0 + j) * NG + g) * p.ks_stride_idx] * qss;
      int limit = kv_len - j * BN;
      if (p.causal) limit = min(limit, p.causal_q_offset + q_row - j * BN + 1);
      limit -= half * HC;   // visible keys among this thread's 32 columns
      const bool masked_tile = (kv_len - j * BN < BN) || (p.causal && (j + 1) * BN > p.causal_q_offset + qt * BM + 1);

      mbar_wait_wd(s_full + (j & 1), s_parity(j));
      tc_fence_after();
      uint32_t s[HC];
      tmem_ld32(tS, s);
      tc_wait_ld();

      auto tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        if constexpr (MASKED) {
#pragma unroll
          for (int i = 0; i < HC; ++i)
            if (i >= limit) s[i] = uint32_t(kIntSentinel);
        }
        float mx = kMaskValue;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          int v;
          if constexpr (kKT) {
            v = kIntSentinel;
#pragma unroll
            for (int i8 = 0; i8 < HC; i8 += 8) v = __vimax3_s32(v, int(s[i8 + 2 * g]), int(s[i8 + 2 * g + 1]));
          } else {
            int v0 = kIntSentinel, v1 = kIntSentinel, v2 = kIntSentinel, v3 = kIntSentinel;
#pragma unroll
            for (int i = 0; i < HC; i += 8) {
              v0 = __vimax3_s32(v0, int(s[i]), int(s[i + 1]));
              v1 = __vimax3_s32(v1, int(s[i + 2]), int(s[i + 3]));
              v2 = __vimax3_s32(v2, int(s[i + 4]), int(s[i + 5]));
              v3 = __vimax3_s32(v3, int(s[i + 6]), int(s[i + 7]));
            }
            v = max(max(v0, v1), max(v2, v3));
          }
          float c = float(v) * coef[g];
          if constexpr (MASKED) c = (v == kIntSentinel) ? kMaskValue : c;
          mx = fmaxf(mx, c);
        }
        // exchange the partial max with the other half of this row (double-buffered on j&1: the partner's read of
        // slot j&1 is ordered before my next write to it by the barrier of tile j+1)
        float* xch = s_xch + (j & 1) * 2 * BM;
        xch[half * BM + row] = mx;
        split_bar_sync(wq);
        mx = fmaxf(mx, xch[(half ^ 1) * BM + row]);

        const float m_new = fmaxf(m, mx - kFp8Offset);   // update_mdo, attn_utils.cuh:377-396
        const float alpha = ex2_approx(m - m_new);
        d *= alpha;
        m = m_new;

        uint64_t coef2[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) coef2[g] = pack_f2(coef[g], coef[g]);
        const uint64_t nm2 = pack_f2(-m_new, -m_new);
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
        uint32_t pk[HC / 4];
#pragma unroll
        for (int w = 0; w < HC / 4; ++w) {
          float e[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const int i = 4 * w + u;
            const int g = kKT ? ((i & 7) >> 1) : 0;
            const uint64_t f2 = pack_f2(__int2float_rn(int(s[i])), __int2float_rn(int(s[i + 1])));
            float y0, y1;
            unpack_f2(ffma2(f2, coef2[g], nm2), y0, y1);
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
        {
          float a0, a1, a2, a3;
          unpack_f2(fadd2(acc[0], acc[1]), a0, a1);
          unpack_f2(fadd2(acc[2], acc[3]), a2, a3);
          d += (a0 + a1) + (a2 + a3);
        }
        tmem_st8(tS, pk);   // P half over the first 8 columns of this thread's own S half

        // correction of this thread's half of the O row (both halves see the same alpha -> same decision)
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
          mbar_wait_wd(s_full + ((j + 1) & 1), s_parity(j + 1));   // step(j-1) retired: PV(j-1) is in O
          tc_fence_after();
          const uint64_t alpha2 = pack_f2(alpha, alpha);
#pragma unroll
          for (int ch = 0; ch < OC / 32; ++ch) {
            uint32_t r[32];
            tmem_ld32(tOh + ch * 32, r);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float lo, hi;
              unpack_f2(fmul2(pack_f2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), lo, hi);
              r[i] = __float_as_uint(lo);
              r[i + 1] = __float_as_uint(hi);
            }
            tmem_st32(tOh + ch * 32, r);
          }
        }
      };
      if (masked_tile) tile(std::true_type{});
      else tile(std::false_type{});

      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full + (j & 1));
    }

    // ---- epilogue: combine the two halves of d, then each half writes its OC columns
    float* xd = s_xch + (n_kv & 1) * 2 * BM;   // slot of parity n_kv: last used at tile n_kv-2, ordered by tile n_kv-1's barrier
    xd[half * BM + row] = d;
    split_bar_sync(wq);
    d += xd[(half ^ 1) * BM + row];

    OutT* orow = reinterpret_cast<OutT*>(p.out) + (varlen ? 0 : int64_t(b) * p.o_stride_b) + int64_t(h) * p.o_stride_h +
                 int64_t(q_off + q_row) * p.o_stride_s + half * OC;
    const bool row_ok = q_row < q_len;
    const float* vs = p.v_scale ? p.v_scale + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + half * OC : nullptr;
    const float* vm = p.v_mean ? p.v_mean + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + half * OC : nullptr;
    if (n_kv > 0) {
      mbar_wait_wd(s_full + ((n_kv + 1) & 1), s_parity(n_kv + 1));
      tc_fence_after();
      const float inv = rcp_approx(d);
#pragma unroll
      for (int ch = 0; ch < OC / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tOh + ch * 32, r);
        tc_wait_ld();
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
    } else if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
      for (int v4 = 0; v4 < OC / 8; ++v4) dst[v4] = make_uint4(0, 0, 0, 0);
    }
    if (p.lse != nullptr && row_ok && half == 0) {
      const int64_t li = varlen ? (int64_t(h) * p.Sq + q_off + q_row) : ((int64_t(b) * p.Hq + h) * p.Sq + q_row);
      p.lse[li] = n_kv > 0 ? lg2_approx(d) + m : -INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<kTmemCols>(tmem_base);
}

template <int D, bool kKT, typename OutT>
int launch_attn_split(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                      cudaStream_t stream) {
  constexpr int NS = (D == 128) ? 5 : 10;
  size_t smem = size_t(BM) * D + size_t(NS) * 2 * BN * D + 4 * BM * sizeof(float) + 512;
  if (smem < 80 * 1024) smem = 80 * 1024;   // keep it at two CTAs per SM (TMEM: 2 x 256 columns)
  auto kern = sage_attn_split_kernel<D, kKT, OutT>;
  static bool configured[64] = {};
  if (int st = ensure_dynamic_smem(kern, smem, configured)) return st;
  kern<<<grid, kSplitThreads, smem, stream>>>(tq, tk, tv, p);
  SAB_CUDA_OK(cudaGetLastError());
  return SAB_OK;
}

#define SAB_INST(D, KT, T) \
  template int launch_attn_split<D, KT, T>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const AttnParams&, dim3, cudaStream_t);
SAB_INST(128, true, __nv_bfloat16)
SAB_INST(128, true, __half)
SAB_INST(128, false, __nv_bfloat16)
SAB_INST(128, false, __half)
SAB_INST(64, true, __nv_bfloat16)
SAB_INST(64, true, __half)
SAB_INST(64, false, __nv_bfloat16)
SAB_INST(64, false, __half)
#undef SAB_INST

}  // namespace sab
