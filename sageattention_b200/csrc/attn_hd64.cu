// Fused INT8-QK / FP8-PV attention for sm_100a, head_dim 64: FOUR co-resident CTAs per SM.
//
// At hd=64 a 128x64 tile costs only 128 tensor-core cycles against 512 MUFU cycles of exponentials, and the kernel
// of attn.cu (two CTAs per SM) leaves the MUFU idle ~45 % of the time behind per-tile latencies.  With hd=64 the O
// accumulator needs only 64 TMEM columns, so one CTA fits in 128 columns (S single-buffered 64 + O 64) and ~50 KB of
// shared memory: four CTAs = four independent softmax warps per scheduler hide each other's mbarrier / TMEM / MMA
// round trips.  256 threads: warps 0-3 softmax (one thread per Q row; the thread also rescales its own O row),
// warp 4 TMA producer, warp 5 MMA issuer + TMEM allocator, warps 6-7 idle (setmaxnreg is per warpgroup: 96 / 32).
// Per tile and CTA the chain is serial — QK(j) -> softmax(j) -> PV(j), QK(j+1) — the overlap comes from the other three
// CTAs.  Numerics identical to attn.cu (reference-exact P, m, d; 64-key tiles; fp32 accumulation in TMEM).
#include "attn_common.cuh"

namespace sab {

constexpr int kHd64Threads = 256;
constexpr uint32_t kHd64TmemCols = 128;
constexpr float kLazyTau = 4.0f;   // lazy-max threshold in binades (same rule as attn_alt.cu)

__device__ __forceinline__ void setmaxnreg_inc_96() { asm volatile("setmaxnreg.inc.sync.aligned.u32 96;"); }
__device__ __forceinline__ void setmaxnreg_dec_32() { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;"); }

template <bool kKT, typename OutT, bool kLazy>
__global__ void __launch_bounds__(kHd64Threads, 4)
sage_attn_hd64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  constexpr int D = 64;
  constexpr int NS = 5;                          // K / V^T ring slots (64-key tiles)
  constexpr uint32_t Q_BYTES = BM * D, K_TILE = BN * D, V_TILE = D * BN;
  constexpr int NG = kKT ? 4 : 1;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + NS * K_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * V_TILE);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;    // phase t: S(t) ready and PV(t-1) accumulated (phase n_kv: O final)
  uint64_t* p_full = bars + 2;    // 128 arrivals: P(j) stored, O rescaled
  uint64_t* kv_full = bars + 3;
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

  if (warp == 4 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    for (int i = 0; i < NS; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<kHd64TmemCols>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;   // S / P at cols [0,64), O at [64,128)

  if (warp >= 4) {
    setmaxnreg_dec_32();
    if (warp == 4) {
      // =============================== TMA producer ===============================
      if (lane == 0 && n_kv > 0) {
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_4d(sQ, &tmQ, q_full, 0, q_off + qt * BM, h, tb);
        int ready_seg = -1;
        for (int j = 0; j < n_kv; ++j) {
          const int s = j % NS;
          const uint32_t ph = (j / NS) & 1;
          int kc = k_off + j * BN, vc = v_off + j * BN, kb = tb;
          if (p.kv_seg_len > 0) {
            const int seg = (j * BN) / p.kv_seg_len;
            kc = vc = j * BN - seg * p.kv_seg_len;
            kb = seg * p.B + b;
            if (p.seg_flags != nullptr && seg != ready_seg) {
              // gather fused into the launch: first tile of a segment -> has the peer copy of this (head group, segment) landed?
              const uint32_t* flag = p.seg_flags + (hk / p.seg_heads) * (p.Sk / p.kv_seg_len) + seg;
              const long long t0 = clock64();
              while (ld_acquire_sys_u32(flag) != p.seg_epoch) {
                __nanosleep(200);
                if (clock64() - t0 > (8ll << 30)) __trap();   // ~4 s: the copies never came; fail the launch instead of hanging
              }
              fence_proxy_async_all();
              ready_seg = seg;
            }
          }
          mbar_wait_wd(kv_empty + s, ph ^ 1);
          mbar_expect_tx(kv_full + s, K_TILE + V_TILE);
          tma_load_4d(sK + s * K_TILE, &tmK, kv_full + s, 0, kc, hk, kb);
          tma_load_4d(sV + s * V_TILE, &tmV, kv_full + s, vc, 0, hk, kb);
        }
      }
    } else if (warp == 5) {
      // =============================== MMA issuer (warp-uniform, one elected lane) ===============================
      if (n_kv > 0) {
        constexpr uint32_t idesc_qk = make_idesc(2, 1, 1, BM, BN);
        constexpr uint32_t idesc_pv = make_idesc(1, 0, 0, BM, D);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dQ = make_smem_desc<64>(smem_u32(sQ));
        const uint64_t dK0 = make_smem_desc<64>(smem_u32(sK));
        const uint64_t dV0 = make_smem_desc<64>(smem_u32(sV));
        auto issue_qk = [&](int t) {
          const uint64_t dK = dK0 + uint64_t(t % NS) * (K_TILE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 32; ++k) umma_i8_ss(tmem_u, dQ + 2 * k, dK + 2 * k, idesc_qk, k > 0);
          }
        };
        mbar_wait_wd(q_full, 0);
        mbar_wait_wd(kv_full + 0, 0);
        tc_fence_after();
        issue_qk(0);
        if (elect_one()) tc_commit(s_full);
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) mbar_wait_wd(kv_full + (j + 1) % NS, ((j + 1) / NS) & 1);
          mbar_wait_wd(p_full, j & 1);
          tc_fence_after();
          const uint64_t dV = dV0 + uint64_t(j % NS) * (V_TILE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BN / 32; ++k) umma_f8_ts(tmem_u + 64, tmem_u + 8 * k, dV + 2 * k, idesc_pv, (j > 0 || k > 0));
          }
          if (j + 1 < n_kv) issue_qk(j + 1);     // overwrites S / P after PV(j) has read P (in-order pipe)
          if (elect_one()) {
            tc_commit(s_full);
            tc_commit(kv_empty + j % NS);
          }
        }
      }
    }
  } else {
    // =============================== softmax / correction / epilogue ===============================
    setmaxnreg_inc_96();
    const int row = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_off;
    const uint32_t tO = tmem_base + lane_off + 64;
    const int q_row = qt * BM + row;

    int q_idx = (q_blk0 + qt) * p.q_mult;
    if (p.q_gran == SAB_GRAN_PER_WARP) q_idx += row >> 5;
    if (p.q_gran == SAB_GRAN_PER_THREAD) q_idx += (row >> 5) * 8 + (row & 7);
    const float* qs_base = p.q_scale + (varlen ? int64_t(h) : (int64_t(b) * p.Hq + h) * p.qs_stride_bh);
    const float* ks_base = p.k_scale + (varlen ? int64_t(hk) : (int64_t(b) * p.Hkv + hk) * p.ks_stride_bh);
    const float qss = qs_base[int64_t(q_idx) * p.qs_stride_idx] * p.sm_scale_log2;

    float m = kMaskValue;
    float d = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      float coef[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) coef[g] = 0.f;
      if (kKT && p.ks_vec4) {   // dense per-thread scales: the four scales of a key tile are one aligned 16-byte word
        const float4 k4 = *reinterpret_cast<const float4*>(ks_base + int64_t(k_blk0 + j) * 4);
        coef[0] = k4.x * qss; coef[NG > 1 ? 1 : 0] = k4.y * qss; coef[NG > 2 ? 2 : 0] = k4.z * qss; coef[NG - 1] = k4.w * qss;
      } else {
#pragma unroll
        for (int g = 0; g < NG; ++g) coef[g] = ks_base[int64_t((k_blk0 + j) * NG + g) * p.ks_stride_idx] * qss;
      }
      int limit = kv_len - j * BN;
      if (p.causal) limit = min(limit, p.causal_q_offset + q_row - j * BN + 1);
      const bool masked_tile = (kv_len - j * BN < BN) || (p.causal && (j + 1) * BN > p.causal_q_offset + qt * BM + 1);

      mbar_wait_wd(s_full, j & 1);
      tc_fence_after();
      uint32_t s[BN];
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        tmem_ld32(tS, lo);
        tmem_ld32(tS + 32, hi);
        tc_wait_ld();
      }

      auto tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        if constexpr (MASKED) {
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
          float c = float(v) * coef[g];
          if constexpr (MASKED) c = (v == kIntSentinel) ? kMaskValue : c;
          mx = fmaxf(mx, c);
        }
        // kLazy: the max moves only when a P of this tile would overflow e4m3 (it grew by more than 2^tau past the point where the
        // current maximum was placed), so the in-line O rescale below — ~70 of this issue-bound kernel's ~420 warp instructions per
        // tile when it runs — happens in well under 1 % of the tiles; else the reference's exact max (update_mdo, attn_utils.cuh:377-396)
        float m_new;
        if constexpr (kLazy) {
          const float m_true = fmaxf(m, mx - (kFp8Offset - kLazyTau));
          m_new = (m_true - m > kLazyTau) ? m_true : m;
        } else {
          m_new = fmaxf(m, mx - kFp8Offset);
        }
        float alpha = 1.0f;
        if (!kLazy || __any_sync(0xffffffffu, m_new != m)) {   // lazy max: rare after the first tiles (ex2(0) = 1 exactly)
          alpha = ex2_approx(m - m_new);
          d *= alpha;
        }
        m = m_new;

        // P = exp2(S*coef - m) -> e4m3 (bit-identical to the reference, see attn.cu), d += sum(P)
        uint64_t coef2[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) coef2[g] = pack_f2(coef[g], coef[g]);
        const uint64_t nm2 = pack_f2(-m_new, -m_new);
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
        tmem_st16(tS, pk);
        // correction: PV(j-1) retired before s_full(j) completed, so this row of O can be rescaled in place
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
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
        }
      };
      if (masked_tile) tile(std::true_type{});
      else tile(std::false_type{});

      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // ---- epilogue: dense outputs are staged in the idle K ring ([128 rows][128 B], 128-byte swizzle) and leave through one
    // TMA store per CTA (full lines, rows past Sq clipped by the map); packed varlen rows are stored directly
    const bool row_ok = q_row < q_len;
    auto out_row = [&]() {
      return reinterpret_cast<OutT*>(p.out) + int64_t(h) * p.o_stride_h + int64_t(q_off + q_row) * p.o_stride_s;
    };
    const bool use_tma = kTmaStoreEpilogue && p.o_tma != 0;
    static_assert(D * sizeof(OutT) == 128 && NS * K_TILE >= BM * 128, "staging tile is 128 rows x 128 bytes");
    const uint32_t stage_row = smem_u32(sK) + row * 128;
    const float* vs = p.v_scale ? p.v_scale + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D : nullptr;
    const float* vm = p.v_mean ? p.v_mean + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D : nullptr;
    if (n_kv > 0) {
      mbar_wait_wd(s_full, n_kv & 1);   // phase n_kv: PV(n_kv-1) accumulated
      tc_fence_after();
      const float inv = rcp_approx(d);
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tO + ch * 32, r);
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
        if (use_tma) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4)
            st_shared_v4(stage_row + (((ch * 4 + v4) ^ (row & 7)) << 4), o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
        } else if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(out_row() + ch * 32);
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) dst[v4] = make_uint4(o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
        }
      }
    } else if (use_tma) {
#pragma unroll
      for (int v4 = 0; v4 < D / 8; ++v4) st_shared_v4(stage_row + (v4 << 4), 0u, 0u, 0u, 0u);
    } else if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(out_row());
#pragma unroll
      for (int v4 = 0; v4 < D / 8; ++v4) dst[v4] = make_uint4(0, 0, 0, 0);
    }
    if (use_tma) {
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) {
        tma_store_4d(&p.o_map, sK, 0, q_row - row, blockIdx.y, blockIdx.z);
        tma_store_commit();
        tma_store_wait_read();
      }
    }
    if (p.lse != nullptr && row_ok) {
      const int64_t li = varlen ? (int64_t(h) * p.Sq + q_off + q_row) : ((int64_t(b) * p.Hq + h) * p.Sq + q_row);
      p.lse[li] = n_kv > 0 ? lg2_approx(d) + m : -INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<kHd64TmemCols>(tmem_base);
}

template <bool kKT, typename OutT, bool kLazy>
int launch_attn_hd64(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                     cudaStream_t stream) {
  constexpr int NS = 5;
  const size_t smem = size_t(BM) * 64 + size_t(NS) * 2 * BN * 64 + 256;   // 49.3 KB -> four CTAs per SM
  auto kern = sage_attn_hd64_kernel<kKT, OutT, kLazy>;
  static bool configured[64] = {};
  if (int st = ensure_dynamic_smem(kern, smem, configured)) return st;
  kern<<<grid, kHd64Threads, smem, stream>>>(tq, tk, tv, p);
  SAB_CUDA_OK(cudaGetLastError());
  return SAB_OK;
}

#define SAB_INST(KT, T, LZ) \
  template int launch_attn_hd64<KT, T, LZ>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const AttnParams&, dim3, cudaStream_t);
SAB_INST(true, __nv_bfloat16, true)
SAB_INST(true, __half, true)
SAB_INST(false, __nv_bfloat16, true)
SAB_INST(false, __half, true)
SAB_INST(true, __nv_bfloat16, false)
SAB_INST(true, __half, false)
SAB_INST(false, __nv_bfloat16, false)
SAB_INST(false, __half, false)
#undef SAB_INST

}  // namespace sab
