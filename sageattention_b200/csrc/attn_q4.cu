// Fused INT8-QK / FP8-PV attention for sm_100a (B200), head_dim 128: ONE CTA per SM, FOUR softmax warpgroups.
//
// attn_alt.cu (two CTAs per SM, two softmax warpgroups each) leaves the MUFU 31 % idle: with 256 TMEM columns per CTA the e4m3
// P(j) has to alias its own S buffer, so QK(j+2) cannot be issued before PV(j) has consumed P(j), and every warpgroup waits
// ~350+ cycles per tile for its next S (21 % of the softmax warps' time).  One CTA owns all 512 TMEM columns here:
//   S   [0,256)    four 64-column buffers, tile j in buffer j & 3, softmax by warpgroup j & 3
//   P   [256,320)  four 16-column e4m3 buffers of their own
//   O   [384,512)  fp32 accumulator
// so QK(j+4) only needs the owner of tile j to have READ S(j) (`s_free`, signalled in the middle of the exponentials), not
// PV(j): the next S of a warpgroup is ready long before it is needed.  Two issuing warps feed the tensor pipe independently —
// one for QK^T (waits s_free / K tiles, commits s_full), one for PV (waits p_full / V tiles, commits pv_done) — and K and V
// travel in separate rings with their own producers, because K(j+4) is consumed about four tiles before V(j).
// Everything on the softmax side is attn_alt.cu's: one thread per row and tile, lazy running max (SAB_ALT_TAU), the running max
// chained through shared memory (`m_full`), in-line O rescale (rare), partial row sums per warpgroup combined in the epilogue,
// TMA-store epilogue.
// 640 threads: warps 0-15 softmax (104 registers), 16 K/Q producer, 17 QK issuer + TMEM allocator, 18 PV issuer, 19 V producer
// (64 registers): 512 x 104 + 128 x 64 = 640 x 96, the CTA's register pool.
#include "attn_common.cuh"

namespace sab {

constexpr int kQ4Threads = 640;
constexpr uint32_t kQ4TmemCols = 512;
#ifndef SAB_ALT_TAU
#define SAB_ALT_TAU 4
#endif

__device__ __forceinline__ void setmaxnreg_inc_104() { asm volatile("setmaxnreg.inc.sync.aligned.u32 104;"); }
__device__ __forceinline__ void setmaxnreg_dec_64() { asm volatile("setmaxnreg.dec.sync.aligned.u32 64;"); }
__device__ __forceinline__ void q4_bar_sync_all() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the four softmax warpgroups

template <bool kKT, typename OutT>
__global__ void __launch_bounds__(kQ4Threads, 1)
sage_attn_q4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  constexpr int D = 128;
  constexpr uint32_t K_TILE = BN * D, V_TILE = D * BN, Q_BYTES = BM * D;
  constexpr int NK = 8, NV = 8;     // ring slots (64-key tiles)
  constexpr int NG = kKT ? 4 : 1;
  constexpr int NW = 4;             // softmax warpgroups = S / P buffers
  constexpr int OC = D / NW;        // O columns each warpgroup writes in the epilogue

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + NK * K_TILE;
  float* s_m = reinterpret_cast<float*>(sV + NV * V_TILE);     // [4 buffers][128 rows] running max m(j)
  float* s_x = s_m + NW * BM;                                   // [4 warpgroups][2][128] epilogue exchange: d, last max
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + 2 * NW * BM);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;            // [4] QK(t) retired: S(t) in buffer t & 3
  uint64_t* s_free = s_full + NW;         // [4] 128 arrivals: the owner of tile t has read S(t) for the last time
  uint64_t* p_full = s_free + NW;         // [4] 128 arrivals: P(t) stored (and O rescaled when the max moved)
  uint64_t* pv_done = p_full + NW;        // [4] PV(t) retired: P buffer t & 3 free, O holds tiles <= t
  uint64_t* m_full = pv_done + NW;        // [4] 128 arrivals: m(t) published in s_m[t & 3]
  uint64_t* k_full = m_full + NW;
  uint64_t* k_empty = k_full + NK;
  uint64_t* v_full = k_empty + NK;
  uint64_t* v_empty = v_full + NV;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(v_empty + NV);

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

  if (warp == 16 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < NW; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(s_free + i, 128);
      mbar_init(p_full + i, 128);
      mbar_init(pv_done + i, 1);
      mbar_init(m_full + i, 128);
    }
    for (int i = 0; i < NK; ++i) {
      mbar_init(k_full + i, 1);
      mbar_init(k_empty + i, 1);
    }
    for (int i = 0; i < NV; ++i) {
      mbar_init(v_full + i, 1);
      mbar_init(v_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc<kQ4TmemCols>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp >= 16) {
    setmaxnreg_dec_64();
    if (warp == 16 || warp == 19) {
      // =============================== TMA producers: warp 16 Q + K ring, warp 19 V ring ===============================
      const bool is_k = warp == 16;
      if (lane == 0 && n_kv > 0) {
        if (is_k) {
          mbar_expect_tx(q_full, Q_BYTES);
          tma_load_4d(sQ, &tmQ, q_full, 0, q_off + qt * BM, h, tb);
        }
        int ready_seg = -1;
        for (int j = 0; j < n_kv; ++j) {
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
          if (is_k) {
            const int s = j % NK;
            mbar_wait_wd(k_empty + s, ((j / NK) & 1) ^ 1);
            mbar_expect_tx(k_full + s, K_TILE);
            tma_load_4d(sK + s * K_TILE, &tmK, k_full + s, 0, kc, hk, kb);
          } else {
            const int s = j % NV;
            mbar_wait_wd(v_empty + s, ((j / NV) & 1) ^ 1);
            mbar_expect_tx(v_full + s, V_TILE);
            tma_load_4d(sV + s * V_TILE, &tmV, v_full + s, vc, 0, hk, kb);
          }
        }
      }
    } else if (warp == 17) {
      // =============================== QK^T issuer ===============================
      if (n_kv > 0) {   // whole warp runs the loop (uniform control flow); one elected lane issues
        constexpr uint32_t idesc_qk = make_idesc(2, 1, 1, BM, BN);   // s32 <- s8 x s8, 128 x 64
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dQ = make_smem_desc<128>(smem_u32(sQ));
        const uint64_t dK0 = make_smem_desc<128>(smem_u32(sK));
        mbar_wait_wd(q_full, 0);
        for (int t = 0; t < n_kv; ++t) {
          const int st = t % NK;
          // buffer t & 3 held S(t-4): its owner arrives on s_free once its last tcgen05.ld of that tile has completed
          if (t >= NW) mbar_wait_wd(s_free + (t & 3), uint32_t((t >> 2) - 1) & 1u);
          mbar_wait_wd(k_full + st, (t / NK) & 1);
          tc_fence_after();
          const uint64_t dK = dK0 + uint64_t(st) * (K_TILE >> 4);
          const uint32_t tS = tmem_u + (t & 3) * BN;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 32; ++k) umma_i8_ss(tS, dQ + 2 * k, dK + 2 * k, idesc_qk, k > 0);
            tc_commit(s_full + (t & 3));
            tc_commit(k_empty + st);
          }
          __syncwarp();
        }
      }
    } else {
      // =============================== PV issuer (warp 18) ===============================
      if (n_kv > 0) {
        constexpr uint32_t idesc_pv = make_idesc(1, 0, 0, BM, D);    // f32 <- e4m3 x e4m3, 128 x 128
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dV0 = make_smem_desc<64>(smem_u32(sV));
        for (int t = 0; t < n_kv; ++t) {
          const int st = t % NV;
          mbar_wait_wd(v_full + st, (t / NV) & 1);
          mbar_wait_wd(p_full + (t & 3), uint32_t(t >> 2) & 1u);
          tc_fence_after();
          const uint64_t dV = dV0 + uint64_t(st) * (V_TILE >> 4);
          const uint32_t tP = tmem_u + 256 + (t & 3) * 16;   // P(t): 16 columns (32 keys per 8 columns)
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BN / 32; ++k) umma_f8_ts(tmem_u + 384, tP + 8 * k, dV + 2 * k, idesc_pv, (t > 0 || k > 0));
            tc_commit(pv_done + (t & 3));
            tc_commit(v_empty + st);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== softmax of the tiles j = wg, wg+4, ... / in-line correction / epilogue ===============================
    setmaxnreg_inc_104();
    const int wg = warp >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + 384;
    const int q_row = qt * BM + row;

    int q_idx = (q_blk0 + qt) * p.q_mult;
    if (p.q_gran == SAB_GRAN_PER_WARP) q_idx += row >> 5;
    if (p.q_gran == SAB_GRAN_PER_THREAD) q_idx += (row >> 5) * 8 + (row & 7);
    const float* qs_base = p.q_scale + (varlen ? int64_t(h) : (int64_t(b) * p.Hq + h) * p.qs_stride_bh);
    const float* ks_base = p.k_scale + (varlen ? int64_t(hk) : (int64_t(b) * p.Hkv + hk) * p.ks_stride_bh);
    const float qss = qs_base[int64_t(q_idx) * p.qs_stride_idx] * p.sm_scale_log2;

    // Parity waits (completion c of a per-buffer barrier belongs to tile 4c + buffer).  Sound only when the barrier is neither a
    // full cycle ahead of nor behind the waiter:
    //   * s_full[j&3], pv_done[j&3] by the owner of tile j: it follows its own buffer completion by completion;
    //   * m_full[(j-1)&3] and pv_done[(j-1)&3] by the owner of tile j: tile j+3 of that buffer needs m(j) / P(j) from this
    //     warpgroup first (not ahead), and this warpgroup's tile j-4 already needed m(j-5) / waited for PV(j-4), which in-order
    //     retirement puts after PV(j-5) (not behind).
    float m_own = kMaskValue;   // m(j) of this warpgroup's latest tile: the reference its partial sum d is relative to
    float d = 0.f;              // sum of P over THIS warpgroup's tiles, relative to m_own

    for (int j = wg; j < n_kv; j += NW) {
      const int bf = j & 3;     // == wg
      const uint32_t tS = tmem_base + lane_off + bf * BN;
      const uint32_t tP = tmem_base + lane_off + 256 + bf * 16;
      float coef[NG];
      if constexpr (kKT) {
        if (p.ks_vec4) {   // dense: the four per-thread scales of a key tile are one aligned 16-byte word
          const float4 k4 = *reinterpret_cast<const float4*>(ks_base + int64_t(k_blk0 + j) * 4);
          coef[0] = k4.x * qss; coef[1] = k4.y * qss; coef[2] = k4.z * qss; coef[NG - 1] = k4.w * qss;
        } else {
#pragma unroll
          for (int g = 0; g < NG; ++g) coef[g] = ks_base[int64_t((k_blk0 + j) * NG + g) * p.ks_stride_idx] * qss;
        }
      } else {
        coef[0] = ks_base[int64_t(k_blk0 + j) * p.ks_stride_idx] * qss;
      }
      int limit = kv_len - j * BN;
      if (p.causal) limit = min(limit, p.causal_q_offset + q_row - j * BN + 1);
      const bool masked_tile = (kv_len - j * BN < BN) || (p.causal && (j + 1) * BN > p.causal_q_offset + qt * BM + 1);

      mbar_wait_wd(s_full + bf, uint32_t(j >> 2) & 1u);
      tc_fence_after();

      auto tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // ---- pass 1: row max, streamed in two 32-column loads
        int pm[4] = {kIntSentinel, kIntSentinel, kIntSentinel, kIntSentinel};
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t s[32];
          tmem_ld32(tS + 32 * hf, s);
          tc_wait_ld();
          if constexpr (MASKED) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (32 * hf + i >= limit) s[i] = uint32_t(kIntSentinel);
          }
#pragma unroll
          for (int i8 = 0; i8 < 32; i8 += 8) {
#pragma unroll
            for (int g = 0; g < 4; ++g) pm[g] = __vimax3_s32(pm[g], int(s[i8 + 2 * g]), int(s[i8 + 2 * g + 1]));
          }
        }
        float mx = kMaskValue;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int v = kKT ? pm[g] : max(max(pm[0], pm[1]), max(pm[2], pm[3]));
          float c = float(v) * coef[g];
          if constexpr (MASKED) c = (v == kIntSentinel) ? kMaskValue : c;
          mx = fmaxf(mx, c);
        }
        // ---- the running max: m(j-1) comes from warpgroup (j-1) & 3, which published it right after ITS row max.  Slot reuse:
        //      I overwrite slot j & 3 (m(j-4)), last read by the owner of tile j-3 before it published m(j-3); m(j-1), which I
        //      wait for here, came after m(j-2), after m(j-3).
        float m_prev = kMaskValue;
        if (j > 0) {
          mbar_wait_wd(m_full + ((j - 1) & 3), uint32_t((j - 1) >> 2) & 1u);
          m_prev = s_m[((j - 1) & 3) * BM + row];
        }
#if SAB_ALT_TAU > 0
        const float m_true = fmaxf(m_prev, mx - (kFp8Offset - float(SAB_ALT_TAU)));
        const float m_new = (m_true - m_prev > float(SAB_ALT_TAU)) ? m_true : m_prev;
#else
        const float m_new = fmaxf(m_prev, mx - kFp8Offset);   // update_mdo, attn_utils.cuh:377-396
#endif
        s_m[bf * BM + row] = m_new;
        mbar_arrive(m_full + bf);
        float alpha_o = 1.0f;                                   // rescale of O before PV(j): consecutive tiles
        if (__any_sync(0xffffffffu, (m_new != m_prev) | (m_new != m_own))) {   // lazy max: rare after the first tiles (ex2(0) = 1)
          alpha_o = ex2_approx(m_prev - m_new);
          d *= ex2_approx(m_own - m_new);                        // my partial sum: relative to my previous tile (j-4)
        }
        m_own = m_new;

        uint64_t coef2[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) coef2[g] = pack_f2(coef[g], coef[g]);
        const uint64_t nm2 = pack_f2(-m_new, -m_new);
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
        // ---- pass 2: exponentials, again 32 columns at a time; P goes to its own buffer (free once PV(j-4) has retired)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t s[32];
          tmem_ld32(tS + 32 * hf, s);
          tc_wait_ld();
          if (hf == 1) {             // last read of S(j): the QK issuer may overwrite the buffer with S(j+4)
            tc_fence_before();
            mbar_arrive(s_free + bf);
          }
          uint32_t pk[8];
#pragma unroll
          for (int w = 0; w < 8; ++w) {
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
                e[u] = (32 * hf + i < limit) ? e[u] : 0.f;
                e[u + 1] = (32 * hf + i + 1 < limit) ? e[u + 1] : 0.f;
              }
              acc[(w & 1) * 2 + (u >> 1)] = fadd2(acc[(w & 1) * 2 + (u >> 1)], pack_f2(e[u], e[u + 1]));
            }
            pk[w] = pack_e4m3x4(e[0], e[1], e[2], e[3]);
          }
          if (hf == 0 && j >= NW) {   // P buffer j & 3 still holds P(j-4) until PV(j-4) has retired
            mbar_wait_wd(pv_done + bf, uint32_t((j >> 2) - 1) & 1u);
            tc_fence_after();
          }
          tmem_st8(tP + 8 * hf, pk);
        }
        {
          float a0, a1, a2, a3;
          unpack_f2(fadd2(acc[0], acc[1]), a0, a1);
          unpack_f2(fadd2(acc[2], acc[3]), a2, a3);
          d += (a0 + a1) + (a2 + a3);
        }

        // ---- in-line correction of this row of O, when any row of the warp moved its max: needs PV(j-1) accumulated
        if (j > 0 && __any_sync(0xffffffffu, alpha_o != 1.0f)) {
          mbar_wait_wd(pv_done + ((j - 1) & 3), uint32_t((j - 1) >> 2) & 1u);
          tc_fence_after();
          const uint64_t alpha2 = pack_f2(alpha_o, alpha_o);
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
      mbar_arrive(p_full + bf);
    }

    // ---- epilogue: combine the four partial sums relative to the final max, then each warpgroup writes OC columns
    s_x[(wg * 2 + 0) * BM + row] = d;
    s_x[(wg * 2 + 1) * BM + row] = m_own;
    if (n_kv > 0 && wg == ((n_kv - 1) & 3)) {   // owner of the last tile: O is final once PV(n_kv-1) has retired
      mbar_wait_wd(pv_done + ((n_kv - 1) & 3), uint32_t((n_kv - 1) >> 2) & 1u);
      tc_fence_after();
    }
    tc_fence_before();
    q4_bar_sync_all();
    tc_fence_after();
    float m_fin = kMaskValue;
#pragma unroll
    for (int w = 0; w < NW; ++w) m_fin = fmaxf(m_fin, s_x[(w * 2 + 1) * BM + row]);
    d = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) d += s_x[(w * 2 + 0) * BM + row] * ex2_approx(s_x[(w * 2 + 1) * BM + row] - m_fin);

    const bool row_ok = q_row < q_len;
    auto out_row = [&]() {   // packed varlen rows only
      return reinterpret_cast<OutT*>(p.out) + int64_t(h) * p.o_stride_h + int64_t(q_off + q_row) * p.o_stride_s + wg * OC;
    };
    // Dense outputs leave through TMA: a PAIR of warpgroups stages its [128 rows][128 B] half of the tile in the idle K ring
    // (128-byte swizzle; warpgroup 2i + 1 fills the upper four 16-byte chunks of a row) and one thread issues the bulk store.
    const bool use_tma = kTmaStoreEpilogue && p.o_tma != 0;
    static_assert(OC * sizeof(OutT) == 64 && NK * K_TILE >= 2 * BM * 128, "staging: 128 rows x 128 bytes per warpgroup pair");
    const int pair = wg >> 1;
    uint8_t* stage = sK + pair * (BM * 128);
    const uint32_t stage_row = smem_u32(stage) + row * 128;
    const float* vs = p.v_scale ? p.v_scale + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + wg * OC : nullptr;
    const float* vm = p.v_mean ? p.v_mean + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + wg * OC : nullptr;
    uint32_t o16[16];
    if (n_kv > 0) {
      const float inv = rcp_approx(d);
      uint32_t r[32];
      tmem_ld32(tO + wg * OC, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float a = __uint_as_float(r[i]) * inv, c = __uint_as_float(r[i + 1]) * inv;
        if (vs) {
          a *= vs[i];
          c *= vs[i + 1];
        }
        if (vm) {
          a += vm[i];
          c += vm[i + 1];
        }
        o16[i / 2] = pack2<OutT>(a, c);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) o16[i] = 0u;
    }
    if (use_tma) {
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4)
        st_shared_v4(stage_row + ((((wg & 1) * 4 + v4) ^ (row & 7)) << 4), o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
      fence_proxy_async_smem();                                        // generic-proxy stores -> visible to the TMA engine
      if (pair == 0) asm volatile("bar.sync 2, 256;" ::: "memory");    // this pair's half tile is staged
      else asm volatile("bar.sync 3, 256;" ::: "memory");
      if ((threadIdx.x & 255) == 0) {
        tma_store_4d(&p.o_map, stage, pair * 128, q_row - row, blockIdx.y, blockIdx.z);
        tma_store_commit();
        tma_store_wait_read();                                         // the staging buffer must outlive the read
      }
    } else if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(out_row());
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) dst[v4] = make_uint4(o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
    }
    if (p.lse != nullptr && row_ok && wg == 0) {
      const int64_t li = varlen ? (int64_t(h) * p.Sq + q_off + q_row) : ((int64_t(b) * p.Hq + h) * p.Sq + q_row);
      p.lse[li] = n_kv > 0 ? lg2_approx(d) + m_fin : -INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<kQ4TmemCols>(tmem_base);
}

template <bool kKT, typename OutT>
int launch_attn_q4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                   cudaStream_t stream) {
  // Q + 8 K + 8 V tiles + running-max / epilogue exchange + barriers: 151 KB -> one CTA per SM (it owns all 512 TMEM columns)
  const size_t smem = size_t(BM) * 128 + size_t(16) * BN * 128 + 12 * BM * sizeof(float) + 1024;
  auto kern = sage_attn_q4_kernel<kKT, OutT>;
  static bool configured[64] = {};
  if (int st = ensure_dynamic_smem(kern, smem, configured)) return st;
  kern<<<grid, kQ4Threads, smem, stream>>>(tq, tk, tv, p);
  SAB_CUDA_OK(cudaGetLastError());
  return SAB_OK;
}

#define SAB_INST(KT, T) \
  template int launch_attn_q4<KT, T>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const AttnParams&, dim3, cudaStream_t);
SAB_INST(true, __nv_bfloat16)
SAB_INST(true, __half)
SAB_INST(false, __nv_bfloat16)
SAB_INST(false, __half)
#undef SAB_INST

}  // namespace sab
