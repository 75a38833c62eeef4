// Fused INT8-QK / FP8-PV attention for sm_100a (B200) — the PRODUCT kernel of the INT8+FP8 path at head_dim 128.
//
// Tile pipeline shared with attn.cu: one 128-row Q tile per CTA, 64-key tiles (the reference's CTA_K), S double-buffered in TMEM,
// P(j) (e4m3) written over its S buffer and fed to the PV MMA from TMEM, O fp32 in TMEM, two CTAs per SM, in-order tensor pipe
// QK0 QK1 | PV0 QK2 | PV1 QK3 ...  What differs is the softmax side, shaped by the B200 measurements of round 2:
//   * MUFU.EX2 runs at 16 lanes/clk/SM (tools/microbench/mufu_rate.cu: 8.0 cycles per warp instruction and sub-partition), so a
//     128x64 tile costs 512 MUFU cycles against 256 tensor cycles: the exponentials are the roofline of an 8-bit attention kernel,
//     and what the exact kernel of attn.cu lost (MUFU 60 % busy) was the serial per-tile chain of its ONE softmax warp per 32 rows.
//   * TWO softmax warpgroups per CTA on ALTERNATE key tiles (warpgroup 0 even j, warpgroup 1 odd j; warps w and w+4 reach the same
//     TMEM lane quadrant): both S buffers are consumed at once and every scheduler holds four independent exp chains, one thread per
//     row and tile as before.  The online softmax couples consecutive tiles only through the running max:
//       m(j) from m(j-1)                            one float per row through shared memory + one mbarrier (m_full)
//       O *= 2^(m(j-1) - m(j)) before PV(j)          done in-line by the thread that owns tile j (after step(j-1) retired)
//       d                                            each warpgroup keeps the sum of ITS tiles relative to its last max; the two
//                                                    partial sums are combined once in the epilogue
//   * LAZY running max (SAB_ALT_TAU = 4 binades): the max moves only when an element of the tile would overflow e4m3 (P > 448);
//     when it moves, the new row maximum is placed tau binades below 448 (exponent offset 8.807 - tau), so the in-line O rescale —
//     128 TMEM columns read and written per row — runs in ~0.2 % of the warp-tiles instead of ~80 % with the reference's exact max.
//     P keeps e4m3's relative precision (a floating format); only the tail below 2^-9 is cut 2^tau earlier.  P is therefore a
//     different rounding realisation than the reference kernel's (same accuracy against exact attention, tests/test_gpu_parity.py);
//     SAB_ATTN_KERNEL=exact selects the exact-max kernel of attn.cu.
// 384 threads, registers 96 / 96 / 48 via setmaxnreg; no correction warpgroup.  Epilogue: each warpgroup stages its 64 output columns
// in the idle K ring (128-byte swizzle) and one thread issues a TMA bulk tensor store (dense outputs; packed varlen rows are stored directly).
// Measured (B200, hd128 S=8192 non-causal, kernel only): 1307 TFLOP/s at tau = 0, 1438 at tau = 3, 1452 at tau = 4 (exact kernel: 1277);
// 1492 with the TMA-store epilogue, one float4 K-scale load per tile and the rescale ex2 skipped when the max did not move
// (profiles/r02_bench_n1.json: 0.425 of the 8-bit denominator, XU pipe 70 % busy).
#include "attn_common.cuh"

namespace sab {

constexpr int kAltThreads = 384;   // warpgroups: 0 = softmax of even key tiles, 1 = softmax of odd key tiles, 2 = TMA / MMA / 2 idle
#ifndef SAB_ALT_TAU
#define SAB_ALT_TAU 4   // lazy-max threshold in binades (0: the reference's exact running max; measured on B200: 1307 / 1438 / 1452 TFLOP/s at tau 0 / 3 / 4)
#endif

// 128 x (96 + 96 + 48) = 30720 = the CTA's register pool (80 x 384)
__device__ __forceinline__ void setmaxnreg_inc_96a() { asm volatile("setmaxnreg.inc.sync.aligned.u32 96;"); }
// both softmax warpgroups (256 threads); barrier 0 is __syncthreads
__device__ __forceinline__ void alt_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int D, bool kKT, typename OutT>
__global__ void __launch_bounds__(kAltThreads, 2)
sage_attn_alt_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  constexpr uint32_t K_TILE = BN * D;
  constexpr uint32_t V_TILE = D * BN;
  constexpr int NS = (D == 128) ? 5 : 10;
  constexpr int SWQK = (D == 128) ? 128 : 64;
  constexpr uint32_t Q_BYTES = BM * D;
  constexpr int NG = kKT ? 4 : 1;
  constexpr int OC = D / 2;         // O columns each warpgroup writes in the epilogue

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + NS * K_TILE;
  float* s_m = reinterpret_cast<float*>(sV + NS * V_TILE);     // [2 tile parity][128 rows] running max m(j)
  float* s_x = s_m + 2 * BM;                                    // [2 warpgroups][2][128] epilogue exchange: d, last max
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + 4 * BM);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;    // [2] step(t) retired: S(t+2) ready in buffer t&1 AND PV(t) accumulated into O
  uint64_t* p_full = bars + 3;    // [2] 128 arrivals: the warpgroup of tile j stored P(j) and rescaled O
  uint64_t* m_full = bars + 5;    // [2] 128 arrivals: m(j) published in s_m[j&1]
  uint64_t* kv_full = bars + 7;
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
      mbar_init(p_full + i, 128);
      mbar_init(m_full + i, 128);
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

  if (warp >= 8) {
    setmaxnreg_dec_48();
    if (warp == 8) {
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
          const uint32_t tP = tmem_u + (j & 1) * BN;   // P(j): 16 columns over the start of its S buffer (32 keys per 8 columns)
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BN / 32; ++k) umma_f8_ts(tmem_u + 128, tP + 8 * k, dV + 2 * k, idesc_pv, (j > 0 || k > 0));
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
    // =============================== softmax of the tiles j = wg, wg+2, ... / in-line correction / epilogue ===============================
    setmaxnreg_inc_96a();
    const int wg = warp >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + 128;
    const int q_row = qt * BM + row;

    int q_idx = (q_blk0 + qt) * p.q_mult;
    if (p.q_gran == SAB_GRAN_PER_WARP) q_idx += row >> 5;
    if (p.q_gran == SAB_GRAN_PER_THREAD) q_idx += (row >> 5) * 8 + (row & 7);
    const float* qs_base = p.q_scale + (varlen ? int64_t(h) : (int64_t(b) * p.Hq + h) * p.qs_stride_bh);
    const float* ks_base = p.k_scale + (varlen ? int64_t(hk) : (int64_t(b) * p.Hkv + hk) * p.ks_stride_bh);
    const float qss = qs_base[int64_t(q_idx) * p.qs_stride_idx] * p.sm_scale_log2;

    // Parity waits on the two s_full barriers (completion c of s_full[i]: c = 0 the prologue commit of QK(i), then the
    // retirement of step(2(c-1)+i)).  A parity wait is only sound when the barrier is neither a full cycle ahead of nor
    // behind the waiter (tests/test_alt_protocol_model.py):
    //   * S(j) by the owner of tile j: it follows s_full[j&1] completion by completion;
    //   * step(j-1) by the owner of tile j (correction): at that point step(j-3) has retired (S(j) was ready) and step(j+1)
    //     cannot have (the in-order MMA warp still waits for this warpgroup's p_full(j)), so the barrier is at most one
    //     completion away;
    //   * step(n_kv-1) in the epilogue: ONLY by the owner of the last tile; the other warpgroup — which may not have touched
    //     that barrier for a while, or at all when n_kv == 1 — learns it through the named barrier both meet at.
    auto wait_S_ready = [&](int t) { mbar_wait_wd(s_full + (t & 1), uint32_t(t >> 1) & 1u); };
    auto wait_step_retired = [&](int t) { mbar_wait_wd(s_full + (t & 1), uint32_t((t >> 1) + 1) & 1u); };

    float m_own = kMaskValue;   // m(j) of this warpgroup's latest tile: the reference its partial sum d is relative to
    float d = 0.f;              // sum of P over THIS warpgroup's tiles, relative to m_own

    for (int j = wg; j < n_kv; j += 2) {
      const uint32_t tS = tmem_base + lane_off + (j & 1) * BN;
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

      wait_S_ready(j);
      tc_fence_after();

      auto tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // ---- pass 1: row max, streamed in two 32-column loads (a 64-column row plus the packed P does not fit the 96
        //      registers of this warpgroup; TMEM reads are cheap and the other three softmax warps of the scheduler cover
        //      the load latency).  pm: per scale group (kKT) or four independent chains.
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
        // ---- the running max: m(j-1) comes from the other warpgroup (slot (j-1)&1; it published it right after ITS row max).
        //      Slot reuse: I overwrite slot j&1, last read by the other warpgroup at tile j-1 — which it finished reading
        //      before it published m(j-1), which I wait for here.
        float m_prev = kMaskValue;
        if (j > 0) {
          mbar_wait_wd(m_full + ((j - 1) & 1), uint32_t((j - 1) >> 1) & 1u);
          m_prev = s_m[((j - 1) & 1) * BM + row];
        }
#if SAB_ALT_TAU > 0
        const float m_true = fmaxf(m_prev, mx - (kFp8Offset - float(SAB_ALT_TAU)));
        const float m_new = (m_true - m_prev > float(SAB_ALT_TAU)) ? m_true : m_prev;
#else
        const float m_new = fmaxf(m_prev, mx - kFp8Offset);   // update_mdo, attn_utils.cuh:377-396
#endif
        s_m[(j & 1) * BM + row] = m_new;
        mbar_arrive(m_full + (j & 1));
        float alpha_o = 1.0f;                                   // rescale of O before PV(j): consecutive tiles
        if (__any_sync(0xffffffffu, (m_new != m_prev) | (m_new != m_own))) {   // lazy max: rare after the first tiles (ex2(0) = 1)
          alpha_o = ex2_approx(m_prev - m_new);
          d *= ex2_approx(m_own - m_new);                        // my partial sum: relative to my previous tile (j-2)
        }
        m_own = m_new;

        uint64_t coef2[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) coef2[g] = pack_f2(coef[g], coef[g]);
        const uint64_t nm2 = pack_f2(-m_new, -m_new);
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
        // ---- pass 2: exponentials, again 32 columns at a time; P half hf (8 columns) overwrites S columns [8 hf, 8 hf + 8),
        //      which both passes have consumed by then (half 0 is read before the first store, half 1 lives at [32,64))
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t s[32];
          tmem_ld32(tS + 32 * hf, s);
          tc_wait_ld();
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
#ifdef SAB_POLY_EXP_PAIRS
              if (((i >> 1) & 3) < SAB_POLY_EXP_PAIRS) {
                ex2_poly2(y0, y1, e[u], e[u + 1]);
              } else {
                e[u] = ex2_approx(y0);
                e[u + 1] = ex2_approx(y1);
              }
#else
              e[u] = ex2_approx(y0);
              e[u + 1] = ex2_approx(y1);
#endif
              if constexpr (MASKED) {
                e[u] = (32 * hf + i < limit) ? e[u] : 0.f;
                e[u + 1] = (32 * hf + i + 1 < limit) ? e[u + 1] : 0.f;
              }
              acc[(w & 1) * 2 + (u >> 1)] = fadd2(acc[(w & 1) * 2 + (u >> 1)], pack_f2(e[u], e[u + 1]));
            }
            pk[w] = pack_e4m3x4(e[0], e[1], e[2], e[3]);
          }
          tmem_st8(tS + 8 * hf, pk);
        }
        {
          float a0, a1, a2, a3;
          unpack_f2(fadd2(acc[0], acc[1]), a0, a1);
          unpack_f2(fadd2(acc[2], acc[3]), a2, a3);
          d += (a0 + a1) + (a2 + a3);
        }

        // ---- in-line correction of this row of O, when any row of the warp moved its max: needs PV(j-1) accumulated, i.e.
        //      step(j-1) retired — the phase of s_full[(j+1)&1] the other warpgroup waits for to start tile j+1.  No cycle:
        //      step(j-1) needs p_full(j-1) from the other warpgroup, whose own correction waits for step(j-2), which needs
        //      p_full(j-2) — signalled by this warpgroup one iteration ago.
        if (j > 0 && __any_sync(0xffffffffu, alpha_o != 1.0f)) {
          wait_step_retired(j - 1);
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
      mbar_arrive(p_full + (j & 1));
    }

    // ---- epilogue: combine the two partial sums relative to the final max m(n_kv-1), then each warpgroup writes OC columns
    s_x[(wg * 2 + 0) * BM + row] = d;
    s_x[(wg * 2 + 1) * BM + row] = m_own;
    if (n_kv > 0 && wg == ((n_kv - 1) & 1)) {   // owner of the last tile: O is final once step(n_kv-1) retired
      wait_step_retired(n_kv - 1);
      tc_fence_after();
    }
    tc_fence_before();
    alt_bar_sync();
    tc_fence_after();
    const float d_o = s_x[((wg ^ 1) * 2 + 0) * BM + row], m_o = s_x[((wg ^ 1) * 2 + 1) * BM + row];
    const float m_fin = fmaxf(m_own, m_o);     // the running max is non-decreasing, so the later tile's value is the larger
    d = d * ex2_approx(m_own - m_fin) + d_o * ex2_approx(m_o - m_fin);

    const uint32_t tOh = tO + wg * OC;
    const bool row_ok = q_row < q_len;
    auto out_row = [&]() {   // packed varlen rows only
      return reinterpret_cast<OutT*>(p.out) + int64_t(h) * p.o_stride_h + int64_t(q_off + q_row) * p.o_stride_s + wg * OC;
    };
    // Dense outputs leave through TMA: each warpgroup stages its [128 rows][128 B] half of the tile in the (now idle) K ring
    // with the 128-byte swizzle (conflict-free 16-byte shared stores) and one thread issues a bulk tensor store, which writes
    // full lines and clips the rows past Sq.  Packed varlen outputs (per-sequence clipping) keep the direct row stores.
    const bool use_tma = kTmaStoreEpilogue && p.o_tma != 0;
    static_assert(OC * sizeof(OutT) == 128 && NS * K_TILE >= 2 * BM * 128, "staging tile is 128 rows x 128 bytes per warpgroup");
    uint8_t* stage = sK + wg * (BM * 128);
    const uint32_t stage_row = smem_u32(stage) + row * 128;
    const float* vs = p.v_scale ? p.v_scale + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + wg * OC : nullptr;
    const float* vm = p.v_mean ? p.v_mean + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + wg * OC : nullptr;
    if (n_kv > 0) {
      const float inv = rcp_approx(d);   // O is final: the owner of the last tile saw step(n_kv-1) retire before the barrier above
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
      for (int v4 = 0; v4 < OC / 8; ++v4) st_shared_v4(stage_row + (v4 << 4), 0u, 0u, 0u, 0u);
    } else if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(out_row());
#pragma unroll
      for (int v4 = 0; v4 < OC / 8; ++v4) dst[v4] = make_uint4(0, 0, 0, 0);
    }
    if (use_tma) {
      fence_proxy_async_smem();                                        // generic-proxy stores -> visible to the TMA engine
      if (wg == 0) asm volatile("bar.sync 2, 128;" ::: "memory");      // this warpgroup's half tile is staged
      else asm volatile("bar.sync 3, 128;" ::: "memory");
      if ((threadIdx.x & 127) == 0) {
        tma_store_4d(&p.o_map, stage, wg * 128, q_row - row, blockIdx.y, blockIdx.z);
        tma_store_commit();
        tma_store_wait_read();                                         // the staging buffer must outlive the read
      }
    }
    if (p.lse != nullptr && row_ok && wg == 0) {
      const int64_t li = varlen ? (int64_t(h) * p.Sq + q_off + q_row) : ((int64_t(b) * p.Hq + h) * p.Sq + q_row);
      p.lse[li] = n_kv > 0 ? lg2_approx(d) + m_fin : -INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<kTmemCols>(tmem_base);
}

template <int D, bool kKT, typename OutT>
int launch_attn_alt(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                    cudaStream_t stream) {
  constexpr int NS = (D == 128) ? 5 : 10;
  size_t smem = size_t(BM) * D + size_t(NS) * 2 * BN * D + 6 * BM * sizeof(float) + 512;
  if (smem < 80 * 1024) smem = 80 * 1024;   // keep it at two CTAs per SM (TMEM: 2 x 256 columns)
  auto kern = sage_attn_alt_kernel<D, kKT, OutT>;
  static bool configured[64] = {};
  if (int st = ensure_dynamic_smem(kern, smem, configured)) return st;
  kern<<<grid, kAltThreads, smem, stream>>>(tq, tk, tv, p);
  SAB_CUDA_OK(cudaGetLastError());
  return SAB_OK;
}

#define SAB_INST(D, KT, T) \
  template int launch_attn_alt<D, KT, T>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const AttnParams&, dim3, cudaStream_t);
SAB_INST(128, true, __nv_bfloat16)
SAB_INST(128, true, __half)
SAB_INST(128, false, __nv_bfloat16)
SAB_INST(128, false, __half)
#undef SAB_INST

}  // namespace sab
