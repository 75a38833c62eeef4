// Fused INT8-QK / FP8-PV attention for sm_100a (B200) — the product kernel of the INT8+FP8 path ("lazy" softmax).
//
// Same tile pipeline as attn.cu (one 128-row Q tile per CTA, 64-key tiles = the reference's CTA_K, S double-buffered in
// TMEM, P(j) written over its S buffer and fed to the PV MMA from TMEM, O fp32 in TMEM, two CTAs per SM, in-order tensor
// pipe QK0 QK1 | PV0 QK2 | PV1 QK3 ...), but the softmax is restructured around what the B200 measurements showed
// (DESIGN.md section 4.3): with 8-bit operands a tile costs 256 tensor cycles against 512 MUFU cycles, and the exact kernel
// kept the MUFU only 60 % busy because every tile ran a serial chain  wait -> TMEM load -> row max -> alpha (MUFU) ->
// publish -> 64 exponentials -> P store -> hand-off  (~1400 cycles) in each softmax warp.  Here:
//   * LAZY running max (threshold tau = 4 binades): the max moves only when an element of the tile would overflow e4m3
//     (P > 448); when it moves, the new row maximum is placed tau binades below 448 (exponent offset 8.807 - tau), so it
//     rarely moves again.  P keeps e4m3's relative precision (a float format); only the tail below 2^-9 is cut 2^tau earlier.
//   * SPECULATIVE single pass: a tile is first evaluated against the CURRENT max — exponentials, row sum, e4m3 packing and
//     the integer row maximum (DPX 3-input max, off the critical path) in ONE loop with no dependency on the maximum.  The
//     check "would any P exceed 448?" is one warp vote at the end; in the ~1 % of warp-tiles where it fails, and in masked
//     tiles / the first tile, the classic two-step tile runs instead (max -> alpha -> exponentials) and the same thread
//     rescales its row of O in TMEM in-line (after step(j-1) retired).  There is no correction warpgroup and no alpha
//     hand-off: 256 threads (warps 0-3 softmax / epilogue, one thread per Q row; warp 4 TMA; warp 5 MMA + TMEM allocator).
//   * S(j+1) is prefetched from TMEM into a second register set INSIDE the exponential loop of tile j (its mbarrier
//     round trip and the tcgen05.ld latency hide under queued MUFU work); registers 208 / 48 via setmaxnreg.
// Numerics: m, P and d are relative to the lazy max, so P is no longer bit-identical to the reference kernel's (same e4m3
// rounding of a 2^k-shifted value); O and the LSE agree with the reference to its own quantisation noise (tests state the
// bound).  The exact-max kernel of attn.cu stays selectable (SAB_ATTN_KERNEL=exact) and serves the debug dumps.
#include "attn_common.cuh"

namespace sab {

constexpr int kLazyThreads = 256;
#ifndef SAB_LAZY_TAU
#define SAB_LAZY_TAU 4
#endif
#ifndef SAB_POLY_EXP_PAIRS
#define SAB_POLY_EXP_PAIRS 0
#endif
// A/B switches of the softmax loop (measured on B200)
#ifndef SAB_LZ_KSPRE      // K dequant scales of tile j+1 fetched during tile j
#define SAB_LZ_KSPRE 0
#endif
#ifndef SAB_LZ_TESTWAIT   // non-blocking test of s_full(j+1) a few iterations before the prefetch point
#define SAB_LZ_TESTWAIT 0
#endif
#ifndef SAB_LZ_ITHR       // overflow check on integer thresholds instead of the scaled float maximum
#define SAB_LZ_ITHR 0
#endif
#ifndef SAB_LZ_DEFER      // hand-off of a speculative tile's P deferred into the next tile's loop
#define SAB_LZ_DEFER 0
#endif
#ifndef SAB_LZ_PREFETCH_AT   // exp-loop iteration (4 elements each) at which S(j+1) is prefetched
#define SAB_LZ_PREFETCH_AT (SAB_LZ_DEFER ? 12 : 4)
#endif

// launch_bounds(256, 2) -> 128 registers per thread at launch = a 32768-register pool per CTA, re-split 208 / 48 (multiples of 16, see DESIGN.md on setmaxnreg granularity)
__device__ __forceinline__ void setmaxnreg_inc_208() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;"); }
__device__ __forceinline__ void setmaxnreg_dec_48b() { asm volatile("setmaxnreg.dec.sync.aligned.u32 48;"); }

// The registers written by an asynchronous tcgen05.ld are only defined after tcgen05.wait::ld; the compiler sees no
// dependency between the two asm statements and the register uses, so every register is "touched" by an empty volatile
// asm after the wait (no instruction is emitted) to pin the uses behind it.
template <int N>
__device__ __forceinline__ void reg_fence(uint32_t (&r)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
}

template <int D, bool kKT, typename OutT, bool kSeg>
__global__ void __launch_bounds__(kLazyThreads, 2)
sage_attn_lazy_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr uint32_t K_TILE = BN * D;
  constexpr uint32_t V_TILE = D * BN;
  constexpr int NS = (D == 128) ? 5 : 10;
  constexpr int SWQK = (D == 128) ? 128 : 64;
  constexpr uint32_t Q_BYTES = BM * D;
  constexpr int NG = kKT ? 4 : 1;
  constexpr float kTau = float(SAB_LAZY_TAU);
  constexpr float kOff = kFp8Offset - kTau;     // where a NEW running max is placed: 448 / 2^tau

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + NS * K_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * V_TILE);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;    // [2] step(t) retired: S(t+2) ready in buffer t&1 AND PV(t) accumulated into O
  uint64_t* p_full = bars + 3;    // [2] 128 arrivals: P(j) stored (and O rescaled when the max moved)
  uint64_t* kv_full = bars + 5;   // [NS]
  uint64_t* kv_empty = kv_full + NS;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(kv_empty + NS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---------------- work decode (uniform across the CTA)
  int qt = blockIdx.x;
  if (p.causal) qt = p.n_q_tiles - 1 - qt;  // heaviest tiles first
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

  // ---------------- one-time setup
  if (warp == 4 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(p_full + i, 128);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<kTmemCols>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  auto s_parity = [](int t) { return uint32_t(t >> 1) & 1u; };

  if (warp >= 4) {
    setmaxnreg_dec_48b();
    if (warp == 4) {
      // =============================== TMA producer ===============================
      if (lane == 0 && n_kv > 0) {
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_4d(sQ, &tmQ, q_full, 0, q_off + qt * BM, h, tb);
        [[maybe_unused]] int ready_seg = -1;
        for (int j = 0; j < n_kv; ++j) {
          const int s = j % NS;
          const uint32_t ph = (j / NS) & 1;
          int kc = k_off + j * BN, vc = v_off + j * BN, kb = tb;
          if (p.kv_seg_len > 0) {  // all-gathered layout: segment-major
            const int seg = (j * BN) / p.kv_seg_len;
            kc = vc = j * BN - seg * p.kv_seg_len;
            kb = seg * p.B + b;
            if constexpr (kSeg) {
              if (seg != ready_seg) {   // first tile of a segment: has the peer copy of this (head group, segment) landed?
                const uint32_t* flag = p.seg_flags + (hk / p.seg_heads) * (p.Sk / p.kv_seg_len) + seg;
                const long long t0 = clock64();
                while (ld_acquire_sys_u32(flag) != p.seg_epoch) {
                  __nanosleep(200);
                  if (clock64() - t0 > (8ll << 30)) __trap();   // ~4 s: the copies never came; fail the launch instead of hanging the GPU
                }
                fence_proxy_async_all();
                ready_seg = seg;
              }
            }
          }
          mbar_wait_wd(kv_empty + s, ph ^ 1);
#ifdef SAB_DBG_NO_TMA     // timing experiment only (results are garbage): the ring turns over without any K/V traffic
          mbar_arrive(kv_full + s);
#else
          mbar_expect_tx(kv_full + s, K_TILE + V_TILE);
          tma_load_4d(sK + s * K_TILE, &tmK, kv_full + s, 0, kc, hk, kb);
          tma_load_4d(sV + s * V_TILE, &tmV, kv_full + s, vc, 0, hk, kb);
#endif
        }
      }
    } else if (warp == 5) {
      // =============================== MMA issuer ===============================
      // step(t) = PV(t) ; QK(t+2) ; commit -> s_full[t&1].   In-order tensor pipe:  QK0 QK1 | PV0 QK2 | PV1 QK3 | ...
      if (n_kv > 0) {   // whole warp runs the loop (uniform control flow); one elected lane issues
        constexpr uint32_t idesc_qk = make_idesc(2, 1, 1, BM, BN);  // s32 <- s8 x s8, 128 x 64
        constexpr uint32_t idesc_pv = make_idesc(1, 0, 0, BM, D);   // f32 <- e4m3 x e4m3, 128 x D
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
#ifndef SAB_DBG_NO_MMA    // timing experiment only: no tensor-core work, the commits still flow
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 32; ++k) umma_i8_ss(tS, dQ + 2 * k, dK + 2 * k, idesc_qk, k > 0);
          }
#endif
        };
        mbar_wait_wd(q_full, 0);
        issue_qk(0, true);
        if (elect_one()) tc_commit(s_full + 0);
        if (n_kv > 1) {
          issue_qk(1, true);
          if (elect_one()) tc_commit(s_full + 1);
        }
#ifdef SAB_TIMELINE
        const bool tl_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
        long long* tl = reinterpret_cast<long long*>(p.dbg) + 4096;
#endif
        for (int j = 0; j < n_kv; ++j) {
          SAB_TL(8);
          if (j + 2 < n_kv) mbar_wait_wd(kv_full + (j + 2) % NS, ((j + 2) / NS) & 1);
          mbar_wait_wd(p_full + (j & 1), (j >> 1) & 1);   // P(j) stored, O rescaled if needed (also: S(j) fully consumed)
          tc_fence_after();
          SAB_TL(9);
          const int st = j % NS;
          const uint64_t dV = dV0 + uint64_t(st) * (V_TILE >> 4);
          const uint32_t tP = tmem_u + (j & 1) * BN;
#ifndef SAB_DBG_NO_MMA
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BN / 32; ++k) umma_f8_ts(tmem_u + 128, tP + 8 * k, dV + 2 * k, idesc_pv, (j > 0 || k > 0));
          }
#endif
          SAB_TL(10);
          if (j + 2 < n_kv) issue_qk(j + 2, false);
          if (elect_one()) {
            tc_commit(s_full + (j & 1));
            tc_commit(kv_empty + st);
          }
          SAB_TL(11);
        }
      }
    }
  } else {
    setmaxnreg_inc_208();
    // =============================== softmax / in-line correction / epilogue ===============================
    const int row = warp * 32 + lane;  // TMEM lane == Q row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + 128;
    const int q_row = qt * BM + row;

    int q_idx = (q_blk0 + qt) * p.q_mult;
    if (p.q_gran == SAB_GRAN_PER_WARP) q_idx += row >> 5;
    if (p.q_gran == SAB_GRAN_PER_THREAD) q_idx += (row >> 5) * 8 + (row & 7);
    const float* qs_base = p.q_scale + (varlen ? int64_t(h) : (int64_t(b) * p.Hq + h) * p.qs_stride_bh);
    const float* ks_base = p.k_scale + (varlen ? int64_t(hk) : (int64_t(b) * p.Hkv + hk) * p.ks_stride_bh);
    const float qss = qs_base[int64_t(q_idx) * p.qs_stride_idx] * p.sm_scale_log2;

    float m = kMaskValue;  // running (lazy) max, log2 units, includes the exponent offset
    float d = 0.f;         // running sum of fp32 P relative to m
#ifdef SAB_TIMELINE
    const bool tl_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    long long* tl = reinterpret_cast<long long*>(p.dbg) + 4096;
#endif
    // first tile that needs masking (ragged end of the keys / causal diagonal): all earlier tiles are fully visible to all 128 rows
    int j_mask0 = kv_len / BN;
    if (p.causal) j_mask0 = min(j_mask0, (p.causal_q_offset + qt * BM + 1) / BN);
    // K dequant scales of the NEXT tile are fetched one tile ahead (their L1/L2 latency used to sit at the top of every tile)
    const float* ks_ptr = ks_base + int64_t(k_blk0) * NG * p.ks_stride_idx;
    const int64_t ks_step = int64_t(NG) * p.ks_stride_idx;
    const bool ks_vec = kKT && p.ks_stride_idx == 1 && (reinterpret_cast<uintptr_t>(ks_ptr) & 15) == 0;
    float coef_cur[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) coef_cur[g] = (n_kv > 0 ? ks_ptr[int64_t(g) * p.ks_stride_idx] : 0.f) * qss;
    bool pending = false;   // the hand-off of the previous tile's P (wait::st + arrive on p_full) is still owed

    // One tile.  `s`: the 64 S values of tile j (already loaded and waited for); `nxt`: register set the S values of tile j+1
    // are prefetched into (tcgen05.ld issued inside the exponential loop; the caller waits before using them).
    auto do_tile = [&](int j, uint32_t (&s)[BN], uint32_t (&nxt)[BN]) {
      const uint32_t tS = tmem_base + lane_off + (j & 1) * BN;
      float coef[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) coef[g] = coef_cur[g];
      const bool has_next = j + 1 < n_kv;
      [[maybe_unused]] float ks_next[NG];
      auto load_ks = [&](int t, float (&dst)[NG]) {   // the NG dequant scales of key tile t (dense per-thread layout: one 16-byte load)
        if constexpr (kKT) {
          if (ks_vec) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(ks_ptr) + t);
            dst[0] = v4.x; dst[1] = v4.y; dst[2] = v4.z; dst[3] = v4.w;
            return;
          }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) dst[g] = __ldg(ks_ptr + t * ks_step + int64_t(g) * p.ks_stride_idx);
      };
      if constexpr (SAB_LZ_KSPRE != 0) {
        if (has_next) load_ks(j + 1, ks_next);
      } else {
        load_ks(j, coef);
#pragma unroll
        for (int g = 0; g < NG; ++g) coef[g] *= qss;
      }
      const bool masked_tile = j >= j_mask0;
      bool next_issued = false;
      bool nxt_ok = false;

      SAB_TL(0);
      auto hand_off = [&](int t) {   // P(t) is complete in TMEM: let the MMA warp issue PV(t)
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(p_full + (t & 1));
      };
      // S(j+1) ready == step(j-1) retired (the same mbarrier phase): once waited for, PV(j-1) is also accumulated in O
      auto prefetch_next = [&]() {
        if (has_next) {
          SAB_TL(5);
          if (!nxt_ok) mbar_wait_wd(s_full + ((j + 1) & 1), s_parity(j + 1));
          tc_fence_after();
          SAB_TL(6);
          const uint32_t tN = tmem_base + lane_off + ((j + 1) & 1) * BN;
          uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&nxt[0]);
          uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&nxt[32]);
#ifndef SAB_DBG_NO_LDTM   // timing experiment only: S is never read (stale registers)
          tmem_ld32(tN, lo);
          tmem_ld32(tN + 32, hi);
#endif
        }
        next_issued = true;
      };

      // exponentials of the tile against the reference max `mm`: P -> TMEM (e4m3), returns sum(P); MASKED zeroes i >= limit.
      // PRE (speculative pass): also gathers the integer row maxima (pm), takes the previous tile's deferred hand-off after the
      // first exponentials and prefetches S(j+1) three quarters into the loop.
      auto exp_row = [&](auto masked_tag, auto pre_tag, float mm, int limit, int (&pm)[4]) -> float {
        constexpr bool MASKED = decltype(masked_tag)::value;
        constexpr bool PRE = decltype(pre_tag)::value;
        uint64_t coef2[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) coef2[g] = pack_f2(coef[g], coef[g]);
        const uint64_t nm2 = pack_f2(-mm, -mm);
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
        uint32_t pk4[4];
#pragma unroll
        for (int w = 0; w < BN / 4; ++w) {
          if constexpr (PRE) {
            if constexpr (SAB_LZ_DEFER != 0) {
              if (w == 1 && pending) {     // the previous tile's P stores have long landed: its hand-off costs no wait here
                hand_off(j - 1);
                pending = false;
              }
            }
            if constexpr (SAB_LZ_TESTWAIT != 0) {   // result used at the prefetch point
              if (w == SAB_LZ_PREFETCH_AT - 3 && has_next) nxt_ok = mbar_test_wait(s_full + ((j + 1) & 1), s_parity(j + 1));
            }
            if (w == SAB_LZ_PREFETCH_AT) prefetch_next();
          }
          float e[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const int i = 4 * w + u;
            const int g = kKT ? ((i & 7) >> 1) : 0;
            const uint64_t f2 = pack_f2(__int2float_rn(int(s[i])), __int2float_rn(int(s[i + 1])));
            float y0, y1;
            unpack_f2(ffma2(f2, coef2[g], nm2), y0, y1);
#if SAB_POLY_EXP_PAIRS > 0
            if (((i >> 1) & 3) < SAB_POLY_EXP_PAIRS) {
              ex2_poly2(y0, y1, e[u], e[u + 1]);
            } else
#endif
            {
#ifdef SAB_DBG_NO_EXP     // timing experiment only: no MUFU
              e[u] = y0 * 0.001f;
              e[u + 1] = y1 * 0.001f;
#else
              e[u] = ex2_approx(y0);
              e[u + 1] = ex2_approx(y1);
#endif
            }
            if constexpr (MASKED) {
              e[u] = (i < limit) ? e[u] : 0.f;
              e[u + 1] = (i + 1 < limit) ? e[u + 1] : 0.f;
            }
            acc[(w & 1) * 2 + (u >> 1)] = fadd2(acc[(w & 1) * 2 + (u >> 1)], pack_f2(e[u], e[u + 1]));
            if constexpr (PRE) {   // four independent chains of the running integer maximum
              const int c = (i >> 1) & 3;
              pm[c] = __vimax3_s32(pm[c], int(s[i]), int(s[i + 1]));
            }
          }
          pk4[w & 3] = pack_e4m3x4(e[0], e[1], e[2], e[3]);
          // P goes out in 4-column pieces as it is produced (short live ranges: tcgen05.st needs consecutive registers).  The
          // speculative pass may do that too: S columns [0,16) — the ones P overwrites — stay in registers for a re-run.
#ifdef SAB_DBG_NO_STTM    // timing experiment only: P is never written
          if ((w & 3) == 3) asm volatile("" :: "r"(pk4[0]), "r"(pk4[1]), "r"(pk4[2]), "r"(pk4[3]));
#else
          if ((w & 3) == 3) tmem_st4(tS + (w >> 2) * 4, pk4[0], pk4[1], pk4[2], pk4[3]);
#endif
        }
        float a0, a1, a2, a3;
        unpack_f2(fadd2(acc[0], acc[1]), a0, a1);
        unpack_f2(fadd2(acc[2], acc[3]), a2, a3);
        return (a0 + a1) + (a2 + a3);
      };

      bool done = false;
      const bool speculate = j > 0 && !masked_tile;
      if (speculate) {
        // ---- speculative single pass against the current max.  "some P of this row would exceed 448" is decided on the integer
        //      maxima: S*coef - m > 8.807  <=>  S > (m + 8.807) / coef (coef > 0); thresholds are computed before the loop.
        [[maybe_unused]] int thr[NG];
        if constexpr (SAB_LZ_ITHR != 0) {
#pragma unroll
          for (int g = 0; g < NG; ++g) thr[g] = __float2int_rd(fminf(__fdividef(m + kFp8Offset, coef[g]), 1.0e9f));
        }
        int pm[4] = {kIntSentinel, kIntSentinel, kIntSentinel, kIntSentinel};
        const float sum = exp_row(std::false_type{}, std::true_type{}, m, 0, pm);
        SAB_TL(1);
        bool viol;
        if constexpr (SAB_LZ_ITHR != 0) {
          if constexpr (kKT) viol = (pm[0] > thr[0]) || (pm[1] > thr[1]) || (pm[2] > thr[2]) || (pm[3] > thr[3]);
          else viol = max(max(pm[0], pm[1]), max(pm[2], pm[3])) > thr[0];
        } else {
          // conservative for per-thread K scales: (largest S of the row) x (largest of the four group scales) — may send a tile to
          // the classic path a little early (the scales of one 64-key block differ by a small factor), never misses an overflow
          const int v = max(max(pm[0], pm[1]), max(pm[2], pm[3]));
          float cmax = coef[0];
#pragma unroll
          for (int g = 1; g < NG; ++g) cmax = fmaxf(cmax, coef[g]);
          viol = float(v) * cmax - m > kFp8Offset;
        }
        if (!__any_sync(0xffffffffu, viol)) {   // no P above 448 in this warp's 32 rows: commit
          d += sum;
          done = true;
        }
      }
      if (!done) {
        // ---- classic tile: row max -> lazy max update -> alpha -> exponentials; this thread rescales its own row of O
        if (pending) {
          hand_off(j - 1);
          pending = false;
        }
        int limit = kv_len - j * BN;
        if (p.causal) limit = min(limit, p.causal_q_offset + q_row - j * BN + 1);
        if (speculate) {   // failed speculation (rare): columns [0,16) of S(j) were kept in registers (P overwrote them in TMEM),
          uint32_t (&mid)[16] = *reinterpret_cast<uint32_t (*)[16]>(&s[16]);   // the rest is re-read from TMEM
          uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
          tmem_ld16(tS + 16, mid);
          tmem_ld32(tS + 32, hi);
          tc_wait_ld();
          reg_fence(s);
        }
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= limit) s[i] = uint32_t(kIntSentinel);
        }
        int pm[4] = {kIntSentinel, kIntSentinel, kIntSentinel, kIntSentinel};
#pragma unroll
        for (int i = 0; i < BN; i += 2) {
          const int c = kKT ? ((i & 7) >> 1) : ((i >> 1) & 3);
          pm[c] = __vimax3_s32(pm[c], int(s[i]), int(s[i + 1]));
        }
        float mx = kMaskValue;   // scaled row maximum from the integer maxima per scale group (scales are positive)
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int v = kKT ? pm[g] : max(max(pm[0], pm[1]), max(pm[2], pm[3]));
          mx = fmaxf(mx, (v == kIntSentinel) ? kMaskValue : float(v) * coef[g]);
        }
        const float m_true = fmaxf(m, mx - kOff);
        const float m_new = (m_true - m > kTau) ? m_true : m;   // == "some P would exceed 448" (mx - m > 8.807)
        const float alpha = ex2_approx(m - m_new);
        d *= alpha;
        m = m_new;
        if (!next_issued) prefetch_next();
        int unused[4];
        const float sum = masked_tile ? exp_row(std::true_type{}, std::false_type{}, m_new, limit, unused)
                                      : exp_row(std::false_type{}, std::false_type{}, m_new, limit, unused);
        d += sum;
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
          if (!has_next) {   // otherwise the prefetch above already waited for this phase
            mbar_wait_wd(s_full + ((j + 1) & 1), s_parity(j + 1));   // step(j-1) retired: PV(j-1) is in O
            tc_fence_after();
          }
          const uint64_t alpha2 = pack_f2(alpha, alpha);
#pragma unroll
          for (int ch = 0; ch < D / 32; ++ch) {
            uint32_t r[32];
            tmem_ld32(tO + ch * 32, r);
            tc_wait_ld();       // (also completes the prefetch of S(j+1))
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
      }
      SAB_TL(2);
      if constexpr (SAB_LZ_KSPRE != 0) {
#pragma unroll
        for (int g = 0; g < NG; ++g) coef_cur[g] = ks_next[g] * qss;
      }
      // the hand-off of a speculative tile is deferred into the next tile's loop (its wait::st + fence + arrive then cost nothing);
      // the MMA warp has a tile of slack before S(j+2) is needed
      if (SAB_LZ_DEFER != 0 && done && has_next) {
        pending = true;
      } else {
        hand_off(j);
      }
      SAB_TL(4);
    };

    if (n_kv > 0) {
      uint32_t sa[BN], sb[BN];
      mbar_wait_wd(s_full + 0, 0);
      tc_fence_after();
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sa[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sa[32]);
        tmem_ld32(tmem_base + lane_off, lo);
        tmem_ld32(tmem_base + lane_off + 32, hi);
      }
      for (int j = 0; j < n_kv; j += 2) {
        tc_wait_ld();
        reg_fence(sa);
        do_tile(j, sa, sb);
        if (j + 1 < n_kv) {
          tc_wait_ld();
          reg_fence(sb);
          do_tile(j + 1, sb, sa);
        }
      }
    }

    // ---- epilogue: O / d * v_scale (+ v_mean) -> fp16/bf16 (qk_int_sv_f8_cuda_sm89.cuh:572-703)
    OutT* orow = reinterpret_cast<OutT*>(p.out) + (varlen ? 0 : int64_t(b) * p.o_stride_b) + int64_t(h) * p.o_stride_h +
                 int64_t(q_off + q_row) * p.o_stride_s;
    const bool row_ok = q_row < q_len;
    const float* vs = p.v_scale ? p.v_scale + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D : nullptr;
    const float* vm = p.v_mean ? p.v_mean + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D : nullptr;
    if (n_kv > 0) {
      mbar_wait_wd(s_full + ((n_kv + 1) & 1), s_parity(n_kv + 1));   // step(n_kv-1) retired: O is final
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
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) dst[v4] = make_uint4(o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
        }
      }
    } else if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
      for (int v4 = 0; v4 < D / 8; ++v4) dst[v4] = make_uint4(0, 0, 0, 0);
    }
    if (p.lse != nullptr && row_ok) {
      const int64_t li = varlen ? (int64_t(h) * p.Sq + q_off + q_row) : ((int64_t(b) * p.Hq + h) * p.Sq + q_row);
      p.lse[li] = n_kv > 0 ? lg2_approx(d) + m : -INFINITY;   // log2 units; the offset cancels: log2(sum 2^x)
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<kTmemCols>(tmem_base);
}

template <int D, bool kKT, typename OutT, bool kSeg>
int launch_attn_lazy(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                     cudaStream_t stream) {
  constexpr int NS = (D == 128) ? 5 : 10;
  size_t smem = size_t(BM) * D + size_t(NS) * 2 * BN * D + 512;
  if (smem < 80 * 1024) smem = 80 * 1024;   // keep it at two CTAs per SM (TMEM: 2 x 256 columns)
  auto kern = sage_attn_lazy_kernel<D, kKT, OutT, kSeg>;
  static bool configured[64] = {};
  if (int st = ensure_dynamic_smem(kern, smem, configured)) return st;
  kern<<<grid, kLazyThreads, smem, stream>>>(tq, tk, tv, p);
  SAB_CUDA_OK(cudaGetLastError());
  return SAB_OK;
}

#define SAB_INST(D, KT, T, SEG) \
  template int launch_attn_lazy<D, KT, T, SEG>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const AttnParams&, dim3, cudaStream_t);
SAB_INST(128, true, __nv_bfloat16, false)
SAB_INST(128, true, __half, false)
SAB_INST(128, false, __nv_bfloat16, false)
SAB_INST(128, false, __half, false)
SAB_INST(128, true, __nv_bfloat16, true)
SAB_INST(128, true, __half, true)
SAB_INST(128, false, __nv_bfloat16, true)
SAB_INST(128, false, __half, true)
#undef SAB_INST

}  // namespace sab
