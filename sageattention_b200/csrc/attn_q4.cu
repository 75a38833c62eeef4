// Fused INT8-QK / FP8-PV attention for sm_100a (B200), head_dim 128: ONE CTA per SM, FOUR softmax warpgroups.
//
// attn_alt.cu (two CTAs per SM, two softmax warpgroups each) leaves the MUFU 31 % idle: with 256 TMEM columns per CTA the e4m3
// P(j) has to alias its own S buffer, so QK(j+2) cannot be issued before PV(j) has consumed P(j), and every warpgroup waits
// ~350+ cycles per tile for its next S (21 % of the softmax warps' time).  One CTA owns all 512 TMEM columns here:
//   S   [0,256)    four 64-column buffers, tile g in buffer g & 3, softmax by warpgroup g & 3
//   P   [256,320)  four 16-column e4m3 buffers of their own
//   O   [384,512)  fp32 accumulator
// so QK(g+4) only needs the owner of tile g to have READ S(g) (`s_free`, signalled in the middle of the exponentials), not
// PV(g): the next S of a warpgroup is ready long before it is needed.  Two issuing warps feed the tensor pipe independently —
// one for QK^T (waits s_free / K tiles, commits s_full), one for PV (waits p_full / V tiles, commits pv_done) — and K and V
// travel in separate rings with their own producers, because K(g+4) is consumed about four tiles before V(g).
//
// Measured (profiles/r02_q4_first_contact.log): +2-3 % over attn_alt.cu at S = 16K-32K, -2 % at S = 8K — the S wait this design
// removes is not where the idle MUFU time goes.  OPT-IN (SAB_ATTN_KERNEL=q4); the product kernel is attn_alt.cu.
//
// A CTA takes SAB_Q4_ITEMS consecutive work items (item = Q tile of one (batch, head); default 1).  The code is written for a stream
// of items: the key tiles of all items of a CTA form ONE sequence g = 0, 1, 2, ...; buffers, warpgroup assignment and mbarrier
// parities run on g across item boundaries, the QK^T issuer runs ahead into the next item (Q is double-buffered), and only O is
// single — the first PV of an item waits until the epilogue of the previous item has loaded O from TMEM (`o_free`).  That form is
// correct (also as a persistent grid of #SMs CTAs walking the whole list) but slow: with more than one copy of the tile body ptxas
// spills 64-88 bytes per thread inside the exponential loop, and 150-200 KB of shared memory leave almost no L1 for local memory.
// Everything on the softmax side is attn_alt.cu's: one thread per row and tile, lazy running max (SAB_ALT_TAU), the running max
// chained through shared memory (`m_full`), in-line O rescale (rare), partial row sums per warpgroup combined in the epilogue,
// TMA-store epilogue (own staging buffer: with several items per CTA the K ring is busy with the next item).
// 640 threads: warps 0-15 softmax (104 registers), 16 Q + K producer, 17 QK issuer + TMEM allocator, 18 PV issuer, 19 V producer
// (64 registers): 512 x 104 + 128 x 64 = 640 x 96, the CTA's register pool.
#include <cstdlib>
#include "attn_common.cuh"

namespace sab {

constexpr int kQ4Threads = 640;
#ifndef SAB_Q4_ITEMS
#define SAB_Q4_ITEMS 1
#endif
constexpr int kQ4Items = SAB_Q4_ITEMS;   // consecutive work items (Q tiles of one head) per CTA
constexpr uint32_t kQ4TmemCols = 512;
#ifndef SAB_ALT_TAU
#define SAB_ALT_TAU 4
#endif

__device__ __forceinline__ void setmaxnreg_inc_104() { asm volatile("setmaxnreg.inc.sync.aligned.u32 104;"); }
__device__ __forceinline__ void setmaxnreg_dec_64() { asm volatile("setmaxnreg.dec.sync.aligned.u32 64;"); }
__device__ __forceinline__ void q4_bar_sync_all() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the four softmax warpgroups

// One work item: the Q tile `qt` of (batch b, head h).  Every role decodes the same list in the same order.
struct Q4Item {
  int qt, h, b, hk, q_len, kv_len, q_off, k_off, v_off, tb, q_blk0, k_blk0, n_kv;
  bool valid;
};
__device__ __forceinline__ Q4Item q4_decode(const AttnParams& p, int it) {
  Q4Item I;
  const int qr = it % p.n_q_tiles;            // Q tiles of one (b, h) are adjacent in the list: their K/V stay in L2
  const int hb = it / p.n_q_tiles;
  I.qt = p.causal ? p.n_q_tiles - 1 - qr : qr;   // causal: heaviest first
  I.h = hb % p.Hq;
  I.b = hb / p.Hq;
  I.hk = I.h / (p.Hq / p.Hkv);
  I.q_len = p.Sq; I.kv_len = p.Sk; I.q_off = 0; I.k_off = 0; I.v_off = 0; I.tb = I.b; I.q_blk0 = 0; I.k_blk0 = 0;
  I.valid = true;
  if (p.cu_q != nullptr) {
    I.q_off = p.cu_q[I.b];
    I.q_len = p.cu_q[I.b + 1] - I.q_off;
    I.k_off = p.cu_k[I.b];
    I.kv_len = p.cu_k[I.b + 1] - I.k_off;
    I.v_off = p.cu_v[I.b];
    I.q_blk0 = p.cu_qs[I.b];
    I.k_blk0 = p.cu_ks[I.b];
    I.tb = 0;
    if (I.qt * BM >= I.q_len) I.valid = false;
  }
  I.n_kv = (I.kv_len + BN - 1) / BN;
  if (p.causal) I.n_kv = min(I.n_kv, (p.causal_q_offset + (I.qt + 1) * BM + BN - 1) / BN);
  return I;
}

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
  constexpr uint32_t STAGE_BYTES = 2 * BM * 128;   // two [128 rows][128 B] halves of an output tile

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                            // [2] Q tiles (item parity)
  uint8_t* sK = sQ + 2 * Q_BYTES;
  uint8_t* sV = sK + NK * K_TILE;
  uint8_t* sStage = sV + NV * V_TILE;                            // epilogue staging for the TMA store
  float* s_m = reinterpret_cast<float*>(sStage + STAGE_BYTES);   // [4 buffers][128 rows] running max m(g)
  float* s_x = s_m + NW * BM;                                    // [2 item parity][4 warpgroups][2][128] epilogue exchange: d, last max
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + 2 * 2 * NW * BM);
  uint64_t* q_full = bars + 0;            // [2]
  uint64_t* q_empty = bars + 2;           // [2] every QK^T of the item that used this Q buffer has retired
  uint64_t* s_full = bars + 4;            // [4] QK(g) retired: S(g) in buffer g & 3
  uint64_t* s_free = s_full + NW;         // [4] 128 arrivals: the owner of tile g has read S(g) for the last time
  uint64_t* p_full = s_free + NW;         // [4] 128 arrivals: P(g) stored (and O rescaled when the max moved)
  uint64_t* pv_done = p_full + NW;        // [4] PV(g) retired: P buffer g & 3 free, O holds the item's tiles <= g
  uint64_t* m_full = pv_done + NW;        // [4] 128 arrivals: m(g) published in s_m[g & 3]
  uint64_t* o_free = m_full + NW;         // [1] 512 arrivals: the epilogue of an item has loaded O from TMEM
  uint64_t* k_full = o_free + 1;
  uint64_t* k_empty = k_full + NK;
  uint64_t* v_full = k_empty + NK;
  uint64_t* v_empty = v_full + NV;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(v_empty + NV);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool varlen = p.cu_q != nullptr;

  if (warp == 16 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full + i, 1);
      mbar_init(q_empty + i, 1);
    }
    for (int i = 0; i < NW; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(s_free + i, 128);
      mbar_init(p_full + i, 128);
      mbar_init(pv_done + i, 1);
      mbar_init(m_full + i, 128);
    }
    mbar_init(o_free, 512);
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

  // Stream counters, advanced identically by every role: g0 = key tiles of all earlier items of this CTA, vi = earlier valid items,
  // qc = earlier items with at least one key tile (they use a Q buffer).
  uint32_t g0 = 0;
  int vi = 0, qc = 0;

  if (warp >= 16) {
    setmaxnreg_dec_64();
    if (warp == 16 || warp == 19) {
      // =============================== TMA producers: warp 16 Q + K ring, warp 19 V ring ===============================
      const bool is_k = warp == 16;
      if (lane == 0) {
        for (int it = int(blockIdx.x) * kQ4Items; it < min(p.n_items, (int(blockIdx.x) + 1) * kQ4Items); ++it) {
          const Q4Item I = q4_decode(p, it);
          if (!I.valid) continue;
          if (I.n_kv > 0 && is_k) {
            const int qb = qc & 1;
            mbar_wait_wd(q_empty + qb, (uint32_t(qc >> 1) & 1u) ^ 1u);
            mbar_expect_tx(q_full + qb, Q_BYTES);
            tma_load_4d(sQ + qb * Q_BYTES, &tmQ, q_full + qb, 0, I.q_off + I.qt * BM, I.h, I.tb);
          }
          int ready_seg = -1;
          for (int j = 0; j < I.n_kv; ++j) {
            int kc = I.k_off + j * BN, vc = I.v_off + j * BN, kb = I.tb;
            if (p.kv_seg_len > 0) {
              const int seg = (j * BN) / p.kv_seg_len;
              kc = vc = j * BN - seg * p.kv_seg_len;
              kb = seg * p.B + I.b;
              if (p.seg_flags != nullptr && seg != ready_seg) {
                // gather fused into the launch: first tile of a segment -> has the peer copy of this (head group, segment) landed?
                const uint32_t* flag = p.seg_flags + (I.hk / p.seg_heads) * (p.Sk / p.kv_seg_len) + seg;
                const long long t0 = clock64();
                while (ld_acquire_sys_u32(flag) != p.seg_epoch) {
                  __nanosleep(200);
                  if (clock64() - t0 > (8ll << 30)) __trap();   // ~4 s: the copies never came; fail the launch instead of hanging
                }
                fence_proxy_async_all();
                ready_seg = seg;
              }
            }
            const uint32_t g = g0 + j;
            if (is_k) {
              const int s = g % NK;
              mbar_wait_wd(k_empty + s, ((g / NK) & 1u) ^ 1u);
              mbar_expect_tx(k_full + s, K_TILE);
              tma_load_4d(sK + s * K_TILE, &tmK, k_full + s, 0, kc, I.hk, kb);
            } else {
              const int s = g % NV;
              mbar_wait_wd(v_empty + s, ((g / NV) & 1u) ^ 1u);
              mbar_expect_tx(v_full + s, V_TILE);
              tma_load_4d(sV + s * V_TILE, &tmV, v_full + s, vc, 0, I.hk, kb);
            }
          }
          g0 += I.n_kv; ++vi; qc += I.n_kv > 0;
        }
      }
    } else if (warp == 17) {
      // =============================== QK^T issuer ===============================
      constexpr uint32_t idesc_qk = make_idesc(2, 1, 1, BM, BN);   // s32 <- s8 x s8, 128 x 64
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t dQ0 = make_smem_desc<128>(smem_u32(sQ));
      const uint64_t dK0 = make_smem_desc<128>(smem_u32(sK));
      for (int it = int(blockIdx.x) * kQ4Items; it < min(p.n_items, (int(blockIdx.x) + 1) * kQ4Items); ++it) {   // whole warp runs the loops (uniform control flow); one elected lane issues
        const Q4Item I = q4_decode(p, it);
        if (!I.valid) continue;
        if (I.n_kv > 0) {
          const int qb = qc & 1;
          const uint64_t dQ = dQ0 + uint64_t(qb) * (Q_BYTES >> 4);
          mbar_wait_wd(q_full + qb, uint32_t(qc >> 1) & 1u);
          for (int j = 0; j < I.n_kv; ++j) {
            const uint32_t g = g0 + j;
            const int st = g % NK;
            // buffer g & 3 held S(g-4): its owner arrives on s_free once its last tcgen05.ld of that tile has completed
            if (g >= NW) mbar_wait_wd(s_free + (g & 3), ((g >> 2) - 1) & 1u);
            mbar_wait_wd(k_full + st, (g / NK) & 1u);
            tc_fence_after();
            const uint64_t dK = dK0 + uint64_t(st) * (K_TILE >> 4);
            const uint32_t tS = tmem_u + (g & 3) * BN;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < D / 32; ++k) umma_i8_ss(tS, dQ + 2 * k, dK + 2 * k, idesc_qk, k > 0);
              tc_commit(s_full + (g & 3));
              tc_commit(k_empty + st);
              if (j == I.n_kv - 1) tc_commit(q_empty + qb);   // the Q buffer may be reloaded (for the item after next)
            }
            __syncwarp();
          }
        }
        g0 += I.n_kv; ++vi; qc += I.n_kv > 0;
      }
    } else {
      // =============================== PV issuer (warp 18) ===============================
      constexpr uint32_t idesc_pv = make_idesc(1, 0, 0, BM, D);    // f32 <- e4m3 x e4m3, 128 x 128
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t dV0 = make_smem_desc<64>(smem_u32(sV));
      for (int it = int(blockIdx.x) * kQ4Items; it < min(p.n_items, (int(blockIdx.x) + 1) * kQ4Items); ++it) {
        const Q4Item I = q4_decode(p, it);
        if (!I.valid) continue;
        // O is single-buffered: the epilogue of the previous item that used it (n_kv > 0) must have loaded it.  Items without key
        // tiles take no part in this hand-shake — their epilogue does not depend on this warp, so counting them would let the
        // softmax warpgroups run two o_free completions ahead of the parity tested here (tests/test_q4_protocol_model.py).
        if (I.n_kv > 0 && qc > 0) mbar_wait_wd(o_free, uint32_t(qc - 1) & 1u);
        for (int j = 0; j < I.n_kv; ++j) {
          const uint32_t g = g0 + j;
          const int st = g % NV;
          mbar_wait_wd(v_full + st, (g / NV) & 1u);
          mbar_wait_wd(p_full + (g & 3), (g >> 2) & 1u);
          tc_fence_after();
          const uint64_t dV = dV0 + uint64_t(st) * (V_TILE >> 4);
          const uint32_t tP = tmem_u + 256 + (g & 3) * 16;   // P(g): 16 columns (32 keys per 8 columns)
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BN / 32; ++k) umma_f8_ts(tmem_u + 384, tP + 8 * k, dV + 2 * k, idesc_pv, (j > 0 || k > 0));
            tc_commit(pv_done + (g & 3));
            tc_commit(v_empty + st);
          }
          __syncwarp();
        }
        g0 += I.n_kv; ++vi; qc += I.n_kv > 0;
      }
    }
  } else {
    // =============================== softmax of the stream tiles g = wg (mod 4) / in-line correction / epilogue ===============================
    setmaxnreg_inc_104();
    const int wg = warp >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + 384;

    // Parity waits (completion c of a per-buffer barrier belongs to stream tile 4c + buffer).  Sound only when the barrier is
    // neither a full cycle ahead of nor behind the waiter:
    //   * s_full[g&3], pv_done[g&3] by the owner of tile g: it owns every tile of that buffer and follows it completion by completion;
    //   * m_full[(g-1)&3] and pv_done[(g-1)&3] by the owner of tile g (same item): tile g+3 of that buffer needs m(g) / P(g) from
    //     this warpgroup or, in the next item, this warpgroup's arrival at the item's epilogue barrier (not ahead); tile g-5 was
    //     completed before this warpgroup's tile g-4 needed it, or before the previous item's epilogue barrier (not behind).
    // FULLY UNROLLED: with a back edge around the tile body ptxas spilled 64-88 bytes per thread inside the exponential loop (and 203 KB of
    // shared memory leave almost no L1 for local memory): 966 instead of 1475 TFLOP/s.  A single pass over the body compiles clean,
    // so the CTA takes a FIXED, small number of items and the item loop disappears at compile time.
#pragma unroll
    for (int u = 0; u < kQ4Items; ++u) {
      const int it = int(blockIdx.x) * kQ4Items + u;
      if (it >= p.n_items) break;
      // Each copy of the item body derives its addresses and thread geometry from opaque values of its own: shared between the
      // copies (common sub-expressions) they stayed live across the tile bodies and were spilled inside the exponential loop.
      uint8_t* sm = smem;
      uint32_t c_tid = threadIdx.x, c_tmem_base = tmem_base;
      asm volatile("" : "+l"(sm), "+r"(c_tid), "+r"(c_tmem_base));
#define Q4_LOCAL(T, name) T const c_##name = reinterpret_cast<T>(sm + (reinterpret_cast<uint8_t*>(name) - smem))
      Q4_LOCAL(uint64_t*, s_full); Q4_LOCAL(uint64_t*, s_free); Q4_LOCAL(uint64_t*, p_full); Q4_LOCAL(uint64_t*, pv_done);
      Q4_LOCAL(uint64_t*, m_full); Q4_LOCAL(uint64_t*, o_free); Q4_LOCAL(float*, s_m); Q4_LOCAL(float*, s_x); Q4_LOCAL(uint8_t*, sStage);
#undef Q4_LOCAL
      const int c_wg = int(c_tid >> 7);
      const int c_row = int(c_tid & 127u);
      const uint32_t c_lane_off = (c_tid & 96u) << 16;
      const uint32_t c_tO = c_tmem_base + c_lane_off + 384;
      // Only what the tile loop needs stays live across it (the warpgroup runs at 104 registers); the epilogue decodes the item again.
      int n_kv, kv_len, cq;       // cq: causal bound of this thread's c_row (key index < cq + c_row is visible); "infinite" when not causal
      const float* ks_ptr;
      float qss;
      {
        const Q4Item I = q4_decode(p, it);
        if (!I.valid) continue;
        n_kv = I.n_kv;
        kv_len = I.kv_len;
        cq = p.causal ? p.causal_q_offset + I.qt * BM + 1 : (1 << 30);
        int q_idx = (I.q_blk0 + I.qt) * p.q_mult;
        if (p.q_gran == SAB_GRAN_PER_WARP) q_idx += c_row >> 5;
        if (p.q_gran == SAB_GRAN_PER_THREAD) q_idx += (c_row >> 5) * 8 + (c_row & 7);
        const float* qs_base = p.q_scale + (varlen ? int64_t(I.h) : (int64_t(I.b) * p.Hq + I.h) * p.qs_stride_bh);
        const float* ks_base = p.k_scale + (varlen ? int64_t(I.hk) : (int64_t(I.b) * p.Hkv + I.hk) * p.ks_stride_bh);
        ks_ptr = ks_base + int64_t(I.k_blk0) * NG * p.ks_stride_idx;
        qss = qs_base[int64_t(q_idx) * p.qs_stride_idx] * p.sm_scale_log2;
      }

      float m_own = kMaskValue;   // m of this warpgroup's latest tile of the item: the reference its partial sum d is relative to
      float d = 0.f;              // sum of P over THIS warpgroup's tiles of the item, relative to m_own

      for (int j = int((uint32_t(c_wg) - g0) & 3u); j < n_kv; j += NW) {
        const uint32_t g = g0 + j;
        const int bf = c_wg;          // == g & 3
        const uint32_t tS = c_tmem_base + c_lane_off + bf * BN;
        const uint32_t tP = c_tmem_base + c_lane_off + 256 + bf * 16;
        float coef[NG];
        if constexpr (kKT) {
          if (p.ks_vec4) {   // dense: the four per-thread scales of a key tile are one aligned 16-byte word
            const float4 k4 = *reinterpret_cast<const float4*>(ks_ptr + int64_t(j) * 4);
            coef[0] = k4.x * qss; coef[1] = k4.y * qss; coef[2] = k4.z * qss; coef[NG - 1] = k4.w * qss;
          } else {
#pragma unroll
            for (int gq = 0; gq < NG; ++gq) coef[gq] = ks_ptr[int64_t(j * NG + gq) * p.ks_stride_idx] * qss;
          }
        } else {
          coef[0] = ks_ptr[int64_t(j) * p.ks_stride_idx] * qss;
        }
        const int limit = min(kv_len, cq + c_row) - j * BN;
        const bool masked_tile = (kv_len - j * BN < BN) || ((j + 1) * BN > cq);

        mbar_wait_wd(c_s_full + bf, (g >> 2) & 1u);
        tc_fence_after();

        auto tile = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
          // ---- pass 1: c_row max, streamed in two 32-column loads
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
              for (int gq = 0; gq < 4; ++gq) pm[gq] = __vimax3_s32(pm[gq], int(s[i8 + 2 * gq]), int(s[i8 + 2 * gq + 1]));
            }
          }
          float mx = kMaskValue;
#pragma unroll
          for (int gq = 0; gq < NG; ++gq) {
            const int v = kKT ? pm[gq] : max(max(pm[0], pm[1]), max(pm[2], pm[3]));
            float c = float(v) * coef[gq];
            if constexpr (MASKED) c = (v == kIntSentinel) ? kMaskValue : c;
            mx = fmaxf(mx, c);
          }
          // ---- the running max: m(g-1) comes from warpgroup (g-1) & 3, which published it right after ITS c_row max.  Slot reuse:
          //      I overwrite slot g & 3 (m(g-4)), last read by the owner of tile g-3 before it published m(g-3); m(g-1), which I
          //      wait for here, came after m(g-2), after m(g-3) (or the previous item's epilogue barrier lies in between).
          float m_prev = kMaskValue;
          if (j > 0) {
            mbar_wait_wd(c_m_full + ((g - 1) & 3), ((g - 1) >> 2) & 1u);
            m_prev = c_s_m[((g - 1) & 3) * BM + c_row];
          }
#if SAB_ALT_TAU > 0
          const float m_true = fmaxf(m_prev, mx - (kFp8Offset - float(SAB_ALT_TAU)));
          const float m_new = (m_true - m_prev > float(SAB_ALT_TAU)) ? m_true : m_prev;
#else
          const float m_new = fmaxf(m_prev, mx - kFp8Offset);   // update_mdo, attn_utils.cuh:377-396
#endif
          c_s_m[bf * BM + c_row] = m_new;
          mbar_arrive(c_m_full + bf);
          float alpha_o = 1.0f;                                   // rescale of O before PV(g): consecutive tiles
          if (__any_sync(0xffffffffu, (m_new != m_prev) | (m_new != m_own))) {   // lazy max: rare after the first tiles (ex2(0) = 1)
            alpha_o = ex2_approx(m_prev - m_new);
            d *= ex2_approx(m_own - m_new);                        // my partial sum: relative to my previous tile (g-4)
          }
          m_own = m_new;

          uint64_t coef2[NG];
#pragma unroll
          for (int gq = 0; gq < NG; ++gq) coef2[gq] = pack_f2(coef[gq], coef[gq]);
          const uint64_t nm2 = pack_f2(-m_new, -m_new);
          uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
          // ---- pass 2: exponentials, again 32 columns at a time; P goes to its own buffer (free once PV(g-4) has retired)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t s[32];
            tmem_ld32(tS + 32 * hf, s);
            tc_wait_ld();
            if (hf == 1) {             // last read of S(g): the QK issuer may overwrite the buffer with S(g+4)
              tc_fence_before();
              mbar_arrive(c_s_free + bf);
            }
            uint32_t pk[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              float e[4];
#pragma unroll
              for (int u = 0; u < 4; u += 2) {
                const int i = 4 * w + u;
                const int gq = kKT ? ((i & 7) >> 1) : 0;
                const uint64_t f2 = pack_f2(__int2float_rn(int(s[i])), __int2float_rn(int(s[i + 1])));
                float y0, y1;
                unpack_f2(ffma2(f2, coef2[gq], nm2), y0, y1);
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
            if (hf == 0 && g >= NW) {   // P buffer g & 3 still holds P(g-4) until PV(g-4) has retired
              mbar_wait_wd(c_pv_done + bf, ((g >> 2) - 1) & 1u);
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

          // ---- in-line correction of this c_row of O, when any c_row of the warp moved its max: needs PV(g-1) accumulated
          if (j > 0 && __any_sync(0xffffffffu, alpha_o != 1.0f)) {
            mbar_wait_wd(c_pv_done + ((g - 1) & 3), ((g - 1) >> 2) & 1u);
            tc_fence_after();
            const uint64_t alpha2 = pack_f2(alpha_o, alpha_o);
#pragma unroll
            for (int ch = 0; ch < D / 32; ++ch) {
              uint32_t r[32];
              tmem_ld32(c_tO + ch * 32, r);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float lo, hi;
                unpack_f2(fmul2(pack_f2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), lo, hi);
                r[i] = __float_as_uint(lo);
                r[i + 1] = __float_as_uint(hi);
              }
              tmem_st32(c_tO + ch * 32, r);
            }
          }
        };
        if (masked_tile) tile(std::true_type{});
        else tile(std::false_type{});

        tc_wait_st();
        tc_fence_before();
        mbar_arrive(c_p_full + bf);
      }

      // ---- epilogue of the item: combine the four partial sums relative to the final max, then each warpgroup writes OC columns
      // (opaque copies: the item is decoded again and the epilogue's address arithmetic is kept from being hoisted out of the item
      //  loop — either would stay live across the tile loop, which runs at 104 registers)
      int it_e = it, row_e = c_row, wg_e = c_wg;
      asm volatile("" : "+r"(it_e), "+r"(row_e), "+r"(wg_e));
      const Q4Item I = q4_decode(p, it_e);
      const int qt = I.qt, h = I.h, b = I.b, hk = I.hk;
      const int q_row = qt * BM + row_e;
      float* sx = c_s_x + (vi & 1) * (2 * NW * BM);   // double-buffered: a fast warpgroup may already deposit the next item's sums
      sx[(wg_e * 2 + 0) * BM + row_e] = d;
      sx[(wg_e * 2 + 1) * BM + row_e] = m_own;
      if (n_kv > 0 && uint32_t(wg_e) == ((g0 + n_kv - 1) & 3u)) {   // owner of the last tile: O is final once its PV has retired
        mbar_wait_wd(c_pv_done + wg_e, ((g0 + n_kv - 1) >> 2) & 1u);
        tc_fence_after();
      }
      tc_fence_before();
      q4_bar_sync_all();
      tc_fence_after();
      float m_fin = kMaskValue;
#pragma unroll
      for (int w = 0; w < NW; ++w) m_fin = fmaxf(m_fin, sx[(w * 2 + 1) * BM + row_e]);
      d = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) d += sx[(w * 2 + 0) * BM + row_e] * ex2_approx(sx[(w * 2 + 1) * BM + row_e] - m_fin);

      uint32_t r[32];
      if (n_kv > 0) {
        tmem_ld32(c_tmem_base + (static_cast<uint32_t>(row_e & ~31) << 16) + 384 + wg_e * OC, r);
        tc_wait_ld();
      }
      tc_fence_before();
      if (n_kv > 0) mbar_arrive(c_o_free);   // the PV issuer may start the next item (O is overwritten by its first PV)

      const bool row_ok = q_row < I.q_len;
      // Dense outputs leave through TMA: a PAIR of warpgroups stages its [128 rows][128 B] half of the tile (128-byte swizzle;
      // warpgroup 2i + 1 fills the upper four 16-byte chunks of a row_e) and one thread issues the bulk store.
      const bool use_tma = kTmaStoreEpilogue && p.o_tma != 0;
      static_assert(OC * sizeof(OutT) == 64, "staging: 128 rows x 128 bytes per warpgroup pair");
      const int pair = wg_e >> 1;
      uint8_t* stage = c_sStage + pair * (BM * 128);
      const uint32_t stage_row = smem_u32(stage) + row_e * 128;
      const float* vs = p.v_scale ? p.v_scale + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + wg_e * OC : nullptr;
      const float* vm = p.v_mean ? p.v_mean + (int64_t(varlen ? 0 : b) * p.Hkv + hk) * D + wg_e * OC : nullptr;
      uint32_t o16[16];
      if (n_kv > 0) {
        const float inv = rcp_approx(d);
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
        // (the previous item's store has finished reading the staging buffer: its issuing thread waited for that before it
        //  arrived at this item's barrier above)
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4)
          st_shared_v4(stage_row + ((((wg_e & 1) * 4 + v4) ^ (row_e & 7)) << 4), o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
        fence_proxy_async_smem();                                        // generic-proxy stores -> visible to the TMA engine
        if (pair == 0) asm volatile("bar.sync 2, 256;" ::: "memory");    // this pair's half tile is staged
        else asm volatile("bar.sync 3, 256;" ::: "memory");
        if ((threadIdx.x & 255) == 0) {
          tma_store_4d(&p.o_map, stage, pair * 128, qt * BM, h, b);
          tma_store_commit();
          tma_store_wait_read();                                         // the staging buffer must outlive the read
        }
      } else if (row_ok) {
        OutT* orow = reinterpret_cast<OutT*>(p.out) + (varlen ? 0 : int64_t(b) * p.o_stride_b) + int64_t(h) * p.o_stride_h +
                     int64_t(I.q_off + q_row) * p.o_stride_s + wg_e * OC;
        uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) dst[v4] = make_uint4(o16[4 * v4], o16[4 * v4 + 1], o16[4 * v4 + 2], o16[4 * v4 + 3]);
      }
      if (p.lse != nullptr && row_ok && wg_e == 0) {
        const int64_t li = varlen ? (int64_t(h) * p.Sq + I.q_off + q_row) : ((int64_t(b) * p.Hq + h) * p.Sq + q_row);
        p.lse[li] = n_kv > 0 ? lg2_approx(d) + m_fin : -INFINITY;
      }
      g0 += n_kv; ++vi;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<kQ4TmemCols>(tmem_base);
}

template <bool kKT, typename OutT>
int launch_attn_q4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p_in, dim3 grid,
                   cudaStream_t stream) {
  // 2 Q + 8 K + 8 V tiles + output staging + running-max / epilogue exchange + barriers: 203 KB, one CTA per SM (it owns all
  // 512 TMEM columns); grid = min(#items, #SMs) persistent CTAs
  const size_t smem = size_t(2) * BM * 128 + size_t(16) * BN * 128 + size_t(2) * BM * 128 + 20 * BM * sizeof(float) + 1024;
  auto kern = sage_attn_q4_kernel<kKT, OutT>;
  static bool configured[64] = {};
  if (int st = ensure_dynamic_smem(kern, smem, configured)) return st;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    SAB_CUDA_OK(cudaGetDevice(&dev));
    SAB_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  AttnParams p = p_in;
  p.n_items = int(grid.x * grid.y * grid.z);
  const int ctas = (p.n_items + kQ4Items - 1) / kQ4Items;
  kern<<<ctas, kQ4Threads, smem, stream>>>(tq, tk, tv, p);
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
