// tcgen05 flash-attention forward on CTA PAIRS (head_dim 128, non-causal, no mask):  O = softmax(scale · Q Kᵀ) V
// Replaces F.scaled_dot_product_attention at actionmesh/model/utils/attention_processor.py:133-139.
//
// A cluster of two CTAs (one TPC) owns 256 query rows of one (batch, head); each CTA holds 128 of them.  Every MMA is a
// `cta_group::2` instruction with M = 256 issued by the leader CTA:
//     S_j  = Q · K_jᵀ      SS form, 256 x 128 x 16: A = each CTA's own Q rows, B = K tile split 64 keys per CTA
//     O   += P_j · V_j     TS form, 256 x 128 x 16: A = bf16 P in each CTA's TMEM, B = V tile split 64 d-columns per CTA
// so each CTA stages only HALF of every K and V tile (half the TMA / shared-memory operand traffic of a single-CTA kernel
// for the same tensor work), and the accumulators of a CTA's 128 rows live in its own TMEM.
//
// TMEM (per CTA, 512 columns):  S/P buffers 0,1,2 at columns 0/128/256 (P_j = bf16 pairs over the first 64 columns of
// its own S_j), O at 384.  With three S buffers the QKᵀ of tile j+2 is issued BEFORE the P·V of tile j, so the tensor pipe
// always has independent work queued while the softmax warps turn S_j into P_j: the QKᵀ -> softmax -> P·V chain of one
// tile no longer gates the next tile (the limiter of the single-CTA kernel, where S and P of a tile share one buffer).
//
// Warps (576 threads per CTA):
//   0-15 softmax.  Warp w owns TMEM lanes 32(w%4) + 16((w/4)%2) .. +15 (16 query rows) and key half w/8 of every tile
//        (64 keys), read with the 16x256b shape: a query row of that half lives in ONE QUAD (thread t: rows t/4 and
//        t/4+8, key columns 8g + 2(t%4), +1), so row reductions are two shuffles.  P goes back with the 16x128b shape, whose
//        fragment (column 4g + t%4) is exactly the bf16 pair the thread just produced; each key half keeps its P inside its
//        own S columns (first 32 of its 64), so a half never overwrites scores the other half still has to read.
//        The reference maximum of a row is fixed after its first key tile; later tiles only check (on the row sums they
//        compute anyway) whether a probability left the safe range.  The two warps sharing 16 rows agree on that with one
//        64-thread named barrier per tile; only then do they take the slow path together: exchange half-row maxima, wait
//        for the outstanding P·V, rescale their halves of the rows' O columns and their partial sums, recompute the tile.
//        Exact after the final 1/rowsum.  exp2 runs on the MUFU and, for a compile-time share of the pairs, as a
//        Cody-Waite + minimax cubic on the FMA pipe.
//   16   TMA producer (both CTAs; bytes of the pair are credited to the leader's `full` barriers)
//   17   MMA issuer (leader CTA only) + TMEM owner
//
// Two softmax organisations share everything else (MODE template parameter):
//   FAST  (the one that runs): same warp -> (rows, key half) map, but NO in-loop maximum exchange and NO rescale: the row
//         reference maxima are fixed on tile 0 (one 512-thread named barrier, once), every warp then turns its 64-key half
//         of EVERY tile into P with no communication at all, and S_j -> P_j takes one half step (~1000 cycles).  That
//         latency is what three S/P buffers can cover: QK (512) + hand-offs (~700) + softmax + P·V (512) per buffer must
//         fit three tile periods; the earlier organisation (two sets of 8 warps taking whole tiles alternately, two half
//         steps per tile) was latency-bound at ~1290 cycles per tile against a tensor floor of 1024
//         (profiles/r02_attn_pair_v8_ab_and_timeline.log).  A row whose probabilities leave the safe range marks its
//         (batch, head, 256-row) unit dirty ...
//   EXACT ... and the exact kernel (per-tile agreement of the two warps of a row group, in-loop slow path) is launched
//         right behind the fast one: it returns at once for clean units and recomputes dirty ones.  Results are exact either
//         way; with q/k RMS-normalised as in this model the fix-up essentially never has work.
//
// Issuer: two barrier polls per tile instead of four: `v_full[i]` carries V_i AND K_{i+2} (the producer loads what one
// iteration consumes onto one barrier), `p_ready` collects both key halves of P_i.  The tensor pipe's queue is shallow, so
// every poll of the issuer is tensor idle time: with four polls per tile the issuer sat ~25 % of its time in them (ncu).
#include <mutex>
#include <type_traits>
#include "common.cuh"
#include "ptx.cuh"
#include "attention.cuh"

namespace amb {

constexpr int PA_THREADS = 576;
constexpr int PA_BK = 128;   // keys per K/V tile
constexpr int PA_NBUF = 3;   // S/P buffers in TMEM
constexpr float PA_SUM_LIMIT = 4096.0f;  // a thread's partial tile sum (16 keys) above this sends its row group down the slow path

template <int KS, int VS>
struct PairSmem {
  static constexpr int Q_BYTES = 128 * 128 * 2;    // my 128 query rows: two 64-column boxes of 16 KB
  static constexpr int KV_BYTES = 64 * 128 * 2;    // my half of a K tile (64 keys x 128 d) or of a V tile (128 keys x 64 d)
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = Q_BYTES;
  static constexpr int V_OFF = K_OFF + KS * KV_BYTES;
  static constexpr int XCH_OFF = V_OFF + VS * KV_BYTES;      // float[3][128 rows]: row maxima / partial sums exchanged between warps
  static constexpr int FLAG_OFF = XCH_OFF + 3 * 128 * 4;      // int[8 row groups][2 tile parities]: slow-path requests
  static constexpr int BAR_OFF = FLAG_OFF + 8 * 2 * 4;
  static constexpr int NUM_BARS = 1 + 2 * KS + 2 * VS + 3 * PA_NBUF;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
};

// which exponential pairs run on the FMA pipe: EMU = quarters of all pairs (0..3; 4 = one eighth); (g, r) = (column group, row A/B)
__host__ __device__ constexpr bool pa_emulated(int emu, int g, int r) {
  return emu == 0 ? false : emu == 4 ? (r == 0 && (g & 3) == 0) : emu == 1 ? (r == 0 && (g & 1) == 0) : emu == 2 ? (((g + r) & 1) == 0) : !(r == 1 && (g & 1) == 1);
}

enum : int { PA_EXACT = 0, PA_FIXUP = 1, PA_FAST = 2 };  // MODE: exact / exact on dirty units only / fast path (fixed reference maxima)
#ifndef AMB_ATTN_TRACE
#define AMB_ATTN_TRACE 0  // 1: compile the clock64 role timeline (tools/attn_trace.py) into the kernels; off in the product build
#endif

#ifndef PA_EMU
#define PA_EMU 1  // FAST: quarters of the exponential pairs computed on the FMA pipe instead of the MUFU (measured best of 0, 1/8, 1, 2)
#endif

template <int KS, int VS, int EMU, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PA_THREADS, 1)
flash_attn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using L = PairSmem<KS, VS>;
  constexpr bool FAST = (MODE == PA_FAST);
  const int unit = (blockIdx.z * gridDim.y + blockIdx.y) * (gridDim.x >> 1) + (blockIdx.x >> 1);
  if (MODE == PA_FIXUP) {
    if (p.dirty[unit] == 0) return;  // both CTAs of the pair read the same flag: uniform exit before any barrier exists
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xch = reinterpret_cast<float*>(smem + L::XCH_OFF);
  volatile int* flags = reinterpret_cast<volatile int*>(smem + L::FLAG_OFF);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);  // leader's copy is live
  uint64_t* k_full = q_full + 1;          // [KS] leader: K_0 and K_1 only (the prologue's tiles)
  uint64_t* k_empty = k_full + KS;        // [KS] both (multicast commit)
  uint64_t* v_full = k_empty + KS;        // [VS] leader: what iteration i consumes, V_i and K_{i+2}
  uint64_t* v_empty = v_full + VS;        // [VS] both
  uint64_t* s_full = v_empty + VS;        // [3]  both: S_j complete in this CTA's TMEM
  uint64_t* p_ready = s_full + PA_NBUF;   // [3]  leader: P_j written: 32 warp arrivals (16 per CTA)
  uint64_t* pv_done = p_ready + PA_NBUF;  // [3]  both: P_j·V_j (and everything before it) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + PA_NBUF);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int q0 = (blockIdx.x >> 1) * 256 + rank * 128;  // first query row of this CTA
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int tiles_per_chunk = (p.sk_chunk + PA_BK - 1) / PA_BK;
  const int n_kv = p.kv_chunks * tiles_per_chunk;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 17) {
    if (lane == 0) {
      for (int i = 0; i < 16; ++i) flags[i] = 0;
      mbar_init(q_full, 1);
      for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
      for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
      for (int b = 0; b < PA_NBUF; ++b) {
        mbar_init(&s_full[b], 1);
        mbar_init(&p_ready[b], 32);
        mbar_init(&pv_done[b], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit / peer-credited TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ===================== TMA producer (both CTAs: my half of every tile) =====================
    if (elect_one()) {
      if (leader) mbar_expect_tx(q_full, 2 * L::Q_BYTES);
      tma_load_4d_pair(smem + L::Q_OFF, &tmQ, q_full, 0, q0, head, batch, kEvictFirst);
      tma_load_4d_pair(smem + L::Q_OFF + 16384, &tmQ, q_full, 64, q0, head, batch, kEvictFirst);
    }
    __syncwarp();
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    int chunk_k = 0, jj_k = 0, chunk_v = 0, jj_v = 0;  // (chunk, tile in chunk) of the next K tile / V tile
    auto load_k = [&](uint64_t* bar) {  // keys [key0 + 64 rank, +64): two 64-column boxes of 8 KB
      if (elect_one()) {
        uint8_t* sk = smem + L::K_OFF + ks * L::KV_BYTES;
        const int key0 = jj_k * PA_BK + 64 * (int)rank;
        tma_load_5d_pair(sk, &tmK, bar, 0, key0, head, batch, chunk_k, kEvictLast);
        tma_load_5d_pair(sk + 8192, &tmK, bar, 64, key0, head, batch, chunk_k, kEvictLast);
      }
      __syncwarp();
      if (++ks == KS) { ks = 0; kph ^= 1; }
      if (++jj_k == tiles_per_chunk) { jj_k = 0; ++chunk_k; }
    };
    for (int t = 0; t < 2 && t < n_kv; ++t) {  // K_0, K_1 on their own barriers (fresh stages: nothing to wait for)
      if (leader && elect_one()) mbar_expect_tx(&k_full[t], 2 * L::KV_BYTES);
      __syncwarp();
      load_k(&k_full[t]);
    }
    for (int i = 0; i < n_kv; ++i) {  // iteration i of the issuer consumes V_i and K_{i+2}: one barrier for both
      const bool has_k = i + 2 < n_kv;
      mbar_wait(&v_empty[vs], vph ^ 1);
      if (has_k) mbar_wait(&k_empty[ks], kph ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(&v_full[vs], (has_k ? 4 : 2) * L::KV_BYTES);
        uint8_t* sv = smem + L::V_OFF + vs * L::KV_BYTES;  // 128 keys x d-columns [64 rank, +64): one box of 16 KB
        tma_load_5d_pair(sv, &tmV, &v_full[vs], 64 * (int)rank, jj_v * PA_BK, head, batch, chunk_v, kEvictLast);
      }
      __syncwarp();
      if (has_k) load_k(&v_full[vs]);
      if (++vs == VS) { vs = 0; vph ^= 1; }
      if (++jj_v == tiles_per_chunk) { jj_v = 0; ++chunk_v; }
    }
  } else if (warp == 17) {
    if (leader) {
      // ===================== MMA issuer (leader CTA; one elected lane inside a converged warp, so that descriptors and
      // addresses stay in uniform registers: a divergent single-thread loop pays R2UR moves per MMA, ~90 cycles each) =========
      constexpr uint32_t idesc_qk = make_idesc_bf16(256, PA_BK, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(256, 128, 0, 1);
      const uint32_t sq_addr = smem_u32(smem + L::Q_OFF);
      const uint32_t sk_addr = smem_u32(smem + L::K_OFF);
      const uint32_t sv_addr = smem_u32(smem + L::V_OFF);
      const uint32_t o_tmem = tmem_base + PA_NBUF * 128;

      auto issue_qk = [&](int buf, int kstage) {
        const uint64_t qd = make_desc_kmajor_sw128(sq_addr);
        const uint64_t kd = make_desc_kmajor_sw128(sk_addr + kstage * L::KV_BYTES);
        const uint32_t d = tmem_base + buf * 128;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)  // d = 128 in 16-wide steps; the second 64-column box of Q is 16 KB, of K 8 KB further
            mma_ss_pair(d, qd + ((k >> 2) * (16384 / 16) + (k & 3) * 2), kd + ((k >> 2) * (8192 / 16) + (k & 3) * 2),
                        idesc_qk, k != 0);
          tc_commit_pair(&s_full[buf]);
          tc_commit_pair(&k_empty[kstage]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int buf, int vstage, bool first) {
        const uint64_t vd = make_desc_mnmajor_sw128(sv_addr + vstage * L::KV_BYTES, 16384);
        const uint32_t pa = tmem_base + buf * 128;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // 128 keys in 16-key steps: 8 packed P columns (key half h = kk / 4 keeps its P in
                                          // the first 32 of its own 64 S columns), 16 V rows (2 KB) per step
            mma_ts_pair(o_tmem, pa + (kk >> 2) * 64 + (kk & 3) * 8, vd + 128 * kk, idesc_pv, (!first || kk != 0) ? 1u : 0u);
          tc_commit_pair(&v_empty[vstage]);
          tc_commit_pair(&pv_done[buf]);
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      int ks = 0, vs = 0;
      uint32_t vph = 0;
      const int npro = n_kv < 2 ? n_kv : 2;
      for (int j = 0; j < npro; ++j) {
        mbar_wait(&k_full[j], 0);
        tc_fence_after();
        issue_qk(j, ks);
        if (++ks == KS) ks = 0;
      }
      int buf = 0, buf2 = 2 % PA_NBUF;  // buffer of tile j / of tile j + 2
      uint32_t bph = 0;
      const bool tracer = AMB_ATTN_TRACE && lane == 0 && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      for (int j = 0; j < n_kv; ++j) {
        const bool tr = tracer && j >= 100 && j < 116;  // role 4 of the debug timeline
        if (tr) p.trace[(4 * 16 + (j - 100)) * 8 + 0] = clock64();
        mbar_wait(&v_full[vs], vph);  // V_j and K_{j+2}
        if (tr) p.trace[(4 * 16 + (j - 100)) * 8 + 1] = clock64();
        tc_fence_after();
        if (j + 2 < n_kv) {  // S_{j+2} first: the tensor pipe has it queued while the softmax warps work on S_j
          issue_qk(buf2, ks);
          if (++ks == KS) ks = 0;
        }
        if (tr) p.trace[(4 * 16 + (j - 100)) * 8 + 2] = clock64();
        mbar_wait(&p_ready[buf], bph);
        if (tr) p.trace[(4 * 16 + (j - 100)) * 8 + 3] = clock64();
        tc_fence_after();
        issue_pv(buf, vs, j == 0);
        if (tr) p.trace[(4 * 16 + (j - 100)) * 8 + 4] = clock64();
        if (++vs == VS) { vs = 0; vph ^= 1; }
        if (++buf == PA_NBUF) { buf = 0; bph ^= 1; }
        if (++buf2 == PA_NBUF) buf2 = 0;
      }
    }
    __syncwarp();
  } else {
    if constexpr (FAST) {
    // ===================== softmax, FAST: warp = (set, lane quarter, 16-row half); set s takes tiles j = s, s+2, ... ============
    const int set = warp >> 3;
    const int quarter = warp & 3;
    const int lane_base = quarter * 32 + ((warp >> 2) & 1) * 16;
    const uint32_t lane_sel = static_cast<uint32_t>(lane_base) << 16;
    const int row_a = lane_base + (lane >> 2);  // my two rows of this CTA's 128-row tile: row_a and row_a + 8
    const int qd = lane & 3;                    // my key columns inside a group of 8: 2 qd, 2 qd + 1
    const float c = p.scale_log2;
    const uint64_t c2 = pk2(c, c);
    const int last_valid = p.sk_chunk - (tiles_per_chunk - 1) * PA_BK;  // valid keys in the last tile of a chunk
    const bool has_tail = last_valid < PA_BK;
    constexpr uint32_t MREF_BAR = 1, SUM_BAR = 2;  // named barriers: reference maxima published / partial sums published
    // debug timeline (amb_debug_set_attn_trace): roles 0/1 = warp 0 of each set of CTA (0,0,0), 5 x 16 x 8 int64 slots
    const bool tracer = AMB_ATTN_TRACE && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (warp & 7) == 0;

    float m_a = 0.f, m_b = 0.f;  // reference maxima (raw score units) of my two rows: fixed by set 0 on tile 0
    float l_a = 0.f, l_b = 0.f;  // partial row sums: my 32 keys of every tile of my set
    bool dirty = false;

    auto half_step_impl = [&](int j, int buf, int h, auto masked_c) {
      constexpr bool masked = decltype(masked_c)::value;  // tail tiles get their own copy: no key masking in the hot one
      const bool tr = tracer && j >= 100 && j < 116;
      const uint32_t s_addr = tmem_base + buf * 128 + h * 64 + lane_sel;
      float s[32];  // s[4g + {0,1}] = row_a, keys 64 h + 8g + 2qd + {0,1};  s[4g + {2,3}] = row_a + 8, same keys
      tmem_ld16_256b_x8f(s_addr, s);
      tmem_wait_ld();
      if (tr) p.trace[(set * 16 + (j - 100)) * 8 + 2 + 2 * h] = clock64();
      uint32_t pk[16];
      const float mb_a = m_a * c, mb_b = m_b * c;
      const uint64_t nmb_a = pk2(-mb_a, -mb_a), nmb_b = pk2(-mb_b, -mb_b);
      uint64_t sum_a = pk2(0.f, 0.f), sum_b = pk2(0.f, 0.f);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float xa0, xa1, xb0, xb1, ea0, ea1, eb0, eb1;
        upk2(fma2(pk2(s[4 * g], s[4 * g + 1]), c2, nmb_a), xa0, xa1);
        upk2(fma2(pk2(s[4 * g + 2], s[4 * g + 3]), c2, nmb_b), xb0, xb1);
        if (pa_emulated(EMU, g, 0)) {
          exp2_poly2(xa0, xa1, ea0, ea1);
        } else {
          ea0 = ex2_approx(xa0);
          ea1 = ex2_approx(xa1);
        }
        if (pa_emulated(EMU, g, 1)) {
          exp2_poly2(xb0, xb1, eb0, eb1);
        } else {
          eb0 = ex2_approx(xb0);
          eb1 = ex2_approx(xb1);
        }
        if (masked) {
          const int key = h * 64 + 8 * g + 2 * qd;
          if (key >= last_valid) ea0 = eb0 = 0.f;
          if (key + 1 >= last_valid) ea1 = eb1 = 0.f;
        }
        sum_a = add2(sum_a, pk2(ea0, ea1));
        sum_b = add2(sum_b, pk2(eb0, eb1));
        pk[2 * g] = pack_bf16(ea0, ea1);
        pk[2 * g + 1] = pack_bf16(eb0, eb1);
      }
      float x0, x1, y0, y1;
      upk2(sum_a, x0, x1);
      upk2(sum_b, y0, y1);
      const float ts_a = x0 + x1, ts_b = y0 + y1;
      // probabilities outside the safe range show up in the sums computed anyway (inf / NaN included): the unit is redone
      dirty |= !(ts_a < PA_SUM_LIMIT) || !(ts_b < PA_SUM_LIMIT);
      tmem_st16_128b_x8(s_addr, pk);  // this half of P_j over the first 32 of its own 64 S columns
      l_a += ts_a;
      l_b += ts_b;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&p_ready[buf]);
        else mbar_arrive_remote(&p_ready[buf], 0);
      }
      if (tr) p.trace[(set * 16 + (j - 100)) * 8 + 3 + 2 * h] = clock64();
    };
    auto half_step = [&](int j, int buf, int h, bool masked) {
      if (masked) half_step_impl(j, buf, h, std::true_type{});
      else half_step_impl(j, buf, h, std::false_type{});
    };

    // Key-half layout: `set` is the key half.  S_j -> P_j takes one half step (~1000 cycles) instead of two, which is what
    // the three S/P buffers can cover: QK (512) + hand-offs (~700) + softmax + P.V (512) per buffer must fit 3 tile periods.
    int buf = 0, jj = 0;
    uint32_t bph = 0;
    for (int j = 0; j < n_kv; ++j) {
      const bool masked = has_tail && (jj == tiles_per_chunk - 1);
      const bool tr = tracer && j >= 100 && j < 116;
      if (tr) p.trace[(set * 16 + (j - 100)) * 8 + 0] = clock64();
      mbar_wait(&s_full[buf], bph);
      if (tr) p.trace[(set * 16 + (j - 100)) * 8 + 1] = clock64();
      tc_fence_after();
      if (j == 0) {
        // every row's reference maximum is anchored on tile 0: my half's maximum, then the partner warp's through smem
        float mx_a = -INFINITY, mx_b = -INFINITY;
        {
          float s[32];
          tmem_ld16_256b_x8f(tmem_base + set * 64 + lane_sel, s);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float a0 = s[4 * g], a1 = s[4 * g + 1], b0 = s[4 * g + 2], b1 = s[4 * g + 3];
            if (masked) {
              const int key = set * 64 + 8 * g + 2 * qd;
              if (key >= last_valid) a0 = b0 = -INFINITY;
              if (key + 1 >= last_valid) a1 = b1 = -INFINITY;
            }
            mx_a = fmaxf(mx_a, fmaxf(a0, a1));
            mx_b = fmaxf(mx_b, fmaxf(b0, b1));
          }
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        float* x_mine = xch + 128 + set * 128;  // (the partial sums reuse these slots after the key loop)
        const float* x_other = xch + 128 + (set ^ 1) * 128;
        if (qd == 0) {
          x_mine[row_a] = mx_a;
          x_mine[row_a + 8] = mx_b;
        }
        named_bar_sync(MREF_BAR, 512);
        m_a = fmaxf(mx_a, x_other[row_a]);
        m_b = fmaxf(mx_b, x_other[row_a + 8]);
      }
      half_step(j, buf, set, masked);
      if (++buf == PA_NBUF) { buf = 0; bph ^= 1; }
      if (++jj == tiles_per_chunk) jj = 0;
    }
    if (__any_sync(0xffffffffu, dirty) && lane == 0) atomicOr(p.dirty + unit, 1);

    // ---- epilogue: O / rowsum -> bf16 -> global (b, s, h, d); set s normalises columns [64 s, 64 s + 64) of its rows
    {
      const int jl = n_kv - 1;
      mbar_wait(&pv_done[jl % PA_NBUF], (jl / PA_NBUF) & 1);
      tc_fence_after();
      l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
      l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
      l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
      l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
      float* x_mine = xch + 128 + set * 128;  // (the first 128 floats hold the reference maxima)
      const float* x_other = xch + 128 + (set ^ 1) * 128;
      if (qd == 0) {
        x_mine[row_a] = l_a;
        x_mine[row_a + 8] = l_b;
      }
      named_bar_sync(SUM_BAR, 512);
      const float inv_a = 1.0f / (l_a + x_other[row_a]), inv_b = 1.0f / (l_b + x_other[row_a + 8]);
      const int qr_a = q0 + row_a, qr_b = qr_a + 8;
      __nv_bfloat16* obase = p.o + (long long)batch * p.o_stride_b + (long long)head * p.o_stride_h + set * 64 + 2 * qd;
      __nv_bfloat16* orow_a = obase + (long long)qr_a * p.o_stride_s;
      __nv_bfloat16* orow_b = obase + (long long)qr_b * p.o_stride_s;
      float o[32];
      tmem_ld16_256b_x8f(tmem_base + PA_NBUF * 128 + set * 64 + lane_sel, o);
      tmem_wait_ld();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (qr_a < p.sq) *reinterpret_cast<uint32_t*>(orow_a + 8 * g) = pack_bf16(o[4 * g] * inv_a, o[4 * g + 1] * inv_a);
        if (qr_b < p.sq) *reinterpret_cast<uint32_t*>(orow_b + 8 * g) = pack_bf16(o[4 * g + 2] * inv_b, o[4 * g + 3] * inv_b);
      }
    }
    } else {
    // ===================== softmax: warp = (lane quarter, 16-row half, key half) =====================
    const int quarter = warp & 3;
    const int kh = warp >> 3;                                    // my key half of every tile: keys [64 kh, 64 kh + 64)
    const int lane_base = quarter * 32 + ((warp >> 2) & 1) * 16;
    const uint32_t lane_sel = static_cast<uint32_t>(lane_base) << 16;
    const int row_a = lane_base + (lane >> 2);  // my two rows of this CTA's 128-row tile: row_a and row_a + 8
    const int qd = lane & 3;                    // my key columns inside a group of 8: 2 qd, 2 qd + 1
    const int rg = warp & 7;                    // row group: the two warps rg and rg + 8 share 16 rows
    const uint32_t pair_bar = 1 + rg;           // their 64-thread named barrier
    const float c = p.scale_log2;
    const uint64_t c2 = pk2(c, c);
    const int last_valid = p.sk_chunk - (tiles_per_chunk - 1) * PA_BK;  // valid keys in the last tile of a chunk
    const uint32_t o_addr = tmem_base + PA_NBUF * 128 + kh * 64 + lane_sel;   // my 64 columns of my rows' O
    float* x_mine = xch + kh * 128;
    const float* x_other = xch + (kh ^ 1) * 128;
    // debug timeline (amb_debug_set_attn_trace): roles 0/1 = softmax warps 0/8 of CTA (0,0,0), 5 x 16 x 8 int64 slots
    const bool tracer = AMB_ATTN_TRACE && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && rg == 0;
    const int trole = kh;

    float m_a = -INFINITY, m_b = -INFINITY;  // reference maxima (raw score units) of my two rows (same in both warps of a group)
    float l_a = 0.f, l_b = 0.f;              // partial row sums: my 16 keys of every tile

    auto softmax_step = [&](int j, int buf, uint32_t bph, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const bool tr = tracer && j >= 100 && j < 116;
      if (tr) p.trace[(trole * 16 + (j - 100)) * 8 + 0] = clock64();
      mbar_wait(&s_full[buf], bph);
      if (tr) p.trace[(trole * 16 + (j - 100)) * 8 + 1] = clock64();
      tc_fence_after();
      const uint32_t s_addr = tmem_base + buf * 128 + kh * 64 + lane_sel;
      float s[32];  // s[4g + {0,1}] = row_a, keys 64 kh + 8g + 2qd + {0,1};  s[4g + {2,3}] = row_a + 8, same keys
      tmem_ld16_256b_x8f(s_addr, s);
      tmem_wait_ld();
      if (tr) p.trace[(trole * 16 + (j - 100)) * 8 + 2] = clock64();

      uint32_t pk[16];
      float ts_a, ts_b;
      auto exps = [&]() {  // P = 2^(c s - c m) for my 32 scores, packed bf16 in the 16x128b fragment order; tile sums
        const float mb_a = m_a * c, mb_b = m_b * c;
        const uint64_t nmb_a = pk2(-mb_a, -mb_a), nmb_b = pk2(-mb_b, -mb_b);
        uint64_t sum_a = pk2(0.f, 0.f), sum_b = pk2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float xa0, xa1, xb0, xb1, ea0, ea1, eb0, eb1;
          upk2(fma2(pk2(s[4 * g], s[4 * g + 1]), c2, nmb_a), xa0, xa1);
          upk2(fma2(pk2(s[4 * g + 2], s[4 * g + 3]), c2, nmb_b), xb0, xb1);
          if (pa_emulated(EMU, g, 0)) {
            exp2_poly2(xa0, xa1, ea0, ea1);
          } else {
            ea0 = ex2_approx(xa0);
            ea1 = ex2_approx(xa1);
          }
          if (pa_emulated(EMU, g, 1)) {
            exp2_poly2(xb0, xb1, eb0, eb1);
          } else {
            eb0 = ex2_approx(xb0);
            eb1 = ex2_approx(xb1);
          }
          if (MASKED) {
            const int key = kh * 64 + 8 * g + 2 * qd;
            if (key >= last_valid) ea0 = eb0 = 0.f;
            if (key + 1 >= last_valid) ea1 = eb1 = 0.f;
          }
          sum_a = add2(sum_a, pk2(ea0, ea1));
          sum_b = add2(sum_b, pk2(eb0, eb1));
          pk[2 * g] = pack_bf16(ea0, ea1);
          pk[2 * g + 1] = pack_bf16(eb0, eb1);
        }
        float x, y;
        upk2(sum_a, x, y);
        ts_a = x + y;
        upk2(sum_b, x, y);
        ts_b = x + y;
      };

      bool slow = (j == 0);
      if (!slow) {
        exps();
        // probabilities outside the safe range show up in the sums computed anyway (inf / NaN included); both warps of the
        // row group must take the same decision: request flag (double-buffered by tile parity) + 64-thread barrier
        volatile int* flag = flags + rg * 2 + (j & 1);
        if (__any_sync(0xffffffffu, !(ts_a < PA_SUM_LIMIT) || !(ts_b < PA_SUM_LIMIT)) && lane == 0) *flag = 1;
        named_bar_sync(pair_bar, 64);
        slow = (*flag != 0);
      }
      if (slow) {
        // (re)anchor the rows' reference maxima on this tile: half-row max inside the quad, exchange with the other key
        // half through shared memory, then rescale my half of the rows' O columns and my partial sums
        float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float a0 = s[4 * g], a1 = s[4 * g + 1], b0 = s[4 * g + 2], b1 = s[4 * g + 3];
          if (MASKED) {
            const int key = kh * 64 + 8 * g + 2 * qd;
            if (key >= last_valid) a0 = b0 = -INFINITY;
            if (key + 1 >= last_valid) a1 = b1 = -INFINITY;
          }
          mx_a = fmaxf(mx_a, fmaxf(a0, a1));
          mx_b = fmaxf(mx_b, fmaxf(b0, b1));
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        if (qd == 0) {
          x_mine[row_a] = mx_a;
          x_mine[row_a + 8] = mx_b;
        }
        named_bar_sync(pair_bar, 64);
        const float mn_a = fmaxf(m_a, fmaxf(mx_a, x_other[row_a])), mn_b = fmaxf(m_b, fmaxf(mx_b, x_other[row_a + 8]));
        if (j > 0) {
          const float al_a = ex2_approx((m_a - mn_a) * c), al_b = ex2_approx((m_b - mn_b) * c);
          // every P·V issued so far has this row group's P_{j-1} as an input, so it is either complete or about to be: wait
          // for it, then O belongs to the softmax warps until P_j is handed over (P·V_j cannot start without my arrival)
          const int pbuf = buf == 0 ? PA_NBUF - 1 : buf - 1;
          mbar_wait(&pv_done[pbuf], buf == 0 ? (bph ^ 1) : bph);
          tc_fence_after();
          float o[32];
          tmem_ld16_256b_x8f(o_addr, o);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            o[4 * g] *= al_a;
            o[4 * g + 1] *= al_a;
            o[4 * g + 2] *= al_b;
            o[4 * g + 3] *= al_b;
          }
          tmem_st16_256b_x8f(o_addr, o);
          tmem_wait_st();
          l_a *= al_a;
          l_b *= al_b;
          if (kh == 0 && lane == 0) flags[rg * 2 + (j & 1)] = 0;
        }
        m_a = mn_a;
        m_b = mn_b;
        exps();
        named_bar_sync(pair_bar, 64);  // exchange buffer and request flag are reusable; both halves' O columns are rescaled
      }
      if (tr) p.trace[(trole * 16 + (j - 100)) * 8 + 3] = clock64();
      tmem_st16_128b_x8(s_addr, pk);  // my half of P_j over the first 32 of my own 64 S columns (my S reads are complete)
      l_a += ts_a;
      l_b += ts_b;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&p_ready[buf]);
        else mbar_arrive_remote(&p_ready[buf], 0);
      }
      if (tr) p.trace[(trole * 16 + (j - 100)) * 8 + 4] = clock64();
    };

    const bool has_tail = last_valid < PA_BK;
    int jj = 0, buf = 0;
    uint32_t bph = 0;
    for (int j = 0; j < n_kv; ++j) {
      const bool tail = has_tail && (jj == tiles_per_chunk - 1);
      if (++jj == tiles_per_chunk) jj = 0;
      if (tail) softmax_step(j, buf, bph, std::true_type{});
      else softmax_step(j, buf, bph, std::false_type{});
      if (++buf == PA_NBUF) { buf = 0; bph ^= 1; }
    }

    // ---- epilogue: O / rowsum -> bf16 -> global (b, s, h, d); my 64 columns of my rows
    {
      const int lbuf = buf == 0 ? PA_NBUF - 1 : buf - 1;       // buffer of the last tile
      mbar_wait(&pv_done[lbuf], buf == 0 ? (bph ^ 1) : bph);  // its parity
      tc_fence_after();
      l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
      l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
      l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
      l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
      if (qd == 0) {
        x_mine[row_a] = l_a;
        x_mine[row_a + 8] = l_b;
      }
      named_bar_sync(pair_bar, 64);
      const float inv_a = 1.0f / (l_a + x_other[row_a]), inv_b = 1.0f / (l_b + x_other[row_a + 8]);
      const int qr_a = q0 + row_a, qr_b = qr_a + 8;
      __nv_bfloat16* obase = p.o + (long long)batch * p.o_stride_b + (long long)head * p.o_stride_h + kh * 64 + 2 * qd;
      __nv_bfloat16* orow_a = obase + (long long)qr_a * p.o_stride_s;
      __nv_bfloat16* orow_b = obase + (long long)qr_b * p.o_stride_s;
      float o[32];
      tmem_ld16_256b_x8f(o_addr, o);
      tmem_wait_ld();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (qr_a < p.sq) *reinterpret_cast<uint32_t*>(orow_a + 8 * g) = pack_bf16(o[4 * g] * inv_a, o[4 * g + 1] * inv_a);
        if (qr_b < p.sq) *reinterpret_cast<uint32_t*>(orow_b + 8 * g) = pack_bf16(o[4 * g + 2] * inv_b, o[4 * g + 3] * inv_b);
      }
    }
    }
    if (MODE == PA_FIXUP && warp == 0 && lane == 0 && leader) p.dirty[unit] = 0;  // (every thread of the pair has read it)
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA frees TMEM / exits while its peer may still read its shared memory or TMEM
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// Scratch of the fast pass -> exact pass hand-over: one int per unit, zero between launches (the exact pass clears what it
// redoes).  One buffer per (device, stream): launches on different streams never share flags.
constexpr int PA_MAX_UNITS = 1 << 18;
static int* pair_dirty_flags(cudaStream_t stream) {
  struct Slot { int dev; cudaStream_t stream; int* ptr; };
  static Slot slots[64] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& sl : slots) {
    if (sl.ptr && sl.dev == dev && sl.stream == stream) return sl.ptr;
    if (!sl.ptr) {
      if (cudaMalloc(&sl.ptr, PA_MAX_UNITS * sizeof(int)) != cudaSuccess) { sl.ptr = nullptr; return nullptr; }
      if (cudaMemsetAsync(sl.ptr, 0, PA_MAX_UNITS * sizeof(int), stream) != cudaSuccess) return nullptr;  // stream-ordered before its first use
      sl.dev = dev;
      sl.stream = stream;
      return sl.ptr;
    }
  }
  return nullptr;
}

int launch_attn_pair(const amb_attn_args* a, long long* trace, cudaStream_t stream) {
  constexpr int KS = 4, VS = 4;
  using L = PairSmem<KS, VS>;
  CUtensorMap tmQ, tmK, tmV;
  int r = encode_attn_maps(a, 128, 128, 64, 128, &tmQ, &tmK, &tmV);
  if (r) return r;
  AttnParams p = make_attn_params(a, trace);
  dim3 grid(2 * ((a->sq + 255) / 256), a->heads, a->batch);  // cluster dims (2,1,1) are compiled in
  const long long units = (long long)(grid.x / 2) * grid.y * grid.z;
  auto exact = flash_attn_pair_kernel<KS, VS, 1, PA_EXACT>;
  if (units > PA_MAX_UNITS) {  // more units than fix-up flags: the exact organisation alone
    r = ensure_smem_optin(exact, L::TOTAL);
    if (r) return r;
    exact<<<grid, PA_THREADS, L::TOTAL, stream>>>(tmQ, tmK, tmV, p);
    AMB_CHECK_CUDA(cudaGetLastError());
    return AMB_OK;
  }
  p.dirty = pair_dirty_flags(stream);
  AMB_CHECK_ARG(p.dirty != nullptr, "flash_attn: could not allocate the fix-up flags");
  auto fast = flash_attn_pair_kernel<KS, VS, PA_EMU, PA_FAST>;  // one exponential pair in four on the FMA pipe (measured best of 0..2)
  auto fixup = flash_attn_pair_kernel<KS, VS, 1, PA_FIXUP>;
  r = ensure_smem_optin(fast, L::TOTAL);
  if (r) return r;
  r = ensure_smem_optin(fixup, L::TOTAL);
  if (r) return r;
  fast<<<grid, PA_THREADS, L::TOTAL, stream>>>(tmQ, tmK, tmV, p);
  AMB_CHECK_CUDA(cudaGetLastError());
  fixup<<<grid, PA_THREADS, L::TOTAL, stream>>>(tmQ, tmK, tmV, p);  // returns at once for every unit the fast pass left clean
  AMB_CHECK_CUDA(cudaGetLastError());
  return AMB_OK;
}

}  // namespace amb
