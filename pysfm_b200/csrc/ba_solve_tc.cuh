// Large reduced camera systems on the 5th-generation tensor cores:  solve_motion_normal_eqns
// (bundle_adjuster.py:281-312) as a BLOCKED right-looking Cholesky whose trailing update
//     A22 -= L21 L21^T          (87 % of the flops at 2,000 cameras)
// runs as exact INT8 products on tcgen05 (SASS UTCIMMA) with INT32 accumulators in tensor memory,
// fed by TMA tensor-map loads (UTMALDG), instead of FP64 DMMA.  tcgen05 has no FP64 kind, so the
// FP64 operand is split Ozaki-style:
//
//   row i of the panel L21 (K = 64 w columns) is scaled by a power of two, x = L_ik 2^-e_i in (-1, 1),
//   and cut into S signed 7-bit digits:  x = sum_p d_p 2^(-6-7p) + O(2^-7S),  d_p in [-64, 64]  (exact in
//   FP64: slice_panel_kernel).  Then
//        sum_k L_ik L_jk = 2^(e_i+e_j) sum_{p,q} 2^(-12-7(p+q)) sum_k d_p[i,k] d_q[j,k]
//   and every inner sum is an INT8 x INT8 -> INT32 product that the tensor core evaluates EXACTLY
//   (|sum| <= 64*64*K*(level+1) < 2^31).  Pairs of one level l = p+q share a weight and accumulate
//   into ONE TMEM accumulator (S accumulators of 128 x 64 INT32 = 64 S <= 448 of the 512 TMEM
//   columns); levels l >= S are dropped (relative 2^-7S of the row scales: 2^-42 at S = 6).  The
//   epilogue reads the S accumulators back (tcgen05.ld), combines them in FP64 by Horner from the
//   smallest level up (exact scalings, one rounding per level) and subtracts from A in place.
//
// What stays on the FP64 pipe: the panel (diagonal block + the tiles below it, 13 % of the flops
// at w = 8), factored by the SAME dataflow kernel as the small systems, launched on the leading w
// tile columns of the trailing matrix (CholArgs::phase = kPhaseWindow), the forward substitution
// inside it, and the backward substitution (one launch, kPhaseBackward).  Per window:
//
//   window_prep_kernel      ticket reset; saves the right-hand side of the tile row right below the
//                           window (the window launch's last chain task C_w overwrites it with a y_w
//                           that lacks this window's terms)
//   chol_dataflow_kernel    L11, L21 (all rows), L11^-1 tiles, y of the window        [DMMA]
//   slice_panel_kernel      row scales + S INT8 slice matrices of L21 (K-major, the TMA source) and
//                           b[rows below] -= L21 y_window                              [FP64, memory]
//   ozaki_syrk_kernel       A22 -= L21 L21^T, lower triangle, 128 x 64 tiles, persistent,
//                           warp-specialised: TMA producer / MMA issuer / 8 epilogue warps; up to four
//                           slice pairs per tcgen05.mma (N = 256)                      [tcgen05]
//
// Accuracy (tests/test_ozaki_model.py holds a numpy restatement of this arithmetic to FP64):
// S = 6 reproduces the FP64 solve of BA reduced systems to ~1e-11 relative at condition 1e4
// (S = 5: 1e-9, S = 7: FP64 level); the int8 slices and INT32 level sums are bit-exact by
// construction and are tested as such against the numpy model on the GPU.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ba {
namespace tc {

constexpr int kM = 128;          // rows of an output tile (UMMA M, one TMEM lane per row)
constexpr int kN = 64;           // columns of an output tile (UMMA N)
constexpr int kThreads = 320;    // warp 0: TMA producer, warp 1: TMEM owner + MMA issuer, warps 2-9: epilogue
constexpr int kEpilogueThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kMaxWindowTiles = 16;

struct SyrkArgs {
  double* __restrict__ A;             // dense lower, column-major, ld
  const double* __restrict__ scale;   // [ld_pad] 2^(e_i - 6)
  double* __restrict__ status;        // solver status word (2 = a wait ran past its deadline)
  unsigned int* __restrict__ abort;   // != 0: an earlier launch of this solve gave up
  unsigned long long* __restrict__ dbg_time;   // optional [8] role timers, summed over CTAs (clocks): see tc_bench perf
  int* __restrict__ dbg_acc;          // optional [S][ld_pad... ] dump of the raw level sums (tests): see dbg_ld
  int ld, ld_pad;
  int c1;        // first row / column of the trailing matrix (multiple of 128)
  int K;         // panel width = contraction length (multiple of BK)
  int n_nb;      // 64-wide column blocks of the trailing matrix
  int ntiles;
  int stages;
  unsigned long long wait_limit_ns;   // budget of the launch for waiting on its barriers (BA_OPT_SPIN_TIMEOUT_MS)
  int dbg_ld;    // row pitch of dbg_acc (0 = off)
  int dbg_skip;  // microbenchmarks only: 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue work, 8 = drain without the FP64
                 // combination, 16 = no load / store of A (results are garbage)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: false = the launch is being abandoned (this wait or another one ran out of time).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* s_abort, unsigned long long t0,
                                          unsigned long long limit_ns) {
  unsigned int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0u) {
      if (*s_abort) return false;
      if (now_ns() - t0 > limit_ns) {
        *s_abort = 1;
        return false;
      }
    }
  }
  return *s_abort == 0;
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {   // arrives on `bar` when every MMA issued so far by this thread is done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] B[smem]^T, INT8 x INT8 -> INT32, issued by ONE thread for the CTA
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane (lane = 32 (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// exact INT32 -> FP64: the integer lands in the low mantissa bits of 2^52 + 2^31, one subtraction
__device__ __forceinline__ double int_to_double(int x) {
  return __hiloint2double(0x43300000, (int)((unsigned int)x ^ 0x80000000u)) - 4503601774854144.0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile whose rows are BK bytes wide and
// swizzled over BK bytes (cute::UMMA::SmemDescriptor): start address >> 4 | LBO (unused for
// swizzled K-major, 1) << 16 | SBO (8 rows) >> 4 << 32 | version 1 << 46 | layout type << 61.
template <int BK>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  constexpr uint64_t layout = BK == 128 ? 2ull : BK == 64 ? 4ull : 6ull;   // SWIZZLE_128B / 64B / 32B
  constexpr uint64_t sbo = (8ull * BK) >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::i8: D = S32 (2 << 4), A and B signed
// 8-bit (1 << 7, 1 << 10), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t instr_desc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
}
constexpr int kMaxPairs = 4;     // pairs (p, q0 .. q0+3) issued as one product with N = 256

// tile t of the trailing matrix's lower triangle -> (128-row block, 64-column block), both relative
// to c1: row block m holds column blocks 0 .. 2m+1 (the last row block of an odd n_nb one less)
__host__ __device__ __forceinline__ void decode_tile(int t, int& mbl, int& nbl) {
  int m = (int)((sqrt(4.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (m > 0 && m * (m + 1) > t) --m;
  while ((m + 1) * (m + 2) <= t) ++m;
  mbl = m;
  nbl = t - m * (m + 1);
}
__host__ __device__ __forceinline__ int count_tiles(int n_nb) {
  const int mb = (n_nb + 1) / 2;
  return mb * (mb + 1) - (n_nb & 1);
}

#define TC_T0() const long long tc_t0__ = DBG && g.dbg_time ? clock64() : 0
#define TC_ADD(slot) do { if (DBG && g.dbg_time) tc_acc[(slot)] += clock64() - tc_t0__; } while (0)

// DBG = true: the instantiation behind the tests and microbenchmarks (level-sum dump, role timers,
// roles switched off); the product launches DBG = false, where all of that is compiled out.
template <int S, int BK, bool DBG>
__global__ void __launch_bounds__(kThreads, 1)   // 10 warps = 3 on one scheduler's 16K registers: 168 per thread
ozaki_syrk_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const SyrkArgs g) {
  constexpr int A_TILE = kM * BK, B_TILE = kN * BK, STAGE = S * (A_TILE + B_TILE);
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_abort;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operand tiles want 1024-byte alignment (the swizzle pattern is a function of the address bits)
  uint8_t* const smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const unsigned long long t0 = now_ns();
  const int nk = g.K / BK;
  const int skip = DBG ? g.dbg_skip : 0;

  if (threadIdx.x == 0) {
    s_abort = (*g.abort != 0u) ? 1 : 0;
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, kEpilogueThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: one warp allocates (and frees), the base address travels through shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  volatile int* const abortp = &s_abort;
  long long tc_acc[3] = {0, 0, 0};   // role-local timers (DBG)

  if (warp == 0) {
    // ================================ TMA producer ========================================
    if (lane == 0 && *abortp == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
        int mbl, nbl;
        decode_tile(t, mbl, nbl);
        const int rowA = g.c1 + mbl * kM, rowB = g.c1 + nbl * kN;
        bool ok = true;
        for (int kc = 0; kc < nk && ok; ++kc) {
          {
            TC_T0();
            ok = mbar_wait(&empty_bar[stage], phase ^ 1u, abortp, t0, g.wait_limit_ns);
            TC_ADD(0);
          }
          if (!ok) break;
          if (skip & 1) {
            mbar_arrive(&full_bar[stage]);
          } else {
            mbar_expect_tx(&full_bar[stage], (uint32_t)STAGE);
            uint8_t* const sA = smem + (size_t)stage * STAGE;
            uint8_t* const sB = sA + S * A_TILE;
#pragma unroll
            for (int p = 0; p < S; ++p) tma_load_2d(&mapA, &full_bar[stage], sA + p * A_TILE, kc * BK, p * g.ld_pad + rowA);
#pragma unroll
            for (int p = 0; p < S; ++p) tma_load_2d(&mapB, &full_bar[stage], sB + p * B_TILE, kc * BK, p * g.ld_pad + rowB);
          }
          if (++stage == g.stages) { stage = 0; phase ^= 1u; }
        }
        if (!ok) break;
      }
      if (DBG && g.dbg_time) atomicAdd(&g.dbg_time[0], (unsigned long long)tc_acc[0]);   // producer: waiting for a free stage
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==========================================
    if (lane == 0 && *abortp == 0) {
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
        {   // the epilogue has drained the accumulators of the previous tile
          TC_T0();
          const bool okt = mbar_wait(&tmem_empty_bar, tphase ^ 1u, abortp, t0, g.wait_limit_ns);
          TC_ADD(0);
          if (!okt) break;
        }
        tcgen05_fence_after();
        bool ok = true;
        for (int kc = 0; kc < nk; ++kc) {
          {
            TC_T0();
            ok = mbar_wait(&full_bar[stage], phase, abortp, t0, g.wait_limit_ns);
            TC_ADD(1);
          }
          if (!ok) break;
          tcgen05_fence_after();
          const uint32_t aA = smem_u32(smem + (size_t)stage * STAGE);
          const uint32_t aB = aA + S * A_TILE;
#pragma unroll
          for (int ks = 0; ks < BK / 32; ++ks) {
            if (skip & 2) break;
            // Slice p of the row block meets slices q = 0 .. S-1-p of the column block, and pair (p, q)
            // belongs to level p+q: the B tiles of consecutive q lie back to back in shared memory
            // (64 rows each, same row pitch) and the accumulators of consecutive levels lie side by
            // side in TMEM, so up to four pairs go out as ONE product with N = 64 nq.  The A tile is
            // then read from shared memory once per four pairs instead of once per pair (at N = 64 the
            // operand reads, not the tensor pipe, set the pace: 65 clk per product measured against 32).
#pragma unroll
            for (int p = 0; p < S; ++p) {
              const uint64_t da = umma_desc<BK>(aA + p * A_TILE + ks * 32);
#pragma unroll
              for (int q0 = 0; q0 < S - p; q0 += kMaxPairs) {
                const int nq = (S - p - q0) < kMaxPairs ? (S - p - q0) : kMaxPairs;
                const uint64_t db = umma_desc<BK>(aB + q0 * B_TILE + ks * 32);
                // first touch of a level in this tile: its p = 0 product over the first 32 bytes of K
                mma_i8(tmem_base + (uint32_t)((p + q0) * kN), da, db, instr_desc(nq * kN), (kc > 0 || ks > 0 || p > 0) ? 1u : 0u);
              }
            }
          }
          tcgen05_commit(&empty_bar[stage]);   // the stage is free when these products have read it
          if (++stage == g.stages) { stage = 0; phase ^= 1u; }
        }
        if (!ok) break;
        tcgen05_commit(&tmem_full_bar);        // all level sums of the tile are in TMEM
        tphase ^= 1u;
      }
      if (DBG && g.dbg_time) {   // MMA issuer: waiting for the epilogue / for operands
        atomicAdd(&g.dbg_time[1], (unsigned long long)tc_acc[0]);
        atomicAdd(&g.dbg_time[2], (unsigned long long)tc_acc[1]);
      }
    }
  } else {
    // ============ epilogue: 8 warps, thread = (row, half of the tile's columns) ============
    const int quad = warp & 3;                 // a warp reaches TMEM lanes 32 (warp % 4) .. +31 only
    const int half = (warp - 2) >> 2;          // warps 2-5: columns 0-31, warps 6-9: columns 32-63
    const int r = 32 * quad + lane;
    constexpr int NC = kN / 2;                 // columns per thread
    uint32_t fphase = 0;
    for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
      int mbl, nbl;
      decode_tile(t, mbl, nbl);
      const int row = g.c1 + mbl * kM + r, col0 = g.c1 + nbl * kN + half * NC;
      const bool row_ok = row < g.ld;
      double* const Arow = g.A + (size_t)col0 * g.ld + row;
      const double srow = row_ok ? __ldg(g.scale + row) : 0.0;
      // The thread's entries of A are fetched NOW, while the products of the tile run (a load - update -
      // store pass behind the products had 16 loads per thread in flight and took 20,000 clocks per
      // tile); entries above the diagonal are neither read nor written.
      double a[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) a[c] = (row_ok && row >= col0 + c && !(skip & (4 | 16))) ? __ldcg(Arow + (size_t)c * g.ld) : 0.0;
      bool ok;
      {
        TC_T0();
        ok = __all_sync(0xffffffffu, mbar_wait(&tmem_full_bar, fphase, abortp, t0, g.wait_limit_ns));   // (tcgen05.ld is warp-collective)
        TC_ADD(0);
      }
      tcgen05_fence_after();
      const long long tc_p1__ = DBG && g.dbg_time ? clock64() : 0;
      // Drain (holds TMEM): the S level sums of the row, 8 columns at a time -- the loads of the next 8
      // columns are in flight while these are combined in FP64 by Horner in 2^-7 from the smallest level
      // up (exact scalings, one rounding per level) and folded into A:
      // A[row, col] -= 2^(e_row + e_col - 12) v.  INT32 -> FP64 goes through the 2^52 trick (LOP + DADD on
      // the FP64 pipe) instead of I2F.
      if (ok && !(skip & 4)) {
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * quad) << 16) + (uint32_t)(half * NC);
        int acc[2][S][8];
#pragma unroll
        for (int l = 0; l < S; ++l) tmem_ld8(taddr + (uint32_t)(l * kN), acc[0][l]);
#pragma unroll
        for (int c8 = 0; c8 < NC / 8; ++c8) {
          double sc[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) sc[c] = __ldg(g.scale + col0 + c8 * 8 + c) * srow;   // exact: powers of two
          tmem_ld_wait();
          if (c8 + 1 < NC / 8) {
#pragma unroll
            for (int l = 0; l < S; ++l) tmem_ld8(taddr + (uint32_t)(l * kN + (c8 + 1) * 8), acc[(c8 + 1) & 1][l]);
          }
          if (DBG && g.dbg_ld > 0 && row_ok) {
#pragma unroll
            for (int l = 0; l < S; ++l)
#pragma unroll
              for (int c = 0; c < 8; ++c) g.dbg_acc[((size_t)l * g.ld_pad + row) * g.dbg_ld + (col0 + c8 * 8 + c)] = acc[c8 & 1][l][c];
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            double h = 0.0;
            if (skip & 8) {   // (microbenchmark: drain only)
              int x = 0;
#pragma unroll
              for (int l = 0; l < S; ++l) x ^= acc[c8 & 1][l][c];
              h = __hiloint2double(x, x);
            } else {
#pragma unroll
              for (int l = S - 1; l >= 0; --l) h = h * 0.0078125 + int_to_double(acc[c8 & 1][l][c]);
            }
            a[c8 * 8 + c] -= h * sc[c];
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar);   // the next tile's products may overwrite the accumulators now
      fphase ^= 1u;
      if (!ok) break;
      const long long tc_p2__ = DBG && g.dbg_time ? clock64() : 0;
      tc_acc[1] += tc_p2__ - tc_p1__;
      if (!(skip & (4 | 16))) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (row_ok && row >= col0 + c) Arow[(size_t)c * g.ld] = a[c];
      }
      if (DBG && g.dbg_time) tc_acc[2] += clock64() - tc_p2__;
    }
    if (DBG && g.dbg_time && threadIdx.x == 64) {   // epilogue (one thread's view): waiting for the products / drain + combine / stores
      atomicAdd(&g.dbg_time[3], (unsigned long long)tc_acc[0]);
      atomicAdd(&g.dbg_time[4], (unsigned long long)tc_acc[1]);
      atomicAdd(&g.dbg_time[5], (unsigned long long)tc_acc[2]);
    }
  }

  // ---- teardown: everybody is done with TMEM before its owner frees it ----
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && s_abort && *g.abort == 0u) {
    atomicExch(g.abort, 1u);
    *g.status = 2.0;
  }
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Row scales, S INT8 slice matrices (K-major: slices[p][row][k], row pitch K bytes) of the panel
// L[c1.., c0..c1) and the right-looking update of the right-hand side, b[row] -= sum_k L[row,k] y[k].
// One CTA per 16 rows (720 CTAs at 2,000 cameras: the kernel lives on loads in flight, not on
// arithmetic); thread = (row r, one of 16 column groups).
constexpr int kSliceRows = 16;
template <int S>
__global__ void __launch_bounds__(256)
slice_panel_kernel(const double* __restrict__ A, int ld, int c0, int K, double* __restrict__ rhs,
                   const double* __restrict__ saved_rhs, int8_t* __restrict__ slices, size_t slice_stride,
                   double* __restrict__ scale) {
  __shared__ double s_y[64 * kMaxWindowTiles];
  __shared__ double s_mx[16][kSliceRows], s_dot[16][kSliceRows], s_mul[kSliceRows];
  const int tid = threadIdx.x, r = tid & (kSliceRows - 1), gq = tid / kSliceRows;
  const int c1 = c0 + K, R0 = c1 + kSliceRows * blockIdx.x;
  for (int k = tid; k < K; k += 256) s_y[k] = rhs[c0 + k];   // y of the window (written by its chain tasks)
  __syncthreads();
  const double* const col = A + (size_t)c0 * ld + R0 + r;
  {
    double mx = 0.0, dot = 0.0;
#pragma unroll 8
    for (int k = gq; k < K; k += 16) {
      const double v = __ldcg(col + (size_t)k * ld);
      mx = fmax(mx, fabs(v));
      dot = fma(v, s_y[k], dot);
    }
    s_mx[gq][r] = mx;
    s_dot[gq][r] = dot;
  }
  __syncthreads();
  if (tid < kSliceRows) {
    double mx = 0.0, dot = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      mx = fmax(mx, s_mx[q][r]);
      dot += s_dot[q][r];
    }
    // |x| 2^-e < 1  (mx = f 2^e, f in [0.5, 1));  non-finite rows (a failed pivot upstream) get e = 0
    const int e = (mx > 0.0 && mx < 1.7e308) ? ilogb(mx) + 1 : 0;
    s_mul[r] = scalbn(1.0, 6 - e);
    scale[R0 + r] = scalbn(1.0, e - 6);
    // the tile row right below the window: its b was saved before the window launch overwrote it
    const int below = R0 + r - c1;
    const double b = (below < 64 && saved_rhs) ? saved_rhs[below] : rhs[R0 + r];
    rhs[R0 + r] = b - dot;
  }
  __syncthreads();
  const double mul = s_mul[r];
  for (int kc = 0; kc < K; kc += 256) {
    if (kc + 16 * gq >= K) break;   // (K is a multiple of 64, not of 256: whole 16-column groups fall in or out)
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __ldcg(col + (size_t)(kc + 16 * gq + i) * ld) * mul;
    uint32_t pk[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p) pk[p][0] = pk[p][1] = pk[p][2] = pk[p][3] = 0u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double t = v[i];
      if (!(fabs(t) <= 64.0)) t = 0.0;   // NaN / Inf
#pragma unroll
      for (int p = 0; p < S; ++p) {
        const double d = rint(t);
        pk[p][i >> 2] |= ((uint32_t)(int)d & 0xffu) << (8 * (i & 3));
        t = (t - d) * 128.0;
      }
    }
#pragma unroll
    for (int p = 0; p < S; ++p)
      *reinterpret_cast<uint4*>(slices + (size_t)p * slice_stride + (size_t)(R0 + r) * K + kc + 16 * gq) =
          make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
  }
}

// ticket reset of the next window launch + copy of the right-hand side block its last chain task clobbers
__global__ void window_prep_kernel(unsigned int* __restrict__ tickets, const double* __restrict__ rhs_block, double* __restrict__ save) {
  if (threadIdx.x == 0) {
    tickets[0] = 0u;
    tickets[1] = 0u;
  }
  if (rhs_block && threadIdx.x < 64) save[threadIdx.x] = rhs_block[threadIdx.x];
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tensor map over the slice matrices seen as ONE 2-D UINT8 tensor [S * ld_pad rows][K bytes], box = box_rows x bk bytes
inline bool make_slice_map(CUtensorMap* map, const void* base, int K, size_t rows, int bk, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K};
  const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapSwizzle sw = bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int S, int BK>
inline size_t syrk_stage_bytes() { return (size_t)S * (kM + kN) * BK; }

// stages that fit the opt-in shared memory (227 KB minus the static part and the alignment slack)
template <int S, int BK>
inline int syrk_stages() {
  const size_t budget = 232448 - 2048;
  int st = (int)(budget / syrk_stage_bytes<S, BK>());
  if (st > kMaxStages) st = kMaxStages;
  return st;
}

template <int S, int BK, bool DBG>
inline cudaError_t launch_syrk_t(const CUtensorMap& mapA, const CUtensorMap& mapB, SyrkArgs g, int grid, cudaStream_t st, bool* attr_flags) {
  // the shared-memory opt-in is a per-device attribute: remembered per handle, not per process
  bool& attr_set = attr_flags[(S - 4) * 4 + (BK == 128 ? 2 : 0) + (DBG ? 1 : 0)];
  g.stages = syrk_stages<S, BK>();
  if (g.stages < 1) return cudaErrorInvalidConfiguration;
  const size_t smem = (size_t)g.stages * syrk_stage_bytes<S, BK>() + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ozaki_syrk_kernel<S, BK, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  ozaki_syrk_kernel<S, BK, DBG><<<grid, kThreads, smem, st>>>(mapA, mapB, g);
  return cudaGetLastError();
}
template <int S, int BK>
inline cudaError_t launch_syrk_d(const CUtensorMap& mapA, const CUtensorMap& mapB, const SyrkArgs& g, int grid, cudaStream_t st, bool* attr_flags) {
  if (g.dbg_acc || g.dbg_time || g.dbg_skip) return launch_syrk_t<S, BK, true>(mapA, mapB, g, grid, st, attr_flags);
  return launch_syrk_t<S, BK, false>(mapA, mapB, g, grid, st, attr_flags);
}

inline cudaError_t launch_syrk(int S, int bk, const CUtensorMap& mapA, const CUtensorMap& mapB, const SyrkArgs& g, int grid, cudaStream_t st,
                               bool* attr_flags /* [16], per handle */) {
  if (bk == 128) {
    switch (S) {
      case 4: return launch_syrk_d<4, 128>(mapA, mapB, g, grid, st, attr_flags);
      case 5: return launch_syrk_d<5, 128>(mapA, mapB, g, grid, st, attr_flags);
      case 6: return launch_syrk_d<6, 128>(mapA, mapB, g, grid, st, attr_flags);
      case 7: return launch_syrk_d<7, 128>(mapA, mapB, g, grid, st, attr_flags);
    }
  } else if (bk == 64) {
    switch (S) {
      case 4: return launch_syrk_d<4, 64>(mapA, mapB, g, grid, st, attr_flags);
      case 5: return launch_syrk_d<5, 64>(mapA, mapB, g, grid, st, attr_flags);
      case 6: return launch_syrk_d<6, 64>(mapA, mapB, g, grid, st, attr_flags);
      case 7: return launch_syrk_d<7, 64>(mapA, mapB, g, grid, st, attr_flags);
    }
  }
  return cudaErrorInvalidValue;
}

inline cudaError_t launch_slice(int S, const double* A, int ld, int c0, int K, double* rhs, const double* saved_rhs, int8_t* slices,
                                size_t slice_stride, double* scale, cudaStream_t st) {
  const int blocks = (ld - (c0 + K)) / kSliceRows;
  if (blocks <= 0) return cudaSuccess;
  switch (S) {
    case 4: slice_panel_kernel<4><<<blocks, 256, 0, st>>>(A, ld, c0, K, rhs, saved_rhs, slices, slice_stride, scale); break;
    case 5: slice_panel_kernel<5><<<blocks, 256, 0, st>>>(A, ld, c0, K, rhs, saved_rhs, slices, slice_stride, scale); break;
    case 6: slice_panel_kernel<6><<<blocks, 256, 0, st>>>(A, ld, c0, K, rhs, saved_rhs, slices, slice_stride, scale); break;
    case 7: slice_panel_kernel<7><<<blocks, 256, 0, st>>>(A, ld, c0, K, rhs, saved_rhs, slices, slice_stride, scale); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace ba
