// Row-gather GEMM on the 5th-gen tensor cores: tcgen05.mma kind::tf32 with the 3xTF32
// split (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) so the reference's fp32 "DFT as a
// convolution" (GTCRN/STFT_Process.py:316,328) keeps fp32-class accuracy.
//
//   C[m,n] = sum_k A(m,k) * W[n,k],  m = (chunk b, row t),  A(m,k) = A[b*sB + (t+t0)*sT + k]
//
// * A is never framed in memory: a 3-D TMA tensor map with row stride sT (< K: rows overlap)
//   gathers 128-row tiles straight from the padded waveform (STFT) or from runs of R
//   consecutive spectrum frames (ISTFT overlap-add).
// * Three operand forms.  PLANES: operands pre-split into tf32 hi / lo planes by their producers (STFT / ISTFT bases, round-1
//   layers): pure TMA -> smem -> tcgen05.mma.  AF (fp32-A): TMA lands the fp32 activation tile, converter warps split it in shared
//   memory, TMA-store epilogue.  TS (AF with 64-wide tiles): the converters write the hi / lo tiles into TENSOR MEMORY with tcgen05.st
//   and the MMAs take A from there (see the comment at the kernel).
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, then 8 / 12 / 16 epilogue warps (TMEM -> registers -> global or a
//   TMA store) and, in the AF / TS forms, 2 .. 10 converter warps.  Producer and MMA warps walk their loops with all 32 lanes and
//   issue under elect.sync (warp-uniform control flow: no per-instruction operand broadcast).  Persistent over output tiles with
//   two TMEM accumulator buffers so the epilogue of tile i overlaps the main loop of i+1.
//
// bf16 variant (TcPlan::bf16, licensed only where the workload spec says "bf16 matmuls": MossFormer2-SE-48K,
// BASELINE.json configs[2]): one bf16 plane per operand, 64-element K blocks (the same 128-byte swizzled rows),
// tcgen05.mma kind::f16 with fp32 accumulation -- 1 MMA per K step instead of 3 and half the operand bytes.
#include "adn.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "tma_utils.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                 // 32 fp32 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 8;              // tf32: 32 bytes of K per instruction
constexpr int NEPI = 16;                // epilogue warps: four per TMEM lane quarter (the epilogue, not the tensor pipe, limits
                                        // small-K and bf16 GEMMs: 8 warps could not issue fast enough)
constexpr int NCONV_AF = 2;             // fp32-A mode: 2 + NEPI + NCONV_AF warps in all; the epilogue needs fewer (TMA stores), the rest convert
constexpr int NTHREADS = 64 + 32 * NEPI; // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int NTHREADS_AF = NTHREADS + 32 * NCONV_AF;
constexpr uint32_t A_TILE_BYTES = BM * BK * 4;   // 16 KiB

// (mbarrier / TMA wrappers: tma_utils.cuh)
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// elect.sync: one lane of a CONVERGED warp.  Unlike `lane == 0`, ptxas knows that exactly one thread runs the guarded block, so
// the uniform-datapath tcgen05 / TMA instructions inside take their operands from uniform registers directly instead of an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall per instruction (ncu + SASS of the 64-wide GEMMs: ~90 clk of issue overhead per
// MMA against 54 clk of tensor-pipe time -- the MMA warp, not the tensor pipe, bounded every N = 64 GEMM).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand from tensor memory (lane = tile row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46)
// (8 rows * 128 B = 1024) | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1,
// A=TF32 [7,10)=2, B=TF32 [10,13)=2, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with BF16 operands: A=BF16 [7,10)=1, B=BF16 [10,13)=1, D=F32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// erf for the GELU epilogue: Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7) on rcp.approx / ex2.approx -- about half the
// instructions of erff; the epilogue, not the tensor pipe, limits the small-K GEMMs that use it.
__device__ __forceinline__ float fast_erf(float x) {
  const float a = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, a, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * __expf(-a * a);
  return copysignf(r, x);
}

__device__ __forceinline__ int16_t to_i16(float y, int mode) {
  if (mode == 1) return (int16_t)(int)(fminf(fmaxf(y, -1.0f), 32767.0f / 32768.0f) * 32768.0f);
  return (int16_t)(int)fminf(fmaxf(y * 32767.0f, -32768.0f), 32767.0f);
}

// ex2.approx.ftz / lg2.approx.ftz: what __expf / __logf evaluate, without their denormal-range guards (FSETP + two predicated
// FMULs per call).  The guarded forms differ only where the result is denormal, and every use below adds 1 to it.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Activation over the 32 registers of an epilogue chunk, the kind fixed at compile time (fp32-A epilogue: one dispatch per chunk).
template <int ACT>
__device__ __forceinline__ void act_all(float (&x)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (ACT == ACT_SWOOSH_L || ACT == ACT_SWOOSH_R) {
      // softplus(y) - 0.08 x, y = x - 4 | 1: max(y, 0) + ln 2 * lg2(1 + 2^(-|y| log2 e)); the log argument is in (1, 2]
      const float y = x[j] - (ACT == ACT_SWOOSH_L ? 4.0f : 1.0f);
      x[j] = fmaf(lg2_ftz(1.0f + ex2_ftz(fabsf(y) * -1.4426950408889634f)), 0.69314718055994531f, fmaxf(y, 0.f)) - 0.08f * x[j];
    } else if (ACT == ACT_SILU) {
      x[j] = x[j] * __fdividef(1.0f, 1.0f + ex2_ftz(x[j] * -1.4426950408889634f));
    } else if (ACT == ACT_RELU) {
      x[j] = fmaxf(x[j], 0.f);
    } else if (ACT == ACT_RELU2) {
      const float r = fmaxf(x[j], 0.f);
      x[j] = r * r;
    } else if (ACT == ACT_GELU) {
      x[j] = 0.5f * x[j] * (1.0f + fast_erf(x[j] * 0.70710678118654752440f));
    } else if (ACT == ACT_TANH) {
      x[j] = tanhf(x[j]);
    } else if (ACT == ACT_SIGMOID) {
      x[j] = 1.0f / (1.0f + expf(-x[j]));
    }
  }
}

constexpr int NCONV_TS = 8;             // TS mode: two converter warps per TMEM lane quarter
constexpr int NE_TS = 8;
constexpr int NTHREADS_TS = 64 + 32 * (NE_TS + NCONV_TS);

template <int BN, bool BF = false, bool AF = false, bool TS = false>
struct Smem {
  static constexpr uint32_t W_TILE_BYTES = BN * BK * 4;           // BN rows of 128 bytes (32 tf32 or 64 bf16)
  static constexpr int PLANES = BF ? 1 : 2;
  // TS: the fp32 A tile as TMA lands it (its tf32 hi / lo operand tiles live in tensor memory) + the W hi / lo tiles
  static constexpr uint32_t STAGE_BYTES = TS ? A_TILE_BYTES + 2 * W_TILE_BYTES : PLANES * (A_TILE_BYTES + W_TILE_BYTES);
  static constexpr int STAGES = TS ? 6 : BF ? 4 : ((BN <= 64) ? 4 : (BN <= 128) ? 3 : 2);
  // TS tensor-memory map: two BN-column accumulators, then STAGES operand slots of 32 hi + 32 lo columns
  static constexpr int ACC_COLS = TS ? BN : 256;
  static constexpr int A_COL0 = 2 * BN;
  static_assert(!TS || 2 * BN + STAGES * 2 * BK <= 512, "TS mode: accumulators + A operand slots exceed tensor memory");
  // epilogue staging behind the pipeline stages (1024-byte aligned): per-warp 8x36 transpose tiles, or, in fp32-A mode, one
  // 16-row x 128-byte swizzled tile per warp that a TMA store drains
  static constexpr uint32_t EPI_STAGE_BYTES = AF ? NEPI * 2048 : 19456 /* >= NEPI * 8 * 36 * 4, multiple of 1024 */;
  static constexpr uint32_t TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
};

// AF (fp32 A operand, ZipEnhancer): the A tensor map is over the fp32 activations themselves.  TMA lands the fp32 tile in the
// hi slot of the stage; the converter warps (6 or 10 of the 20 warps; 640 threads keep the 96-register budget) split every element in place into
// hi = x & 0xFFFFE000 (kept where it is) and lo = x - hi (lo slot) -- the 128-byte swizzle is a permutation of 16-byte chunks,
// so an element-wise pass needs no knowledge of it -- then fence.proxy.async and hand the stage to the MMA warp.  Activations
// live in HBM once, as fp32: half the A bytes, and no producer writes operand planes.
//
// TS (with AF, 64-wide tiles): the A operand of the MMAs comes from TENSOR MEMORY.  ncu on the SS form of the 64-wide implicit-GEMM
// convolutions (profiles/r2e_zip_dense_ss_*): one k block took 1 100 clk against 384 clk of MMA issue (12 x 128x64x8), because
// every MMA streams its 4 KB A tile + 2 KB W tile through the shared-memory port (72 KB per k block = 576 clk at 128 B/clk) and the
// converter warps' loads / stores (395 wavefronts per k block) share that port.  Here the converters (8 warps, lane = tile row)
// read the fp32 tile once and write hi / lo with tcgen05.st into per-stage operand columns; the MMAs read only W from shared
// memory (16 clk per MMA, under the 32 clk issue floor), the stage shrinks to 32 KB (six stages) and nothing is written back to
// shared memory.  Same products in the same order as the SS form: results are bit-identical (tests/test_gpu_ends.py).
template <int BN, int EPI, bool BF, bool AF, bool DK = false, bool TS = false>
__global__ void __launch_bounds__(TS ? NTHREADS_TS : AF ? NTHREADS_AF : NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
               const __grid_constant__ CUtensorMap map_c, const TcArgs g) {
  using S = Smem<BN, BF, AF, TS>;
  static_assert(!TS || (AF && !BF), "TS mode is the fp32-A path");
  constexpr int BKE = BF ? 2 * BK : BK;          // K elements per 128-byte row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S::STAGES * S::STAGE_BYTES + S::EPI_STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + S::STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * S::STAGES;   // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint64_t* conv = tempty + 2;              // [STAGES] (AF only)
  uint32_t* tmem_slot = (uint32_t*)(conv + S::STAGES);
  // epilogue warps 2 .. 2+NE-1, converter warps behind them.  fp32-A mode: a 64-wide tile has only 2 column chunks x 4 lane
  // quarters = 8 epilogue tasks, so 10 warps convert operands there (what bounds the deep-K implicit-GEMM convolutions: 19.7 ->
  // 16.1 ms per step); wider tiles keep all 16 epilogue warps (12 measured slower: ff_in 8.2 -> 10.0 ms) and 2 converters.
  // DK (deep K, >= 512: MossFormer2 fl_in): the epilogue is amortised over many K blocks, conversion is not: 12 epilogue + 6
  // converter warps (fl_in 8.3 -> 7.6 ms per step; at K = 256 the 16-warp epilogue still wins).
  constexpr int NE = TS ? NE_TS : AF ? (BN <= 64 ? 8 : DK ? 12 : NEPI) : NEPI;
  constexpr int NCONV = TS ? NCONV_TS : AF ? NEPI + NCONV_AF - NE : 0;
  float* stage = (float*)(smem + S::STAGES * S::STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = (g.N + BN - 1) / BN;
  const int n_tiles = g.m_tiles * n_tiles_n;
  const int k_blocks = (g.K + BKE - 1) / BKE;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi);
    prefetch_tmap(&map_a_lo);
    prefetch_tmap(&map_w_hi);
    prefetch_tmap(&map_w_lo);
    if (g.k_split > 0) { prefetch_tmap(&map_w2_hi); prefetch_tmap(&map_w2_lo); }
    if (AF) prefetch_tmap(&map_c);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], TS ? 1 + NCONV : 1);   // TS: the converters release the A tile, the MMA commit the W tiles
      if (AF) mbar_init(&conv[i], NCONV);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], NE);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (all lanes walk the loops, one elected lane issues: see elect_one)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile / n_tiles_n, nt = tile - mt * n_tiles_n;
        int b0, t0;
        if (g.bb > 1) { b0 = mt * g.bb; t0 = g.t0; }
        else { b0 = mt / g.tiles_per_chunk; t0 = g.t0 + (mt - b0 * g.tiles_per_chunk) * BM; }
        b0 += g.b_off;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * S::STAGE_BYTES;
          if (elect_one()) {
          const bool skip_w = AF && (g.probe & 4) && !(tile == (int)blockIdx.x && kb < S::STAGES);
          mbar_expect_tx(&full[stage], (uint32_t)(AF ? 1 : S::PLANES) * (uint32_t)(g.bt * g.bb * BK * 4) + (skip_w ? 0u : (uint32_t)S::PLANES * S::W_TILE_BYTES));
          int wb = g.w_batched ? b0 : 0;            // per-batch weights (mask-estimator bands, attention operands)
          int wk = kb * BKE;
          if (g.w_group > 1) { wk += (b0 % g.w_group) * g.w_kstep; wb = b0 / g.w_group; }   // FLASH group of a window
          const bool second = g.k_split > 0 && kb * BKE >= g.k_split;   // second operand of a K-concatenated product
          if (second) wk = kb * BKE - g.k_split;
          int ak = kb * BKE, at = t0;
          if (g.taps > 0) { const int tap = kb / g.tap_kb; ak = g.tap_k0 + (kb - tap * g.tap_kb) * BKE; at = t0 + g.tap_shift[tap]; }
          if (TS) {
            tma_load_3d(st, &map_a_hi, &full[stage], ak, at, b0);
            if (!skip_w) {
              tma_load_3d(st + A_TILE_BYTES, second ? &map_w2_hi : &map_w_hi, &full[stage], wk, nt * BN, wb);
              tma_load_3d(st + A_TILE_BYTES + S::W_TILE_BYTES, second ? &map_w2_lo : &map_w_lo, &full[stage], wk, nt * BN, wb);
            }
          } else if (BF) {
            tma_load_3d(st, &map_a_hi, &full[stage], ak, at, b0);
            tma_load_3d(st + A_TILE_BYTES, second ? &map_w2_hi : &map_w_hi, &full[stage], wk, nt * BN, wb);
          } else {
            tma_load_3d(st, &map_a_hi, &full[stage], ak, at, b0);
            if (!AF) tma_load_3d(st + A_TILE_BYTES, &map_a_lo, &full[stage], ak, at, b0);
            if (!skip_w) {
            tma_load_3d(st + 2 * A_TILE_BYTES, second ? &map_w2_hi : &map_w_hi, &full[stage], wk, nt * BN, wb);
            tma_load_3d(st + 2 * A_TILE_BYTES + S::W_TILE_BYTES, second ? &map_w2_lo : &map_w_lo, &full[stage], wk, nt * BN, wb);
            }
          }
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // All 32 lanes walk the loops (warp-uniform control flow: stage / phase / descriptors stay in uniform registers); one
    // elected lane issues the MMAs and commits of a k block.
    {
      constexpr uint32_t idesc = BF ? make_idesc_bf16(BM, BN) : make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[ab], aphase ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * S::ACC_COLS;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          if (AF) mbar_wait(&conv[stage], phase);  // operand tiles split by the converter warps
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          if (elect_one()) {
            if (TS) {
              const uint32_t a_hi = tmem_base + S::A_COL0 + stage * 2 * BK, a_lo = a_hi + BK;
              const uint64_t w_hi = make_desc(sa + A_TILE_BYTES), w_lo = make_desc(sa + A_TILE_BYTES + S::W_TILE_BYTES);
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);   // +32 B inside the swizzle row
                umma_tf32_ts(d_tmem, a_lo + kk * UMMA_K, w_hi + adv, idesc, (kb | kk) ? 1u : 0u);
                umma_tf32_ts(d_tmem, a_hi + kk * UMMA_K, w_lo + adv, idesc, 1u);
                umma_tf32_ts(d_tmem, a_hi + kk * UMMA_K, w_hi + adv, idesc, 1u);
              }
            } else if (BF) {
              const uint64_t a_d = make_desc(sa), w_d = make_desc(sa + A_TILE_BYTES);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {       // 4 x 16 bf16 = one 128-byte row
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                umma_bf16(d_tmem, a_d + adv, w_d + adv, idesc, (kb | kk) ? 1u : 0u);
              }
            } else {
              const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_TILE_BYTES);
              const uint64_t w_hi = make_desc(sa + 2 * A_TILE_BYTES);
              const uint64_t w_lo = make_desc(sa + 2 * A_TILE_BYTES + S::W_TILE_BYTES);
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);   // +32 B inside the swizzle row
                umma_tf32(d_tmem, a_lo + adv, w_hi + adv, idesc, (kb | kk) ? 1u : 0u);
                umma_tf32(d_tmem, a_hi + adv, w_lo + adv, idesc, 1u);
                umma_tf32(d_tmem, a_hi + adv, w_hi + adv, idesc, 1u);
              }
            }
            umma_commit(&empty[stage]);              // smem slot free once these MMAs retire
            if (kb == k_blocks - 1) umma_commit(&tfull[ab]);   // accumulator complete
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (TS && warp >= 2 + NE) {
    // ===================== fp32 -> tf32 hi / lo converters, operands into tensor memory =====================
    // lane = tile row (a warp may touch the TMEM lane quarter warp % 4 only); the two warps of a quarter take 16 K columns each.
    // The swizzled row is read as four 16-byte chunks: a quarter-warp's eight rows hit eight distinct chunk positions, so the
    // loads are conflict-free.
    const int q = warp & 3, h = (warp - (2 + NE)) >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_dst = tmem_base + ((uint32_t)(q * 32) << 16) + S::A_COL0 + 16 * h;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mt = tile / n_tiles_n;
      const int cb = mt / g.tiles_per_chunk, t = (mt - cb * g.tiles_per_chunk) * BM + r;   // chunk and row of this lane
      float mean = 0.f, rstd = 1.f;
      if (g.a_rowstat && t < g.TM) {
        const float2 ms = __ldg(reinterpret_cast<const float2*>(g.a_rowstat) + ((long long)(cb + g.b_off) * g.TM + t));
        mean = ms.x; rstd = ms.y;
      }
      const int tcol = (g.taps > 0 && g.tap_w > 0) ? t % g.tap_w : 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        const uint8_t* row = smem + stage * S::STAGE_BYTES + r * 128;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(row + (((4 * h + i) ^ (r & 7)) << 4));
        bool zero = false;
        if (g.taps > 0 && g.tap_w > 0) {
          const int col = tcol + g.tap_df[kb / g.tap_kb];
          zero = col < 0 || col >= g.tap_w;
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float y = x[j];
            if (g.a_rowstat) y = (y - mean) * rstd;
            if (zero) y = 0.f;
            const uint32_t hb = __float_as_uint(y) & 0xFFFFE000u;
            hi[4 * i + j] = hb;
            lo[4 * i + j] = __float_as_uint(y - __uint_as_float(hb));
          }
        }
        if (!(g.probe & 1)) {
          tmem_st16(t_dst + stage * 2 * BK, hi);
          tmem_st16(t_dst + stage * 2 * BK + BK, lo);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&conv[stage]);       // operand columns written: the MMA warp may issue
          mbar_arrive(&empty[stage]);      // fp32 tile consumed: with the MMA commit this frees the stage for the producer
        }
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (AF && warp >= 2 + NE) {
    // ===================== fp32 -> tf32 hi / lo converters =====================
    const int cw = warp - (2 + NE);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mt = tile / n_tiles_n;
      const int cb = mt / g.tiles_per_chunk, ct = (mt - cb * g.tiles_per_chunk) * BM;     // chunk and first row of this tile
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        float4* hi = reinterpret_cast<float4*>(smem + stage * S::STAGE_BYTES);
        float4* lo = hi + A_TILE_BYTES / 16;
        const int df = (g.taps > 0 && g.tap_w > 0) ? g.tap_df[kb / g.tap_kb] : 0;
        const int n16 = (g.probe & 1) ? 0 : (int)(A_TILE_BYTES / 16);
        if (!(g.a_rowstat || df)) {                // the common case: split only
#pragma unroll 4
          for (int i = cw * 32 + lane; i < n16; i += NCONV * 32) {
            const float4 v = hi[i];
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
            hi[i] = h;
            lo[i] = l;
          }
        } else {                                   // per-row work: the swizzle permutes 16-byte chunks inside a row only
#pragma unroll 2
          for (int i = cw * 32 + lane; i < n16; i += NCONV * 32) {
            float4 v = hi[i];
            const int t = ct + (i >> 3);
            if (g.a_rowstat && t < g.TM) {
              const float2 ms = __ldg(reinterpret_cast<const float2*>(g.a_rowstat) + ((long long)(cb + g.b_off) * g.TM + t));
              v.x = (v.x - ms.x) * ms.y; v.y = (v.y - ms.x) * ms.y; v.z = (v.z - ms.x) * ms.y; v.w = (v.w - ms.x) * ms.y;
            }
            if (df) {
              const int col = t % g.tap_w + df;
              if (col < 0 || col >= g.tap_w) v = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
            hi[i] = h;
            lo[i] = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[stage]);
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int cidx = (warp - 2) >> 2;              // the four warps of a quarter take every fourth 32-column chunk
    const int r = q * 32 + lane;                   // row of the tile
    const bool store_lane = elect_one();           // issues (and waits for) this warp's TMA stores: bulk groups are per thread
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / n_tiles_n, nt = tile - mt * n_tiles_n;
      const int ab = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int tile_b = mt / g.tiles_per_chunk, tile_t = (mt - tile_b * g.tiles_per_chunk) * BM;   // bb == 1 tiles
      int b, t;
      if (g.bb > 1) { int bi = r / g.bt; b = mt * g.bb + bi; t = r - bi * g.bt; if (bi >= g.bb) b = g.B; }
      else { b = tile_b; t = tile_t + r; }
      b += g.b_off;
      const bool row_ok = (b < g.B) && (t < g.TM);
      // TS tiles are one chunk per epilogue warp: the residual / bypass rows of the chunk (one 128-byte line per lane each) are
      // prefetched into L1 BEFORE the accumulator wait, so their DRAM / L2 latency is off the per-tile chain of the two epilogue
      // warps a scheduler has (register prefetch of the 32 values spills at the 96-register budget)
      if (TS && row_ok && !(g.probe & 2)) {
        const long long op = ((long long)b * g.TM + t) * g.ldc + nt * BN + cidx * 32;
        if (nt * BN + cidx * 32 < g.N) {
          if (g.resid) asm volatile("prefetch.global.L1 [%0];" ::"l"(g.resid + op));
          if (g.resid2) asm volatile("prefetch.global.L1 [%0];" ::"l"(g.resid2 + op));
        }
      }
      mbar_wait(&tfull[ab], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * S::ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = cidx * 32; c0 < BN; c0 += 32 * (NE / 4)) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        const int n0 = nt * BN + c0;
        if (AF) {
          // fp32-only outputs through a TMA store.  The TMEM load gives lane = row with 32 consecutive columns in registers;
          // bias / activation / residuals are applied there, the 32 x 32 block goes -- 16 rows at a time -- into this warp's
          // swizzled 16 x 128-byte staging tile (conflict-free 16-byte writes), and one lane issues cp.async.bulk.tensor
          // (shared -> global): full 128-byte row segments leave the SM asynchronously, rows >= TM and columns >= N are
          // clipped by the tensor map, and the LSU sees no global store at all.  (Measured with ADN_TC_PROBE=2: per-lane
          // st.global epilogues -- row-streaming or transposed -- were 70 % of the time of every small-K GEMM.)
          const bool dead = (g.probe & 2) != 0;
          const long long o = ((long long)b * g.TM + t) * g.ldc + n0;
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          // Valid float4 groups of this chunk (N, BN multiples of 4).  Every optional operand and the activation are dispatched ONCE
          // per chunk, each as its own pass over the 32 registers: ncu on the per-group form (profiles/r2e_zip_ff_*) showed 875
          // warp instructions per chunk, 384 of them the activation itself, at 9.4 clk per issued instruction -- the small-K
          // GEMMs were bound by this instruction stream, not by the stores.
          const int nv = min(8, min(BN - c0, g.N - n0) >> 2);
          if (row_ok && !dead && nv > 0) {
            if (g.rowscale) {
              const float rsc_row = __ldg(g.rowscale + (long long)b * g.TM + t);
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] *= rsc_row;
            }
            // The residual of a plain projection (no activation: every *_out GEMM) is fetched before the bias pass so its latency
            // hides behind it; behind an activation it is fetched afterwards -- 32 more live registers across the activation
            // pass would spill at the 96-register budget of 5 warps per scheduler.
            if (g.act == ACT_NONE && !TS) {
              float4 res4[8];
              if (g.resid) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                  res4[j4] = j4 < nv ? *reinterpret_cast<const float4*>(g.resid + o + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
              if (g.bias) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  if (j4 < nv) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n0) + j4);
                    x[4 * j4] += b4.x; x[4 * j4 + 1] += b4.y; x[4 * j4 + 2] += b4.z; x[4 * j4 + 3] += b4.w;
                  }
                }
              }
              if (g.resid) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  x[4 * j4] += res4[j4].x; x[4 * j4 + 1] += res4[j4].y; x[4 * j4 + 2] += res4[j4].z; x[4 * j4 + 3] += res4[j4].w;
                }
              }
            } else {
              if (g.bias) {
  #pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  if (j4 < nv) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n0) + j4);
                    x[4 * j4] += b4.x; x[4 * j4 + 1] += b4.y; x[4 * j4 + 2] += b4.z; x[4 * j4 + 3] += b4.w;
                  }
                }
              }
              switch (g.act) {                       // (columns >= N carry garbage through the activation: the store clips them)
                case ACT_SWOOSH_L: act_all<ACT_SWOOSH_L>(x); break;
                case ACT_SWOOSH_R: act_all<ACT_SWOOSH_R>(x); break;
                case ACT_SILU: act_all<ACT_SILU>(x); break;
                case ACT_RELU: act_all<ACT_RELU>(x); break;
                case ACT_RELU2: act_all<ACT_RELU2>(x); break;
                case ACT_GELU: act_all<ACT_GELU>(x); break;
                case ACT_TANH: act_all<ACT_TANH>(x); break;
                case ACT_SIGMOID: act_all<ACT_SIGMOID>(x); break;
                case ACT_PRELU: {
                  const float slope = __ldg(g.act_param);
  #pragma unroll
                  for (int j = 0; j < 32; ++j) x[j] = x[j] >= 0.f ? x[j] : slope * x[j];
                } break;
                case ACT_PRELU_VEC: {
  #pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {
                    if (j4 < nv) {
                      const float4 s4 = __ldg(reinterpret_cast<const float4*>(g.act_vec + n0) + j4);
                      x[4 * j4] = x[4 * j4] >= 0.f ? x[4 * j4] : s4.x * x[4 * j4];
                      x[4 * j4 + 1] = x[4 * j4 + 1] >= 0.f ? x[4 * j4 + 1] : s4.y * x[4 * j4 + 1];
                      x[4 * j4 + 2] = x[4 * j4 + 2] >= 0.f ? x[4 * j4 + 2] : s4.z * x[4 * j4 + 2];
                      x[4 * j4 + 3] = x[4 * j4 + 3] >= 0.f ? x[4 * j4 + 3] : s4.w * x[4 * j4 + 3];
                    }
                  }
                } break;
                default: break;
              }
              if (g.resid) {       // (TS: the rows were prefetched into L1 before the accumulator wait; four float4 in flight at a time
                                   //  -- ncu showed the eight-deep form spilling a freshly loaded register, i.e. waiting on the load at once)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  float4 r4[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    r4[i] = (4 * hh + i) < nv ? *reinterpret_cast<const float4*>(g.resid + o + 4 * (4 * hh + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const int j4 = 4 * hh + i;
                    x[4 * j4] += r4[i].x; x[4 * j4 + 1] += r4[i].y; x[4 * j4 + 2] += r4[i].z; x[4 * j4 + 3] += r4[i].w;
                  }
                }
              }
            }
            if (g.resid2) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                if (j4 < nv) {
                  const float4 o4 = *reinterpret_cast<const float4*>(g.resid2 + o + 4 * j4);
                  const float4 s4 = __ldg(reinterpret_cast<const float4*>(g.colscale + n0) + j4);
                  x[4 * j4] = o4.x + (x[4 * j4] - o4.x) * s4.x; x[4 * j4 + 1] = o4.y + (x[4 * j4 + 1] - o4.y) * s4.y;
                  x[4 * j4 + 2] = o4.z + (x[4 * j4 + 2] - o4.z) * s4.z; x[4 * j4 + 3] = o4.w + (x[4 * j4 + 3] - o4.w) * s4.w;
                }
              }
            }
          }
          if (n0 >= g.N || dead) continue;
          if (TS) {
            // eight epilogue warps leave 4 KB of staging per warp: the whole 32 x 32 chunk is staged at once (all lanes, still
            // conflict-free), the two 16-row boxes leave as one bulk group, and the wait for the TMA engine to have READ the
            // tile moves to the next tile -- behind that tile's accumulator wait, TMEM load and arithmetic
            uint8_t* stg4 = reinterpret_cast<uint8_t*>(stage) + (warp - 2) * 4096;
            if (store_lane) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(stg4 + lane * 128 + ((j4 ^ (lane & 7)) << 4)) = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (store_lane) {
              tma_store_3d(&map_c, stg4, n0, tile_t + q * 32, tile_b + g.b_off);
              tma_store_3d(&map_c, stg4 + 2048, n0, tile_t + q * 32 + 16, tile_b + g.b_off);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            continue;
          }
          uint8_t* stg = reinterpret_cast<uint8_t*>(stage) + (warp - 2) * 2048;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if ((lane >> 4) == half) {
              const int rr = lane & 15;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(stg + rr * 128 + ((j4 ^ (rr & 7)) << 4)) = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (store_lane) {
              tma_store_3d(&map_c, stg, n0, tile_t + q * 32 + half * 16, tile_b + g.b_off);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the staging tile may be overwritten
            }
            __syncwarp();
          }
          continue;
        }
        if (EPI == EPI_LIN || EPI == EPI_ISTFT) {
          // Linear layer: v = act(acc * rowscale[m] + bias[n]) (+ residual) -> fp32 and/or tf32 planes.
          // The TMEM load gives lane = row; a per-warp smem transpose turns that into lane = column
          // quad so every global access is a full 128-byte row segment (8 lanes x 16 B, 4 rows/instr).
          float* stg = stage + (warp - 2) * (8 * 36);
          const int cq = lane & 7;
          const bool col_ok = (c0 + 4 * cq < BN) && (n0 + 4 * cq < g.N);      // N, BN, ldc multiples of 4
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          bool have_bias = false;
          // residual / row-scale operands of this 32x32 chunk are fetched up front (8 independent loads
          // in flight) instead of one dependent DRAM round trip per 4-row step
          float4 res4[4][2];
          float rsc[4][2];
#pragma unroll
          for (int hrow = 0; hrow < 4; ++hrow)
#pragma unroll
            for (int it = 0; it < 2; ++it) { res4[hrow][it] = make_float4(0.f, 0.f, 0.f, 0.f); rsc[hrow][it] = 1.0f; }
          if (EPI == EPI_LIN && (g.resid || g.rowscale)) {
#pragma unroll
            for (int hrow = 0; hrow < 4; ++hrow)
#pragma unroll
              for (int it = 0; it < 2; ++it) {
                const int r2 = q * 32 + hrow * 8 + it * 4 + (lane >> 3);
                int b2, t2;
                if (g.bb > 1) { int bi = r2 / g.bt; b2 = mt * g.bb + bi; t2 = r2 - bi * g.bt; if (bi >= g.bb) b2 = g.B; }
                else { b2 = tile_b; t2 = tile_t + r2; }
                b2 += g.b_off;
                const bool ok = col_ok && b2 < g.B && t2 < g.TM;
                const long long m = (long long)b2 * g.TM + t2;
                if (ok && g.resid) res4[hrow][it] = *reinterpret_cast<const float4*>(g.resid + m * g.ldc + n0 + 4 * cq);
                if (ok && g.rowscale) rsc[hrow][it] = __ldg(g.rowscale + m);
              }
          }
#pragma unroll
          for (int hrow = 0; hrow < 4; ++hrow) {             // 8 rows at a time through an 8x36 tile
            if ((lane >> 3) == hrow) {
#pragma unroll
              for (int j0 = 0; j0 < 8; ++j0)
                *reinterpret_cast<float4*>(stg + (lane & 7) * 36 + 4 * j0) =
                    make_float4(__uint_as_float(v[4 * j0]), __uint_as_float(v[4 * j0 + 1]),
                                __uint_as_float(v[4 * j0 + 2]), __uint_as_float(v[4 * j0 + 3]));
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int rl = it * 4 + (lane >> 3);
              const int r2 = q * 32 + hrow * 8 + rl;
              int b2, t2;
              if (g.bb > 1) { int bi = r2 / g.bt; b2 = mt * g.bb + bi; t2 = r2 - bi * g.bt; if (bi >= g.bb) b2 = g.B; }
              else { b2 = tile_b; t2 = tile_t + r2; }
              b2 += g.b_off;
              if (!col_ok || b2 >= g.B || t2 >= g.TM) continue;
              if (EPI == EPI_ISTFT) {
                // overlap-added block row -> 4 consecutive output samples: COLA normalise, scale, convert
                const float4 a4 = *reinterpret_cast<const float4*>(stg + rl * 36 + 4 * cq);
                const float xv[4] = {a4.x, a4.y, a4.z, a4.w};
                const int s0 = (t2 + g.t0) * g.hop + n0 + 4 * cq - g.shift;
                const long long ob = (long long)b2 * g.out_len;
                if (s0 >= 0 && s0 + 3 < g.out_len && ((s0 | g.out_len) & 3) == 0 && n0 + 4 * cq + 3 < g.N) {
                  const float4 nv = __ldg(reinterpret_cast<const float4*>(g.norm + s0));
                  float y[4];
                  if (g.norm_mul) { y[0] = xv[0] * nv.x; y[1] = xv[1] * nv.y; y[2] = xv[2] * nv.z; y[3] = xv[3] * nv.w; }
                  else { y[0] = xv[0] / nv.x; y[1] = xv[1] / nv.y; y[2] = xv[2] / nv.z; y[3] = xv[3] / nv.w; }
                  if (g.out_dtype == ADN_F32) {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + ob + s0) = make_float4(y[0], y[1], y[2], y[3]);
                  } else if (g.out_dtype == ADN_I16) {
                    short q4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) q4[j] = to_i16(y[j], g.i16_mode);
                    *reinterpret_cast<short4*>(reinterpret_cast<int16_t*>(g.out) + ob + s0) = make_short4(q4[0], q4[1], q4[2], q4[3]);
                  } else {
                    __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
                    __half2* dst = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(g.out) + ob + s0);
                    dst[0] = h01; dst[1] = h23;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const int sj = s0 + j;
                    if (n0 + 4 * cq + j >= g.N || sj < 0 || sj >= g.out_len) continue;
                    const float nv = __ldg(g.norm + sj);
                    const float y = g.norm_mul ? xv[j] * nv : xv[j] / nv;
                    if (g.out_dtype == ADN_F32) reinterpret_cast<float*>(g.out)[ob + sj] = y;
                    else if (g.out_dtype == ADN_I16)
                      reinterpret_cast<int16_t*>(g.out)[ob + sj] = to_i16(y, g.i16_mode);
                    else reinterpret_cast<__half*>(g.out)[ob + sj] = __float2half_rn(y);
                  }
                }
                continue;
              }
              if (g.bias && (!have_bias || g.w_batched)) {
                bias4 = __ldg(reinterpret_cast<const float4*>(g.bias + (long long)b2 * g.bias_bstride + n0) + cq);
                have_bias = true;
              }
              const long long m = (long long)b2 * g.TM + t2;
              const float rs = rsc[hrow][it];
              const float4 a4 = *reinterpret_cast<const float4*>(stg + rl * 36 + 4 * cq);
              float x[4] = {a4.x * rs + bias4.x, a4.y * rs + bias4.y, a4.z * rs + bias4.z, a4.w * rs + bias4.w};
              if (g.act == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = 0.5f * x[j] * (1.0f + fast_erf(x[j] * 0.70710678118654752440f));
              } else if (g.act == ACT_TANH) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = tanhf(x[j]);
              } else if (g.act == ACT_SILU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = x[j] * __fdividef(1.0f, 1.0f + __expf(-x[j]));   // ex2.approx / rcp.approx: ~2 ulp,
                                                                                                     // a third of the epilogue's instructions
              } else if (g.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.f);
              } else if (g.act == ACT_RELU2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float r = fmaxf(x[j], 0.f); x[j] = r * r; }
              } else if (g.act == ACT_PRELU) {
                const float slope = __ldg(g.act_param);
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = x[j] >= 0.f ? x[j] : slope * x[j];
              } else if (g.act == ACT_SWOOSH_L || g.act == ACT_SWOOSH_R) {
                const float off = g.act == ACT_SWOOSH_L ? 4.0f : 1.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  // softplus(y) = max(y, 0) + log(1 + exp(-|y|)) on ex2.approx / lg2.approx: the argument of the log is in (1, 2],
                  // absolute error ~1e-7 on values of order 1 (the accurate log1pf / expf pair was the epilogue's critical path)
                  const float y = x[j] - off;
                  x[j] = fmaxf(y, 0.f) + __logf(1.0f + __expf(-fabsf(y))) - 0.08f * x[j];
                }
              }
              const long long o = m * g.ldc + n0 + 4 * cq;
              {
                const float4 r4 = res4[hrow][it];
                x[0] += r4.x; x[1] += r4.y; x[2] += r4.z; x[3] += r4.w;
              }
              if (g.resid2) {
                const float4 o4 = *reinterpret_cast<const float4*>(g.resid2 + o);
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(g.colscale + n0) + cq);
                x[0] = o4.x + (x[0] - o4.x) * s4.x; x[1] = o4.y + (x[1] - o4.y) * s4.y;
                x[2] = o4.z + (x[2] - o4.z) * s4.z; x[3] = o4.w + (x[3] - o4.w) * s4.w;
              }
              if (g.C) *reinterpret_cast<float4*>(g.C + o) = make_float4(x[0], x[1], x[2], x[3]);
              if (g.Chi && !g.Clo) {                 // bf16 operand plane for a bf16 consumer
                __nv_bfloat162 p01 = __floats2bfloat162_rn(x[0], x[1]), p23 = __floats2bfloat162_rn(x[2], x[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&p01);
                pk.y = *reinterpret_cast<uint32_t*>(&p23);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.Chi) + o) = pk;
              } else if (g.Chi) {
                float h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  h[j] = __uint_as_float(__float_as_uint(x[j]) & 0xFFFFE000u);
                  l[j] = x[j] - h[j];
                }
                *reinterpret_cast<float4*>(g.Chi + o) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(g.Clo + o) = make_float4(l[0], l[1], l[2], l[3]);
              }
            }
            __syncwarp();
          }
          continue;
        }
        if (!row_ok || n0 >= g.N) continue;
        if (EPI == EPI_STORE) {
          float* dst = g.C + (long long)b * g.c_sB + (long long)t * g.c_sT + n0;
          if (c0 + 32 <= BN && n0 + 32 <= g.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < BN && n0 + j < g.N) dst[j] = __uint_as_float(v[j]);
          }
        } else {
          // ISTFT: s = (t + t0)*hop + n - shift ; y = acc (/|*) norm[s]
          const int s0 = (t + g.t0) * g.hop + n0 - g.shift;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int s = s0 + j;
            if (c0 + j >= BN || n0 + j >= g.N || s < 0 || s >= g.out_len) continue;
            float x = __uint_as_float(v[j]);
            const float nv = __ldg(g.norm + s);
            x = g.norm_mul ? x * nv : x / nv;
            const long long o = (long long)b * g.out_len + s;
            if (g.out_dtype == ADN_F32) reinterpret_cast<float*>(g.out)[o] = x;
            else if (g.out_dtype == ADN_I16) {
              reinterpret_cast<int16_t*>(g.out)[o] = to_i16(x, g.i16_mode);
            } else reinterpret_cast<__half*>(g.out)[o] = __float2half_rn(x);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
    }
    if (AF && store_lane) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's bulk stores have completed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

bool make_row_map(CUtensorMap* map, const float* base, int k_extent, int rows, long long row_stride, int batches,
                  long long batch_stride, int box_rows, int box_batches, std::string& err, bool bf16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  const int esz = bf16 ? 2 : 4;                  // bf16: `base` holds 2-byte elements, strides are in elements
  if ((row_stride * esz) % 16 || (batch_stride * esz) % 16 || ((uintptr_t)base) % 16) {
    err = "TMA needs 16-byte aligned base and strides";
    return false;
  }
  cuuint64_t dims[3] = {(cuuint64_t)k_extent, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * esz, (cuuint64_t)batch_stride * esz};
  cuuint32_t box[3] = {(cuuint32_t)(bf16 ? 2 * BK : BK), (cuuint32_t)box_rows, (cuuint32_t)box_batches};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r); return false; }
  return true;
}

bool make_weight_map(CUtensorMap* map, const float* base, int k_pad, int n_pad, int box_n, std::string& err,
                     int batches, bool bf16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  const cuuint64_t esz = bf16 ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)n_pad, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)k_pad * esz, (cuuint64_t)k_pad * esz * (cuuint64_t)n_pad};
  cuuint32_t box[3] = {(cuuint32_t)(bf16 ? 2 * BK : BK), (cuuint32_t)box_n, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(W) failed: " + std::to_string((int)r); return false; }
  return true;
}

bool make_tile_map(CUtensorMap* map, const float* base, int cols, int rows, long long row_stride, int batches,
                   long long batch_stride, int box_cols, int box_rows, std::string& err) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  if ((row_stride * 4) % 16 || (batch_stride * 4) % 16 || ((uintptr_t)base) % 16 || box_cols > 256 || box_rows > 256) {
    err = "TMA tile map: 16-byte aligned base / strides and boxes of at most 256 elements per dimension";
    return false;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)batch_stride * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(tile) failed: " + std::to_string((int)r); return false; }
  return true;
}

bool make_store_map(CUtensorMap* map, float* base, int cols, int rows, long long row_stride, int batches, long long batch_stride,
                    std::string& err) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  if ((row_stride * 4) % 16 || (batch_stride * 4) % 16 || ((uintptr_t)base) % 16) {
    err = "TMA store map needs a 16-byte aligned base and strides";
    return false;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)batch_stride * 4};
  cuuint32_t box[3] = {32, 16, 1};                   // the per-warp staging tile of the fp32-A epilogue
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(C) failed: " + std::to_string((int)r); return false; }
  return true;
}

template <int BN, int EPI, bool BF, bool AF = false, bool DK = false, bool TS = false>
static cudaError_t launch_t(const TcPlan& p, const TcArgs& a, int sms, cudaStream_t st) {
  using S = Smem<BN, BF, AF, TS>;
  static unsigned long long configured = 0;      // per device
  auto kern = gemm_tc_kernel<BN, EPI, BF, AF, DK, TS>;
  if (adn_first_use_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL);
    if (e != cudaSuccess) return e;
  }
  const int n_tiles = a.m_tiles * ((a.N + BN - 1) / BN);
  const int grid = n_tiles < sms ? n_tiles : sms;
  kern<<<grid, TS ? NTHREADS_TS : AF ? NTHREADS_AF : NTHREADS, S::TOTAL, st>>>(p.map_a_hi, p.map_a_lo, p.map_w_hi, p.map_w_lo, p.map_w2_hi, p.map_w2_lo, AF ? p.map_c : p.map_a_hi, a);
  return cudaGetLastError();
}

template <int BN>
static cudaError_t launch_bn(const TcPlan& p, const TcArgs& a, int epi, int sms, cudaStream_t st) {
  if (p.bf16) return epi == EPI_LIN ? launch_t<BN, EPI_LIN, true>(p, a, sms, st) : cudaErrorInvalidValue;
  if (p.a_f32) {
    if (epi != EPI_LIN) return cudaErrorInvalidValue;
    // the TMA-store epilogue writes 32-column boxes: a tile width that is not a multiple of 32 (176) would spill its last box
    // into the next N tile's columns, so it is only usable when ONE tile covers N
    if (BN % 32 && a.N > BN) return cudaErrorInvalidValue;
    if (BN == 64) {
      const char* ss = getenv("ADN_TC_SS");   // diagnostics / A-B test: the shared-memory A operand form (read per launch)
      if (!(ss && atoi(ss) != 0)) return launch_t<64, EPI_LIN, false, true, false, true>(p, a, sms, st);
    }
    if (BN > 64 && a.K >= 512 && a.taps == 0) return launch_t<BN, EPI_LIN, false, true, true>(p, a, sms, st);
    return launch_t<BN, EPI_LIN, false, true>(p, a, sms, st);
  }
  if (epi == EPI_STORE) return launch_t<BN, EPI_STORE, false>(p, a, sms, st);
  if (epi == EPI_ISTFT) return launch_t<BN, EPI_ISTFT, false>(p, a, sms, st);
  if (epi == EPI_LIN) return launch_t<BN, EPI_LIN, false>(p, a, sms, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch(const TcPlan& p, const TcArgs& a, int epi, int sms, cudaStream_t st) {
  if (p.bn == 176) return p.bf16 ? cudaErrorInvalidValue : launch_bn<176>(p, a, epi, sms, st);
  if (p.bn == 256) return launch_bn<256>(p, a, epi, sms, st);
  if (p.bn == 128) return launch_bn<128>(p, a, epi, sms, st);
  if (p.bn == 64) return p.bf16 ? cudaErrorInvalidValue : launch_bn<64>(p, a, epi, sms, st);
  return cudaErrorInvalidValue;
}

// x -> (hi, lo): hi = x with the 13 low mantissa bits cleared (exactly representable in tf32),
// lo = x - hi (exact in fp32).
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = x[i];
    float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

void split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st) {
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, hi, lo, n);
}

}  // namespace tc
