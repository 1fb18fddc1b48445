// tcgen05 dense path: fp32-accurate GEMM on the 5th-gen tensor cores via a
// 3xTF32 split, TMA-fed, accumulators in TMEM, two-level accumulation.
//
//   D[M,N] = A[M,K] * B[N,K]^T      (both operands K-major)
//   a = a_hi + a_lo,  a_hi = rna_tf32(a), a_lo = rna_tf32(a - a_hi)   (same for b)
//   a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi      (dropped a_lo*b_lo ~ 2^-22 |ab|)
//
// Accuracy note (measured, profiles/r1_accuracy_probe.txt): the tensor core's
// fp32 accumulator TRUNCATES on every accumulate (bias -1.1e-8*K relative for
// same-sign data with one long TMEM chain).  So the K loop is cut into chunks of
// 128: hi*hi partial sums accumulate in TMEM for one chunk only, then the
// epilogue warps add the chunk into fp32 REGISTERS with round-to-nearest while
// the MMA warp fills the other TMEM buffer; the small cross terms get their own
// TMEM accumulator (their truncation error is 2^-11 smaller).
//
// Persistent, warp-specialised CTA (one per SM, 192 threads):
//   warp 0  TMA producer : 4 tiles/stage (A_hi, A_lo, B_hi, B_lo; 128 rows x 32 fp32,
//                          SWIZZLE_128B, one contiguous 16 KiB box each) into a
//                          3-stage smem ring, mbarrier expect_tx
//   warp 1  MMA issuer   : 12 x tcgen05.mma.kind::tf32 (M128 N128 K8) per stage;
//                          tcgen05.commit frees the stage / publishes the chunk
//   warps 2-5 epilogue   : tcgen05.ld chunk -> register accumulate; at tile end
//                          bias/ReLU | ReLU-mask | split-K partial, staged through
//                          smem so every global access is a full 128 B line
// Operand planes are produced by the split pre-pass below in k-block-major layout
//   plane[kb][row][32]   (kb = k / 32)
// so each TMA box is one contiguous 16 KiB read whatever the logical row stride
// (the transposed operands of dW = X^T dZ have a 128 KiB row stride otherwise).
//
// Reference arithmetic being replaced: tf.layers.dense / tf.matmul and their gradients,
//   adanet/examples/simple_dnn.py:72-86,103-110.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

#include "dense_simt.cuh"
#include "dense_tc.cuh"

namespace adn {
namespace tc {

static constexpr int BM = 128, BN = 128, BK = 32;
static constexpr int STAGES = 3;
static constexpr int CHUNK_KB = 4;                    // k-blocks per TMEM accumulation chunk (K = 128)
static constexpr int TILE_BYTES = 128 * BK * 4;       // 16 KiB
static constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A_hi A_lo B_hi B_lo
static constexpr int EPI_STAGE_FLOATS = 32 * 33;      // per epilogue warp: output slice, transposed through smem
static constexpr int EPI_MASK_FLOATS = 32 * 36;       // per epilogue warp: ReLU-mask slice landed by cp.async (16 B rows)
static constexpr int EPI_BYTES = 4 * (EPI_STAGE_FLOATS + EPI_MASK_FLOATS) * 4;
static constexpr int BAR_BYTES = 256;
// dynamic smem is declared __align__(1024) (SWIZZLE_128B atoms need it); no slack is left: 232192 of 232448 B
static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
static constexpr int NUM_THREADS = 192;
static constexpr int TMEM_COLS = 512;                 // H0 [0,128) H1 [128,256) S0 [256,384) S1 [384,512)
static constexpr int MAX_SPLITS = 64;

enum { EPI_BIAS_ACT = 0, EPI_MASK = 1, EPI_PARTIAL = 2 };

struct GemmParams {
  float* out;
  int M, N, ldc;
  int tiles_m, tiles_n, splits;
  int total_kb;        // K blocks of 32 over the whole (padded) K
  int kb_per_split;
  const float* bias;   // EPI_BIAS_ACT
  int act;
  const float* mask;   // EPI_MASK (nullable)
  int ldmask;
  int chunk_kb;        // k-blocks per TMEM accumulation chunk (tuning knob, default CHUNK_KB)
  int merge_small;     // 1: cross terms share the chunk accumulator (tuning knob, default 0)
};

// ---------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a broken pipeline traps (CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((it & 1023u) == 1023u) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();   // ~2 s at 2 GHz
    }
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major)
//   [32,46) SBO >> 4 (8 rows * 128 B = 1024 B -> 64) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
// a_major=b_major=K (0), n_dim=N>>3 [17,23), m_dim=M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// descriptors passed as 32-bit halves: only the low word (address field) differs between operands
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t d_hi, uint32_t idesc,
                                          uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(da_lo), "r"(db_lo), "r"(d_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// work item -> (tile_m, tile_n, split).  n fastest so concurrently resident CTAs share A tiles.
struct Item {
  int m0, n0, kb0, nkb, split;
};
__device__ __forceinline__ Item decode_item(const GemmParams& g, int item) {
  Item it;
  const int tiles = g.tiles_m * g.tiles_n;
  it.split = item / tiles;
  const int t = item - it.split * tiles;
  const int tm = t / g.tiles_n;
  it.m0 = tm * BM;
  it.n0 = (t - tm * g.tiles_n) * BN;
  it.kb0 = it.split * g.kb_per_split;
  it.nkb = min(g.total_kb, it.kb0 + g.kb_per_split) - it.kb0;
  return it;
}

// ---------------------------------------------------------------------------------
// GEMM kernel
// ---------------------------------------------------------------------------------
template <int EPI, int CHUNK, int MERGE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const GemmParams g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // SWIZZLE_128B tiles must sit on 1024 B boundaries
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;          // [2]       MMA -> epilogue (chunk ready)
  uint64_t* acc_empty = bars + 2 * STAGES + 2;     // [2]       epilogue -> MMA (chunk drained), count 4
  uint64_t* s_empty = bars + 2 * STAGES + 4;       // [2]       epilogue -> MMA (small-term acc read), count 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_items = g.tiles_m * g.tiles_n * g.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi);
    tma_prefetch_desc(&map_b_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(smem_u32(&acc_full[b]), 1);
        mbar_init(smem_u32(&acc_empty[b]), 4);
        mbar_init(smem_u32(&s_empty[b]), 4);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;   // ring slot / phase, advanced incrementally (no div/mod in the loop)
      const uint32_t smem0 = smem_u32(smem);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(g, item);
        for (int kb = 0; kb < it.nkb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, STAGE_BYTES);
          const uint32_t base = smem0 + s * STAGE_BYTES;
          const int kc = it.kb0 + kb;
          tma_load_3d(&map_a_hi, fb, base + 0 * TILE_BYTES, 0, it.m0, kc);
          tma_load_3d(&map_a_lo, fb, base + 1 * TILE_BYTES, 0, it.m0, kc);
          tma_load_3d(&map_b_hi, fb, base + 2 * TILE_BYTES, 0, it.n0, kc);
          tma_load_3d(&map_b_lo, fb, base + 3 * TILE_BYTES, 0, it.n0, kc);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // One thread feeds the tensor core: a 128x128x8 TF32 MMA retires every 64 clk, so the issue loop
    // must stay well under 64 clk per MMA -- ring counters are incremental, descriptors are built from
    // 32-bit halves (only the 14-bit address field of the low word moves), no div/mod anywhere.
    // The whole warp runs the loop converged (all lanes wait on the barriers); only the issue itself is
    // under elect.sync, so every operand is warp-uniform and lives in uniform registers.
    {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      const uint32_t desc_hi = (uint32_t)(make_desc(0) >> 32);
      const uint32_t desc_lo0 = (uint32_t)make_desc(smem_u32(smem));   // stage 0, tile 0 (A_hi)
      uint32_t s = 0, ph = 0, gchunk = 0, tile_i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tile_i) {
        const Item it = decode_item(g, item);
        const uint32_t acc_s = tmem_base + 256 + (tile_i & 1) * 128;
        mbar_wait(smem_u32(&s_empty[tile_i & 1]), ((tile_i >> 1) & 1) ^ 1);   // small-term accumulator free
        tc_fence_after();
        uint32_t s_accum = 0;                      // first small-term MMA of the tile overwrites
        for (int kb = 0; kb < it.nkb; kb += CHUNK, ++gchunk) {
          const uint32_t b = gchunk & 1;
          mbar_wait(smem_u32(&acc_empty[b]), ((gchunk >> 1) & 1) ^ 1);      // chunk buffer drained
          tc_fence_after();
          const uint32_t acc_h = tmem_base + b * 128;
          const int nk = min(CHUNK, it.nkb - kb);
          uint32_t h_accum = 0;                    // first hi*hi MMA of the chunk overwrites
          for (int kk = 0; kk < nk; ++kk) {
            mbar_wait(smem_u32(&full_bar[s]), ph);
            tc_fence_after();
            const uint32_t lo = desc_lo0 + s * (STAGE_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 8; ++k) {   // UMMA_K = 8 tf32 = 32 B -> +2 in the (>>4) address field
                const uint32_t a_hi = lo + 0 * (TILE_BYTES >> 4) + 2 * k, a_lo = lo + 1 * (TILE_BYTES >> 4) + 2 * k;
                const uint32_t b_hi = lo + 2 * (TILE_BYTES >> 4) + 2 * k, b_lo = lo + 3 * (TILE_BYTES >> 4) + 2 * k;
                if (MERGE) {
                  umma_tf32(acc_h, a_hi, b_hi, desc_hi, idesc, (k == 0) ? h_accum : 1u);
                  umma_tf32(acc_h, a_lo, b_hi, desc_hi, idesc, 1u);
                  umma_tf32(acc_h, a_hi, b_lo, desc_hi, idesc, 1u);
                } else {
                  umma_tf32(acc_s, a_lo, b_hi, desc_hi, idesc, (k == 0) ? s_accum : 1u);
                  umma_tf32(acc_s, a_hi, b_lo, desc_hi, idesc, 1u);
                  umma_tf32(acc_h, a_hi, b_hi, desc_hi, idesc, (k == 0) ? h_accum : 1u);
                }
              }
              umma_commit(smem_u32(&empty_bar[s]));  // frees this smem stage when the MMAs retire
              if (kk == nk - 1) umma_commit(smem_u32(&acc_full[b]));   // chunk (and small terms) complete
            }
            __syncwarp();
            s_accum = 1u;
            h_accum = 1u;
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ================= epilogue warps 2..5: TMEM lane quadrant = warp % 4 =================
    const int quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    float* stage = epi_stage + quad * (EPI_STAGE_FLOATS + EPI_MASK_FLOATS);
    float* mstage = stage + EPI_STAGE_FLOATS;
    uint32_t gchunk = 0, tile_i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tile_i) {
      const Item it = decode_item(g, item);
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      const int nchunks = (it.nkb + CHUNK - 1) / CHUNK;
      for (int c = 0; c < nchunks; ++c, ++gchunk) {
        const uint32_t b = gchunk & 1;
        mbar_wait(smem_u32(&acc_full[b]), (gchunk >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_base + b * 128 + q * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(r[j]);   // fp32 RN adds
        }
        if (c == nchunks - 1 && !MERGE) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + 256 + (tile_i & 1) * 128 + q * 32, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&acc_empty[b]));
          if (c == nchunks - 1) mbar_arrive(smem_u32(&s_empty[tile_i & 1]));
        }
      }
      // ---- tile output: per-warp smem transpose so each global access is one full 128 B row segment ----
      float* out = g.out;
      if (EPI == EPI_PARTIAL) out += (size_t)it.split * g.M * g.N;
      const int mrow0 = it.m0 + quad * 32;
      const int rmax = min(32, g.M - mrow0);
      // ReLU-mask slices land in smem through cp.async (no registers: the 128 accumulators stay live)
      const bool mask_async = (EPI == EPI_MASK) && g.mask != nullptr && ((g.ldmask & 3) == 0) &&
                              ((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cbase = it.n0 + q * 32;
        const int col = cbase + lane;
        if (EPI == EPI_MASK && mask_async) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int id = lane + 32 * i;           // 16-byte chunk id: 8 chunks per 128 B row
            const int r = id >> 3, c4 = (id & 7) * 4;
            int bytes = (g.N - (cbase + c4)) * 4;
            bytes = (r < rmax) ? max(0, min(16, bytes)) : 0;
            const float* src = bytes > 0 ? g.mask + (size_t)(mrow0 + r) * g.ldmask + cbase + c4 : g.mask;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(mstage + r * 36 + c4)), "l"(src),
                         "r"(bytes)
                         : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = acc[q * 32 + j];
        if (EPI == EPI_MASK && mask_async) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        float bias_v = 0.f;
        if (EPI == EPI_BIAS_ACT && g.bias && col < g.N) bias_v = __ldg(g.bias + col);
        if (col < g.N) {
          float* op = out + (size_t)mrow0 * g.ldc + col;
          if (rmax == 32) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              float v = stage[r * 33 + lane];
              if (EPI == EPI_BIAS_ACT) {
                v += bias_v;
                if (g.act == ADN_ACT_RELU) v = fmaxf(v, 0.f);
              } else if (EPI == EPI_MASK) {
                if (mask_async) {
                  if (!(mstage[r * 36 + lane] > 0.f)) v = 0.f;
                } else if (g.mask && !(__ldg(g.mask + (size_t)(mrow0 + r) * g.ldmask + col) > 0.f)) {
                  v = 0.f;
                }
              }
              op[(size_t)r * g.ldc] = v;
            }
          } else {
            for (int r = 0; r < rmax; ++r) {
              float v = stage[r * 33 + lane];
              if (EPI == EPI_BIAS_ACT) {
                v += bias_v;
                if (g.act == ADN_ACT_RELU) v = fmaxf(v, 0.f);
              } else if (EPI == EPI_MASK) {
                if (mask_async) {
                  if (!(mstage[r * 36 + lane] > 0.f)) v = 0.f;
                } else if (g.mask && !(__ldg(g.mask + (size_t)(mrow0 + r) * g.ldmask + col) > 0.f)) {
                  v = 0.f;
                }
              }
              op[(size_t)r * g.ldc] = v;
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------
// hi/lo split pre-pass into k-block-major planes  plane[kb][row][32]
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
  lo = __uint_as_float(l);
}

// src[rows, cols] row-major (K = cols) -> hi/lo[nkb][rows][32], zero padded in K
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, int rows, int cols,
             int nkb) {
  const int vec_per_row = nkb * 8;                       // float4 per padded row
  const int64_t nvec = (int64_t)rows * vec_per_row;
  const bool vec_src = ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec_per_row);
    const int c = (int)(i % vec_per_row) * 4;
    float v[4];
    if (vec_src && c + 3 < cols) {
      float4 t = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * cols + c));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = (c + q < cols) ? __ldg(src + (size_t)r * cols + c + q) : 0.f;
    }
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_tf32(v[q], h[q], l[q]);
    const size_t dst = ((size_t)(c >> 5) * rows + r) * 32 + (c & 31);
    *reinterpret_cast<float4*>(hi + dst) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + dst) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// src[rows, cols] row-major, K = rows -> hiT/loT[nkb][cols][32] (plane row = source column), zero padded in K
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ src, float* __restrict__ hiT, float* __restrict__ loT, int rows,
                       int cols) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;   // blockIdx.x = k-block
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? __ldg(src + (size_t)r * cols + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;     // plane row
    if (c < cols) {
      float h, l;
      split_tf32(tile[tx][i], h, l);
      const size_t dst = ((size_t)blockIdx.x * cols + c) * 32 + tx;
      hiT[dst] = h;
      loT[dst] = l;
    }
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled g_encode = nullptr;

int init() {
  static std::once_flag once;
  static int rc = ADN_OK;
  std::call_once(once, []() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) {
      (void)cudaGetLastError();
      rc = fail(ADN_ERR_CUDA, "tc::init: cuTensorMapEncodeTiled entry point unavailable");
      return;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    bool ok = true;
#define ADN_TC_ATTR(E, C, MG) \
  ok = ok && (cudaFuncSetAttribute(tc_gemm_kernel<E, C, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess)
#define ADN_TC_ATTR3(C, MG) ADN_TC_ATTR(EPI_BIAS_ACT, C, MG); ADN_TC_ATTR(EPI_MASK, C, MG); ADN_TC_ATTR(EPI_PARTIAL, C, MG)
    ADN_TC_ATTR3(4, 0); ADN_TC_ATTR3(8, 0); ADN_TC_ATTR3(1024, 0);
    ADN_TC_ATTR3(4, 1); ADN_TC_ATTR3(8, 1); ADN_TC_ATTR3(1024, 1);
#undef ADN_TC_ATTR3
#undef ADN_TC_ATTR
    if (!ok) {
      (void)cudaGetLastError();
      rc = fail(ADN_ERR_CUDA, "tc::init: cudaFuncSetAttribute(smem=%d) failed", SMEM_BYTES);
    }
  });
  return rc;
}

static inline int64_t nkb_of(int64_t k) { return ceil_div(k, BK); }

// thresholds: the tensor path pays a split pre-pass; skinny layers stay on CUDA cores
bool fwd_supported(int64_t batch, int64_t in, int64_t out) { return batch >= 128 && in >= 32 && out >= 64; }
bool bwd_supported(int64_t batch, int64_t in, int64_t out) { return batch >= 128 && in >= 32 && out >= 64; }

struct Planes {
  float* hi;
  float* lo;
  int64_t rows, nkb;
};

static int make_map(CUtensorMap* map, const float* plane, int64_t rows, int64_t nkb) {
  if (!g_encode) return fail(ADN_ERR_CUDA, "tc: adn_init() was not called");
  cuuint64_t gdim[3] = {(cuuint64_t)BK, (cuuint64_t)rows, (cuuint64_t)nkb};
  cuuint64_t gstride[2] = {(cuuint64_t)BK * sizeof(float), (cuuint64_t)rows * BK * sizeof(float)};
  cuuint32_t box[3] = {(cuuint32_t)BK, 128u, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(plane), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ADN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld nkb=%lld", (int)r,
                                     (long long)rows, (long long)nkb);
  return ADN_OK;
}

static int do_split(const float* src, int64_t rows, int64_t cols, Planes& p, cudaStream_t st) {
  const int64_t nvec = rows * p.nkb * 8;
  int blocks = (int)std::min<int64_t>(ceil_div(nvec, 256), (int64_t)sm_count() * 16);
  split_kernel<<<blocks, 256, 0, st>>>(src, p.hi, p.lo, (int)rows, (int)cols, (int)p.nkb);
  ADN_CHECK_LAUNCH("tc split");
  return ADN_OK;
}

static int do_split_T(const float* src, int64_t rows, int64_t cols, Planes& p, cudaStream_t st) {
  dim3 grid((unsigned)p.nkb, (unsigned)ceil_div(cols, 32));
  split_transpose_kernel<<<grid, 256, 0, st>>>(src, p.hi, p.lo, (int)rows, (int)cols);
  ADN_CHECK_LAUNCH("tc split_transpose");
  return ADN_OK;
}

template <int EPI>
static int launch_gemm(const Planes& a, const Planes& b, GemmParams g, cudaStream_t st, const char* what) {
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a.hi, a.rows, a.nkb))) return rc;
  if ((rc = make_map(&ma_lo, a.lo, a.rows, a.nkb))) return rc;
  if ((rc = make_map(&mb_hi, b.hi, b.rows, b.nkb))) return rc;
  if ((rc = make_map(&mb_lo, b.lo, b.rows, b.nkb))) return rc;
  // tuning knobs (experiments only; defaults are the shipped configuration)
  static const int env_chunk = getenv("ADN_TC_CHUNK") ? atoi(getenv("ADN_TC_CHUNK")) : CHUNK_KB;
  static const int env_merge = getenv("ADN_TC_MERGE") ? atoi(getenv("ADN_TC_MERGE")) : 0;
  g.tiles_m = (int)ceil_div(g.M, BM);
  g.tiles_n = (int)ceil_div(g.N, BN);
  const int items = g.tiles_m * g.tiles_n * g.splits;
  const int grid = std::min(items, sm_count());
#define ADN_TC_LAUNCH(C, MG) tc_gemm_kernel<EPI, C, MG><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, g)
  if (env_merge) {
    if (env_chunk >= 1024) ADN_TC_LAUNCH(1024, 1);
    else if (env_chunk >= 8) ADN_TC_LAUNCH(8, 1);
    else ADN_TC_LAUNCH(4, 1);
  } else {
    if (env_chunk >= 1024) ADN_TC_LAUNCH(1024, 0);
    else if (env_chunk >= 8) ADN_TC_LAUNCH(8, 0);
    else ADN_TC_LAUNCH(4, 0);
  }
#undef ADN_TC_LAUNCH
  ADN_CHECK_LAUNCH(what);
  return ADN_OK;
}

// carve 256B-aligned plane pairs out of the caller's workspace
struct Carver {
  char* p;
  char* end;
  bool ok = true;
  Planes planes(int64_t rows, int64_t nkb) {
    Planes pl{nullptr, nullptr, rows, nkb};
    const int64_t bytes = align_up(rows * nkb * BK * (int64_t)sizeof(float), 256);
    if (p + 2 * bytes > end) { ok = false; return pl; }
    pl.hi = reinterpret_cast<float*>(p);
    pl.lo = reinterpret_cast<float*>(p + bytes);
    p += 2 * bytes;
    return pl;
  }
  float* floats(int64_t n) {
    const int64_t bytes = align_up(n * (int64_t)sizeof(float), 256);
    if (p + bytes > end) { ok = false; return nullptr; }
    float* r = reinterpret_cast<float*>(p);
    p += bytes;
    return r;
  }
};

static inline int64_t plane_pair_bytes(int64_t rows, int64_t nkb) {
  return 2 * align_up(rows * nkb * BK * (int64_t)sizeof(float), 256);
}

int64_t dense_fwd_workspace_bytes(int64_t batch, int64_t in, int64_t out) {
  if (!fwd_supported(batch, in, out)) return 0;
  return plane_pair_bytes(batch, nkb_of(in)) + plane_pair_bytes(out, nkb_of(in)) + 1024;
}

// dW split-K: the partial buffer bounds the split count (<= 16M floats, <= 64 splits)
static int max_dw_splits(int64_t in, int64_t out) {
  int64_t s = (16LL << 20) / std::max<int64_t>(1, in * out);
  if (s > MAX_SPLITS) s = MAX_SPLITS;
  if (s < 1) s = 1;
  return (int)s;
}

// pick S minimising (persistent rounds) x (k-blocks per item) + a per-split reduction cost
static int dw_splits(int64_t tiles, int64_t kblocks, int max_s) {
  const int sms = sm_count();
  int best = 1;
  double best_t = 1e30;
  for (int s = 1; s <= max_s && s <= kblocks; ++s) {
    const int64_t kps = ceil_div(kblocks, s);
    const int64_t s_eff = ceil_div(kblocks, kps);
    const int64_t rounds = ceil_div(tiles * s_eff, sms);
    const double t = (double)rounds * ((double)kps + 6.0) + 0.75 * (double)s_eff;   // 6: tile prologue/epilogue in k-block units
    if (t < best_t) { best_t = t; best = (int)s_eff; }
  }
  return best;
}

int64_t dense_bwd_workspace_bytes(int64_t batch, int64_t in, int64_t out) {
  if (!bwd_supported(batch, in, out)) return 0;
  const int64_t kb_b = nkb_of(batch), kb_o = nkb_of(out);
  int64_t b = plane_pair_bytes(batch, kb_o)   // dz       [B, out]   (A of dX)
              + plane_pair_bytes(in, kb_o)    // w        [in, out]  (B of dX)
              + plane_pair_bytes(in, kb_b)    // x^T      [in, B]    (A of dW)
              + plane_pair_bytes(out, kb_b);  // dz^T     [out, B]   (B of dW)
  b += align_up((int64_t)max_dw_splits(in, out) * in * out * (int64_t)sizeof(float), 256);   // dW split-K partials
  b += align_up(64 * out * (int64_t)sizeof(float), 256);                   // db partials
  return b + 1024;
}

int dense_fwd(const float* x, const float* w, const float* b, float* y, int64_t batch, int64_t in, int64_t out,
              int act, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < dense_fwd_workspace_bytes(batch, in, out))
    return fail(ADN_ERR_WORKSPACE, "tc dense_fwd: workspace %lld < %lld bytes", (long long)ws_bytes,
                (long long)dense_fwd_workspace_bytes(batch, in, out));
  Carver c{reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255),
           reinterpret_cast<char*>(ws) + ws_bytes};
  const int64_t nkb = nkb_of(in);
  Planes px = c.planes(batch, nkb);   // A = x        [M=B,   K=in]
  Planes pw = c.planes(out, nkb);     // B = w^T      [N=out, K=in]
  if (!c.ok) return fail(ADN_ERR_WORKSPACE, "tc dense_fwd: workspace carve failed");
  int rc;
  if ((rc = do_split(x, batch, in, px, st))) return rc;
  if ((rc = do_split_T(w, in, out, pw, st))) return rc;
  GemmParams g{};
  g.out = y; g.M = (int)batch; g.N = (int)out; g.ldc = (int)out;
  g.total_kb = (int)nkb; g.kb_per_split = (int)nkb; g.splits = 1;
  g.bias = b; g.act = act;
  return launch_gemm<EPI_BIAS_ACT>(px, pw, g, st, "tc dense_fwd gemm");
}

int dense_bwd(const float* x, const float* w, const float* dz, float* dx, float* dw, float* db, int64_t batch,
              int64_t in, int64_t out, int x_relu_mask, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < dense_bwd_workspace_bytes(batch, in, out))
    return fail(ADN_ERR_WORKSPACE, "tc dense_bwd: workspace %lld < %lld bytes", (long long)ws_bytes,
                (long long)dense_bwd_workspace_bytes(batch, in, out));
  Carver c{reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255),
           reinterpret_cast<char*>(ws) + ws_bytes};
  const int64_t kb_o = nkb_of(out), kb_b = nkb_of(batch);
  const int max_s = max_dw_splits(in, out);
  Planes pdz = c.planes(batch, kb_o);
  Planes pw = c.planes(in, kb_o);
  Planes pxT = c.planes(in, kb_b);
  Planes pdzT = c.planes(out, kb_b);
  float* part = c.floats((int64_t)max_s * in * out);
  float* dbpart = c.floats(64 * out);
  if (!c.ok) return fail(ADN_ERR_WORKSPACE, "tc dense_bwd: workspace carve failed");
  int rc;
  // ---- dW[in,out] = x^T[in,B] * (dz^T[out,B])^T, split-K over the batch ----
  if ((rc = do_split_T(x, batch, in, pxT, st))) return rc;
  if ((rc = do_split_T(dz, batch, out, pdzT, st))) return rc;
  {
    const int S = dw_splits(ceil_div(in, BM) * ceil_div(out, BN), kb_b, max_s);
    GemmParams g{};
    g.M = (int)in; g.N = (int)out; g.ldc = (int)out;
    g.total_kb = (int)kb_b; g.kb_per_split = (int)ceil_div(kb_b, S);
    g.splits = (int)ceil_div(kb_b, g.kb_per_split);
    g.out = (g.splits == 1) ? dw : part;
    if ((rc = launch_gemm<EPI_PARTIAL>(pxT, pdzT, g, st, "tc dW gemm"))) return rc;
    if (g.splits > 1 && (rc = simt::reduce_partials(part, dw, in * out, g.splits, in * out, st))) return rc;
  }
  if (db && (rc = simt::colsum(dz, db, batch, out, dbpart, st))) return rc;
  // ---- dX[B,in] = dz[B,out] * (w[in,out])^T, ReLU mask from x ----
  if (dx) {
    if ((rc = do_split(dz, batch, out, pdz, st))) return rc;
    if ((rc = do_split(w, in, out, pw, st))) return rc;
    GemmParams g{};
    g.out = dx; g.M = (int)batch; g.N = (int)in; g.ldc = (int)in;
    g.total_kb = (int)kb_o; g.kb_per_split = (int)kb_o; g.splits = 1;
    g.mask = x_relu_mask ? x : nullptr; g.ldmask = (int)in;
    if ((rc = launch_gemm<EPI_MASK>(pdz, pw, g, st, "tc dX gemm"))) return rc;
  }
  return ADN_OK;
}

}  // namespace tc
}  // namespace adn
