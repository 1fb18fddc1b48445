// Plane-native tcgen05 dense pipeline: fp32-accurate GEMMs (3xTF32) whose operands
// AND results live in HBM as pre-split TF32 planes, so no conversion pass sits
// between the layers of a subnetwork.
//
// Split planes of a matrix T[rows, cols] (fp32):
//     hi = rna_tf32(T)          lo = rna_tf32(T - hi)
//   each stored k-block-major   plane[cols/32][rows][32]   (cols zero padded to 32),
//   plus sign bits  bits[cols/32][rows]  (uint32, bit j = T[row, kb*32+j] > 0): the ReLU mask of
//   the backward pass costs 1/32 of a plane and one coalesced word per row instead of a 4 B/element read;
//   one buffer: hi plane, lo plane, sign bits (adn_query(ADN_Q_PLANES_BYTES)).
// One layout serves every GEMM of training because tcgen05 takes either operand
// K-major or MN-major straight from shared memory:
//     K  = cols of T : box {32, 128 rows, 1 kb}   -> K-major  [128 rows][32 k]      SWIZZLE_128B
//     K  = rows of T : box {32, 32 rows, 4 kb}    -> MN-major [4][32 k][32 mn]     SWIZZLE_128B_ATOM_32B
//   both boxes are 16 KiB and contiguous per 16 KiB / 4 KiB run.
//     fwd  Y = X W       A = Xp  K-major (K=in)    B = Wp  MN-major (N=out, K=in)
//     dX   = dZ W^T      A = dZp K-major (K=out)   B = Wp  K-major  (N=in,  K=out)
//     dW   = X^T dZ      A = Xp  MN-major (M=in)   B = dZp MN-major (N=out), K = batch
//   -> no transposed copies, and the epilogue of one GEMM writes the planes the
//   next one reads (bias+ReLU for fwd, ReLU mask + column sums for dX).
//
// Arithmetic: a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (fp32 accumulate in TMEM).
// The tensor core truncates its accumulator on every add, so hi*hi partial sums
// stay in TMEM for 128 K only and are then added in registers with RN
// (profiles/r1a_accuracy_probe_*.txt); cross terms use their own accumulator.
//
// Kernel: persistent, one CTA per SM, 576 threads, warp-specialised, grouped (up to 8 GEMMs per launch)
//   warp 0    TMA producer (3-stage ring, 64 KiB per stage: A_hi A_lo B_hi B_lo)
//   warp 1    MMA issuer (elected lane; per K=8 step one N=256 MMA a_hi x [b_hi|b_lo] + one N=128 MMA a_lo x b_hi)
//   warps 2-17 epilogue (TMEM lane quadrant = warp % 4, column group = (warp-2)/4, 32 accumulators each):
//             tcgen05.ld -> registers (bias/ReLU + sign bits | sign-bit mask) -> per-warp swizzled smem
//             transpose -> column sums -> hi/lo split -> stores that cover whole 32 B sectors of 8 rows
//
// Reference arithmetic replaced: tf.layers.dense and its gradients,
//   adanet/examples/simple_dnn.py:72-86,103-110.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "dense_simt.cuh"
#include "planes.cuh"

namespace adn {
namespace pl {

static constexpr int BM = 128, BN = 128, BK = 32;
static constexpr int STAGES = 3;
static constexpr int CHUNK = 4;                       // k-blocks per TMEM accumulation chunk (K = 128)
static constexpr int TILE_BYTES = 128 * BK * 4;       // 16 KiB
static constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A_hi A_lo B_hi B_lo
static constexpr int EPI_WARPS = 8;
static constexpr int EPI_STAGE_FLOATS = 32 * 32;      // per epilogue warp: 32x32 slice transposed through smem (swizzled)
static constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_FLOATS * 4;
static constexpr int BAR_BYTES = 256;
static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
static constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;   // CTA-pair kernel: 8 epilogue warps x 128 columns
// single-CTA kernel: 16 epilogue warps x 32 columns (TMEM lane quadrant = warp % 4, column group = (warp-2)/4):
// the short-K layer waves are bound by epilogue latency per warp, so thread-level parallelism is what helps;
// each warp stages 32x16 floats (2 KB) at a time to stay inside the 32 KB left beside the 3-stage ring.
static constexpr int EPI_WARPS1 = 16;
static constexpr int NUM_THREADS1 = 64 + 32 * EPI_WARPS1;
static constexpr int EPI_STAGE_FLOATS1 = 32 * 16;
static constexpr int TMEM_COLS = 512;                 // two chunk buffers of [H: 128 | S: 128] columns (pair kernel: H 256 | S 256)
static constexpr int MAX_SPLITS = 64;

enum { EPI_BIAS_ACT = 0, EPI_MASK = 1, EPI_PARTIAL = 2 };

struct GemmParams {
  int M, N;
  int tiles_m, tiles_n, splits;
  int total_kb;          // k-blocks of 32 over the whole K
  int kb_per_split;
  int a_mn, b_mn;        // operand majorness (0 = K-major box, 1 = MN-major box)
  int out_planes;        // 1: result written as split planes (out / out_lo / out_bits), 0: dense fp32
  int m_fastest;         // work-item order: 1 = row blocks fastest (write locality), 0 = column blocks fastest (A reuse in L2)
  // output: dense row-major (ldc) / split-K partial [split][M][N], or planes
  float* out;            // dense base | hi plane base
  float* out_lo;         // lo plane base (OUT_PLANES)
  int ldc;
  int out_nkb;           // planes: k-blocks of the output tensor (ceil(N/32))
  const float* bias;     // EPI_BIAS_ACT (nullable)
  int act;
  uint32_t* out_bits;    // planes + EPI_BIAS_ACT: sign bits of the output
  const uint32_t* mask_bits;  // EPI_MASK (nullable): sign bits of a [M, N] tensor; out = bit ? out : 0
  float* colsum_part;    // EPI_MASK (nullable): [ceil(M/32)][colsum_ld] per-32-row column sums of out
  int colsum_ld;
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

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), 32-bit elements:
//   [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout type
//   K-major  tile [128 rows][32 k], SWIZZLE_128B (type 2): SBO = 1024 (8 rows x 128 B), LBO unused (=1);
//            next K=8 step: +32 B
//   MN-major tile [4 mn-blocks][32 k][32 mn]: 32-bit operands must use the 32 B-granular 128 B swizzle
//            SWIZZLE_128B_BASE32B (type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; cute Layout_MN_SW128_32B_Atom:
//            atom = 32 mn x 4 k rows): LBO = 4096 (next 32-wide mn block), SBO = 512 (next 4 k rows);
//            next K=8 step: +1024 B
__device__ __forceinline__ uint32_t desc_hi_word(int mn_major) {
  return mn_major ? ((uint32_t)(512 >> 4) | (1u << 14) | (1u << 29)) : ((uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29));
}
__device__ __forceinline__ uint32_t desc_lo_word(uint32_t smem_addr, int mn_major) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((mn_major ? (4096u >> 4) : 1u) << 16);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
// a_major [15], b_major [16] (0 = K, 1 = MN), n_dim=N>>3 [17,23), m_dim=M>>4 [24,29).
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t da_hi, uint32_t db_hi,
                                          uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(da_lo), "r"(db_lo), "r"(da_hi), "r"(db_hi), "r"(idesc), "r"(accum)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// hi = rna_tf32(v), lo = rna_tf32(v - hi).  cvt.rna.tf32.f32 is emulated in SASS (add, NaN/Inf test, select,
// mask: 4 instructions); on the bit pattern it is "add half a TF32 ulp to the magnitude, clear the low 13
// bits", done here in two integer ops (Inf stays Inf, NaN stays NaN, finite values are bit-identical).
__device__ __forceinline__ float rna_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = rna_tf32(v);
  lo = rna_tf32(v - hi);
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
  if (g.m_fastest) {            // consecutive items = consecutive row blocks of the same column block
    const int tn = t / g.tiles_m;
    it.n0 = tn * BN;
    it.m0 = (t - tn * g.tiles_m) * BM;
  } else {
    const int tm = t / g.tiles_n;
    it.m0 = tm * BM;
    it.n0 = (t - tm * g.tiles_n) * BN;
  }
  it.kb0 = it.split * g.kb_per_split;
  it.nkb = min(g.total_kb, it.kb0 + g.kb_per_split) - it.kb0;
  return it;
}

// One 32-row x 32-column slice of a tile: lane = row holds its 32 accumulators a[0..31] (columns
// cbase..cbase+31).  Applies bias/ReLU (+ sign bits) or the sign-bit ReLU mask in the register layout,
// transposes through `stage` (16 B chunks XOR-swizzled by row: conflict-free both ways) and writes with
// lane = 4-column group of 4 rows, so every global access covers whole 128 B lines (planes: one contiguous
// 512 B run per instruction).  Shared by the 1-CTA and the CTA-pair kernels.
template <int EPI, int SW>     // SW = staged columns per pass: 32 (4 KB per warp) or 16 (2 KB per warp, two passes)
__device__ __forceinline__ void emit_slice(const GemmParams& g, const bool OUT_PLANES, float* a, uint32_t mwq, float* stage,
                                           int lane, int mrow0, int cbase, int rows_ok, float* dense, bool dense_vec) {
  constexpr int CH = SW / 4;               // 16 B chunks per staged row
  constexpr int RPI = 32 / CH;             // rows covered by one transposed instruction (4 or 8)
  const int cc = lane % CH;                // 16 B chunk (4 columns) this lane owns after the transpose
  const int rsub = lane / CH;
  const int my_row = mrow0 + lane;
  const int kbo = cbase >> 5;
  const bool live = (OUT_PLANES ? (kbo < g.out_nkb) : (cbase < g.N)) && rows_ok > 0;   // warp-uniform
  // valid columns of this slice as a bit mask (warp-uniform); all ones for interior tiles
  const uint32_t cmask = (cbase + 32 <= g.N) ? 0xffffffffu : ((cbase < g.N) ? ((1u << (g.N - cbase)) - 1u) : 0u);
  if (EPI == EPI_BIAS_ACT) {
    if (g.bias) {
      if (cmask == 0xffffffffu && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + cbase) + jj);   // warp-uniform address
          a[4 * jj + 0] += bv.x; a[4 * jj + 1] += bv.y;
          a[4 * jj + 2] += bv.z; a[4 * jj + 3] += bv.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((cmask >> j) & 1u) a[j] += __ldg(g.bias + cbase + j);
      }
    }
    if (g.act == ADN_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = fmaxf(a[j], 0.f);
    }
    if (OUT_PLANES) {
      uint32_t bits = 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) bits |= (a[j] > 0.f) ? (1u << j) : 0u;
      bits &= cmask;
      if (live && my_row < g.M) g.out_bits[(size_t)kbo * g.M + my_row] = bits;
    }
    if (cmask != 0xffffffffu) {       // K padding of the next GEMM must be exact zeros
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (!((cmask >> j) & 1u)) a[j] = 0.f;
    }
  } else if (EPI == EPI_MASK) {
    const uint32_t keep = mwq & cmask;
    if (keep != 0xffffffffu) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (!((keep >> j) & 1u)) a[j] = 0.f;
    }
  }
  // chunk swizzle by row: both the row-wise float4 writes and the transposed float4 reads are bank-conflict free
#define ADN_SWZ(r) (SW == 32 ? ((r) & 7) : (((r) >> 1) & 3))
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 32 / SW; ++h) {
    {
      float4* srow = reinterpret_cast<float4*>(stage + lane * SW);
#pragma unroll
      for (int jj = 0; jj < CH; ++jj)
        srow[jj ^ ADN_SWZ(lane)] = make_float4(a[h * SW + 4 * jj], a[h * SW + 4 * jj + 1], a[h * SW + 4 * jj + 2],
                                               a[h * SW + 4 * jj + 3]);
    }
    __syncwarp();
    if (live) {
      const int c4 = h * SW + cc * 4;       // column offset inside the 32-column slice
      const int col = cbase + c4;
      if (SW == 16) { cs[0] = cs[1] = cs[2] = cs[3] = 0.f; }
      float* hp = OUT_PLANES ? g.out + ((size_t)kbo * g.M + mrow0 + rsub) * 32 + c4 : nullptr;
      float* lp = OUT_PLANES ? g.out_lo + ((size_t)kbo * g.M + mrow0 + rsub) * 32 + c4 : nullptr;
#pragma unroll
      for (int i = 0; i < 32 / RPI; ++i) {
        const int r = rsub + RPI * i;
        const bool rv = (rows_ok == 32) || (r < rows_ok);
        float4 t = reinterpret_cast<const float4*>(stage + r * SW)[cc ^ ADN_SWZ(r)];
        float v[4] = {t.x, t.y, t.z, t.w};
        if (EPI == EPI_MASK) {
#pragma unroll
          for (int k = 0; k < 4; ++k) cs[k] += rv ? v[k] : 0.f;
        }
        if (rv) {
          if (OUT_PLANES) {
            float hh[4], ll[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_tf32(v[k], hh[k], ll[k]);
            *reinterpret_cast<float4*>(hp + i * RPI * 32) = make_float4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<float4*>(lp + i * RPI * 32) = make_float4(ll[0], ll[1], ll[2], ll[3]);
          } else {
            float* op = dense + (size_t)(mrow0 + r) * g.ldc + col;
            if (dense_vec && col + 3 < g.N) {
              *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (col + k < g.N) op[k] = v[k];
            }
          }
        }
      }
      if (EPI == EPI_MASK && g.colsum_part) {
        // rows of this lane: rsub + RPI*i; fold the lanes that own the same columns in a fixed order
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (SW == 16) cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4);
          cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
          cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
        }
        if (rsub == 0) {
          float* cp = g.colsum_part + (size_t)(mrow0 >> 5) * g.colsum_ld + col;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (col + k < g.colsum_ld) cp[k] = cs[k];
        }
      }
    }
    __syncwarp();
  }
#undef ADN_SWZ
}

// ---------------------------------------------------------------------------------
// GEMM kernel (grouped): one persistent launch runs the tiles of up to MAX_GROUP independent GEMMs of
// the same epilogue kind -- the same layer of every candidate subnetwork of an AdaNet iteration -- so the
// launch, prologue and pipeline fill/drain are paid once per layer wave instead of once per candidate,
// and the short-K tiles of narrow candidates hide behind the long-K tiles of wide ones.
// Work items are numbered problem after problem; CTA c takes items c, c + grid, ... (every CTA gets the
// same share of every problem).  A group of one is the plain single-GEMM call.
// ---------------------------------------------------------------------------------
static constexpr int MAX_GROUP = 8;

struct alignas(64) Problem {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  GemmParams g;
  int item0;             // first work item of this problem
  int pad_[3];
};
struct alignas(64) Group {
  Problem p[MAX_GROUP];
  int n;                 // problems
  int total_items;
};

// advance `cur` to the problem that owns `item` (items are visited in increasing order)
__device__ __forceinline__ int find_problem(const Group& grp, int cur, int item) {
  while (cur + 1 < grp.n && item >= grp.p[cur + 1].item0) ++cur;
  return cur;
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS1, 1)
pl_gemm_kernel(const __grid_constant__ Group grp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // SWIZZLE_128B tiles must sit on 1024 B boundaries
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;          // [2]       MMA -> epilogue (chunk ready)
  uint64_t* acc_empty = bars + 2 * STAGES + 2;     // [2]       epilogue -> MMA (chunk drained), count EPI_WARPS
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_items = grp.total_items;

  if (warp == 0 && lane < grp.n) {       // every problem's descriptors: each CTA visits every problem
    tma_prefetch_desc(&grp.p[lane].a_hi);
    tma_prefetch_desc(&grp.p[lane].a_lo);
    tma_prefetch_desc(&grp.p[lane].b_hi);
    tma_prefetch_desc(&grp.p[lane].b_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(smem_u32(&acc_full[b]), 1);
        mbar_init(smem_u32(&acc_empty[b]), EPI_WARPS1);
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
      uint32_t s = 0, ph = 0;
      int cur = 0;
      const uint32_t smem0 = smem_u32(smem);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        cur = find_problem(grp, cur, item);
        const Problem& pr = grp.p[cur];
        const GemmParams& g = pr.g;
        const Item it = decode_item(g, item - pr.item0);
        for (int kb = 0; kb < it.nkb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, STAGE_BYTES);
          const uint32_t base = smem0 + s * STAGE_BYTES;
          const int kc = it.kb0 + kb;
          // K-major box {32, 128 rows, 1 kb} at (0, row0, kc); MN-major box {32, 32 rows, 4 kb} at (0, kc*32, mn0/32)
          const int a1 = g.a_mn ? kc * BK : it.m0, a2 = g.a_mn ? (it.m0 >> 5) : kc;
          const int b1 = g.b_mn ? kc * BK : it.n0, b2 = g.b_mn ? (it.n0 >> 5) : kc;
          tma_load_3d(&pr.a_hi, fb, base + 0 * TILE_BYTES, 0, a1, a2);
          tma_load_3d(&pr.a_lo, fb, base + 1 * TILE_BYTES, 0, a1, a2);
          tma_load_3d(&pr.b_hi, fb, base + 2 * TILE_BYTES, 0, b1, b2);
          tma_load_3d(&pr.b_lo, fb, base + 3 * TILE_BYTES, 0, b1, b2);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The warp runs the loop converged (all lanes wait on the barriers); only the issue is under
    // elect.sync, so every operand is warp-uniform.  A 128x128x8 TF32 MMA retires every 64 clk: the
    // issue loop keeps ring counters incremental and builds descriptors from 32-bit halves.
    {
      const uint32_t smem0 = smem_u32(smem);
      uint32_t s = 0, ph = 0, gchunk = 0;
      int cur = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        cur = find_problem(grp, cur, item);
        const GemmParams& g = grp.p[cur].g;
        const Item it = decode_item(g, item - grp.p[cur].item0);
        // B_hi and B_lo tiles are adjacent in the stage, so ONE N=256 MMA computes a_hi x [b_hi | b_lo] (hi*hi into
        // columns [0,128), hi*lo into [128,256) of the chunk buffer) and a second N=128 MMA adds a_lo x b_hi to the
        // cross-term half: 20 KiB of operand reads per K=8 step instead of 24 (the 128x128 single-CTA tile is
        // bound by shared-memory traffic, not by MMA issue), same 192 clk of tensor work.
        const uint32_t idesc256 = make_idesc(BM, 2 * BN, g.a_mn, g.b_mn);
        const uint32_t idesc128 = make_idesc(BM, BN, g.a_mn, g.b_mn);
        const uint32_t dah = desc_hi_word(g.a_mn), dbh = desc_hi_word(g.b_mn);
        const uint32_t a_lo0 = desc_lo_word(smem0, g.a_mn);
        const uint32_t b_lo0 = desc_lo_word(smem0 + 2 * TILE_BYTES, g.b_mn);
        const uint32_t a_step = g.a_mn ? (1024u >> 4) : (32u >> 4);   // address-field advance per K=8 MMA
        const uint32_t b_step = g.b_mn ? (1024u >> 4) : (32u >> 4);
        for (int kb = 0; kb < it.nkb; kb += CHUNK, ++gchunk) {
          const uint32_t b = gchunk & 1;
          mbar_wait(smem_u32(&acc_empty[b]), ((gchunk >> 1) & 1) ^ 1);      // chunk buffer drained
          tc_fence_after();
          const uint32_t acc = tmem_base + b * 256;     // [H: 128 cols | S: 128 cols]
          const int nk = min(CHUNK, it.nkb - kb);
          uint32_t accum = 0;                      // first MMA of the chunk overwrites both halves
          for (int kk = 0; kk < nk; ++kk) {
            mbar_wait(smem_u32(&full_bar[s]), ph);
            tc_fence_after();
            const uint32_t so = s * (STAGE_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 8; ++k) {
                const uint32_t a_hi = a_lo0 + so + k * a_step, a_lo = a_hi + (TILE_BYTES >> 4);
                const uint32_t b_hi = b_lo0 + so + k * b_step;
                umma_tf32(acc, a_hi, b_hi, dah, dbh, idesc256, (k == 0) ? accum : 1u);       // hi*hi | hi*lo
                umma_tf32(acc + 128, a_lo, b_hi, dah, dbh, idesc128, 1u);                    // + lo*hi
              }
              umma_commit(smem_u32(&empty_bar[s]));  // frees this smem stage when the MMAs retire
              if (kk == nk - 1) umma_commit(smem_u32(&acc_full[b]));   // chunk complete
            }
            __syncwarp();
            accum = 1u;
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ================= epilogue warps 2..17 =================
    const int quad = warp & 3;                       // TMEM lane quadrant a warp may read = warp % 4
    const int cgrp = (warp - 2) >> 2;                // which 32 of the tile's 128 columns
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t col_base = (uint32_t)(cgrp * 32);
    float* stage = epi_stage + (warp - 2) * EPI_STAGE_FLOATS1;
    uint32_t gchunk = 0;
    int cur = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      cur = find_problem(grp, cur, item);
      const GemmParams& g = grp.p[cur].g;
      const Item it = decode_item(g, item - grp.p[cur].item0);
      const int mrow0 = it.m0 + quad * 32;
      const int ncol0 = it.n0 + (int)col_base;        // first output column of this warp
      const int my_row = mrow0 + lane;
      // ReLU mask: one sign-bit word per (row, 32-column block), fetched before the accumulators are awaited
      uint32_t mw = 0xffffffffu;
      if (EPI == EPI_MASK && g.mask_bits) {
        const int kbo = ncol0 >> 5;
        mw = (my_row < g.M && kbo < g.out_nkb) ? __ldg(g.mask_bits + (size_t)kbo * g.M + my_row) : 0u;
      }
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      const int nchunks = (it.nkb + CHUNK - 1) / CHUNK;
      for (int c = 0; c < nchunks; ++c, ++gchunk) {
        const uint32_t b = gchunk & 1;
        mbar_wait(smem_u32(&acc_full[b]), (gchunk >> 1) & 1);
        tc_fence_after();
        {
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(tmem_base + lane_base + b * 256 + col_base, r0);          // hi*hi partial sums of this chunk
          tmem_ld32_nowait(tmem_base + lane_base + b * 256 + 128 + col_base, r1);    // cross terms of this chunk
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r0[j]);   // fp32 RN adds
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r1[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty[b]));
      }
      // ---- tile output: this warp's 32 rows x 32 columns (one plane row segment of 128 B per row) ----
      float* dense = g.out;
      if (EPI == EPI_PARTIAL) dense += (size_t)it.split * g.M * g.N;
      const bool out_planes = g.out_planes != 0;
      const bool dense_vec = !out_planes && ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(dense) & 15) == 0);
      const int rows_ok = min(32, g.M - mrow0);          // warp-uniform; <= 0: nothing to write
      emit_slice<EPI, 16>(g, out_planes, acc, mw, stage, lane, mrow0, ncol0, rows_ok, dense, dense_vec);
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
// CTA-pair GEMM kernel (tcgen05 cta_group::2): one 256 x 256 tile per pair of SMs.
//
// Per SM and k-block the pair tile moves (128 A rows + 128 B rows) x 32 x 8 B = 64 KiB for 1536 clk of
// tensor work -- half the operand traffic per flop of the 128x128 single-CTA tile, which is what that
// kernel is bound by (TMA latency cover and shared-memory traffic) on the big layers.
//   CTA rank r of the pair owns A rows / D rows [m0 + 128 r, +128) and supplies B rows
//   [n0 + r n_inst/2, + n_inst/2); the leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256,
//   N = n_inst <= 256, trimmed to the live columns in steps of 64) for both SMs.
//   TMEM per SM (512 columns): H [0,256) = hi*hi partial sums of ONE 128-K chunk, S [256,512) = cross
//   terms of the whole tile.  H is single-buffered: the next chunk starts with its 8 cross-term MMAs
//   (1024 clk of tensor work) while the epilogue warps of both CTAs drain H into registers.
//   Barriers: TMA of both CTAs -> leader's full[s] (tx bytes of both); commit multicast -> both CTAs'
//   empty[s] / acc_full; epilogue warps of both CTAs -> leader's acc_empty / s_empty (remote arrive).
// Arithmetic (MMA order per accumulator, RN register adds) is identical to pl_gemm_kernel.
// ---------------------------------------------------------------------------------
static constexpr int BM2 = 256, BN2 = 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals come from other CTAs of the cluster
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((it & 1023u) == 1023u) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, uint32_t bar_cluster, uint32_t dst, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t da_hi, uint32_t db_hi,
                                              uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(da_lo), "r"(db_lo), "r"(da_hi), "r"(db_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

struct Item2 {
  int m0, n0, kb0, nkb, split, n_inst;
};
__device__ __forceinline__ Item2 decode_item2(const GemmParams& g, int item) {
  Item2 it;
  const int tiles = g.tiles_m * g.tiles_n;       // 256 x 256 tiles
  it.split = item / tiles;
  const int t = item - it.split * tiles;
  const int tm = t / g.tiles_n;
  it.m0 = tm * BM2;
  it.n0 = (t - tm * g.tiles_n) * BN2;
  it.kb0 = it.split * g.kb_per_split;
  it.nkb = min(g.total_kb, it.kb0 + g.kb_per_split) - it.kb0;
  it.n_inst = min(BN2, ((g.N - it.n0 + 63) >> 6) << 6);   // live columns, in steps of 64 (32 per CTA half)
  return it;
}

template <int EPI, int OUT_PLANES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
pl_gemm2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const GemmParams g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]  leader's: TMA of both CTAs -> MMA
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]  each CTA's: MMA commit (multicast) -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;          // [1]       each CTA's: MMA commit (multicast) -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 1;     // [1]       leader's: epilogue warps of both CTAs -> MMA (H drained)
  uint64_t* s_empty = bars + 2 * STAGES + 2;       // [1]       leader's: epilogue warps of both CTAs -> MMA (S drained)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 3);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
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
      mbar_init(smem_u32(acc_full), 1);
      mbar_init(smem_u32(acc_empty), 2 * EPI_WARPS);
      mbar_init(smem_u32(s_empty), 2 * EPI_WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs: own A rows, own half of B) =================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      const uint32_t smem0 = smem_u32(smem);
      for (int item = pair; item < n_items; item += npairs) {
        const Item2 it = decode_item2(g, item);
        const int a_row = it.m0 + (int)rank * 128;
        const int b_row = it.n0 + (int)rank * (it.n_inst >> 1);
        for (int kb = 0; kb < it.nkb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[s]), 2 * STAGE_BYTES);   // bytes of both CTAs
          const uint32_t fb = mapa_rank(smem_u32(&full_bar[s]), 0);                 // leader's barrier
          const uint32_t base = smem0 + s * STAGE_BYTES;
          const int kc = it.kb0 + kb;
          const int a1 = g.a_mn ? kc * BK : a_row, a2 = g.a_mn ? (a_row >> 5) : kc;
          const int b1 = g.b_mn ? kc * BK : b_row, b2 = g.b_mn ? (b_row >> 5) : kc;
          tma_load_3d_2sm(&map_a_hi, fb, base + 0 * TILE_BYTES, 0, a1, a2);
          tma_load_3d_2sm(&map_a_lo, fb, base + 1 * TILE_BYTES, 0, a1, a2);
          tma_load_3d_2sm(&map_b_hi, fb, base + 2 * TILE_BYTES, 0, b1, b2);
          tma_load_3d_2sm(&map_b_lo, fb, base + 3 * TILE_BYTES, 0, b1, b2);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      const uint32_t dah = desc_hi_word(g.a_mn), dbh = desc_hi_word(g.b_mn);
      const uint32_t smem0 = smem_u32(smem);
      const uint32_t a_lo0 = desc_lo_word(smem0, g.a_mn);
      const uint32_t b_lo0 = desc_lo_word(smem0 + 2 * TILE_BYTES, g.b_mn);
      const uint32_t a_step = g.a_mn ? (1024u >> 4) : (32u >> 4);
      const uint32_t b_step = g.b_mn ? (1024u >> 4) : (32u >> 4);
      const uint32_t acc_h = tmem_base, acc_s = tmem_base + 256;
      uint32_t s = 0, ph = 0, gchunk = 0, tile_i = 0;
      for (int item = pair; item < n_items; item += npairs, ++tile_i) {
        const Item2 it = decode_item2(g, item);
        const uint32_t idesc = make_idesc(BM2, it.n_inst, g.a_mn, g.b_mn);
        mbar_wait_cluster(smem_u32(s_empty), (tile_i & 1) ^ 1);     // cross-term accumulator drained by both CTAs
        tc_fence_after();
        uint32_t s_accum = 0;
        for (int kb = 0; kb < it.nkb; kb += CHUNK, ++gchunk) {
          const int nk = min(CHUNK, it.nkb - kb);
          uint32_t h_accum = 0;
          for (int kk = 0; kk < nk; ++kk) {
            mbar_wait(smem_u32(&full_bar[s]), ph);
            tc_fence_after();
            const uint32_t so = s * (STAGE_BYTES >> 4);
            if (kk == 0) {
              // new chunk: cross terms first, so the tensor pipe stays busy while H is being drained
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                  const uint32_t a_hi = a_lo0 + so + k * a_step, a_lo = a_hi + (TILE_BYTES >> 4);
                  const uint32_t b_hi = b_lo0 + so + k * b_step, b_lo = b_hi + (TILE_BYTES >> 4);
                  umma_tf32_2sm(acc_s, a_lo, b_hi, dah, dbh, idesc, (k == 0) ? s_accum : 1u);
                  umma_tf32_2sm(acc_s, a_hi, b_lo, dah, dbh, idesc, 1u);
                }
              }
              __syncwarp();
              mbar_wait_cluster(smem_u32(acc_empty), (gchunk & 1) ^ 1);   // H drained by both CTAs
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                  const uint32_t a_hi = a_lo0 + so + k * a_step;
                  const uint32_t b_hi = b_lo0 + so + k * b_step;
                  umma_tf32_2sm(acc_h, a_hi, b_hi, dah, dbh, idesc, (k == 0) ? 0u : 1u);
                }
                umma_commit_2sm(smem_u32(&empty_bar[s]));
                if (nk == 1) umma_commit_2sm(smem_u32(acc_full));
              }
            } else {
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                  const uint32_t a_hi = a_lo0 + so + k * a_step, a_lo = a_hi + (TILE_BYTES >> 4);
                  const uint32_t b_hi = b_lo0 + so + k * b_step, b_lo = b_hi + (TILE_BYTES >> 4);
                  umma_tf32_2sm(acc_s, a_lo, b_hi, dah, dbh, idesc, 1u);
                  umma_tf32_2sm(acc_s, a_hi, b_lo, dah, dbh, idesc, 1u);
                  umma_tf32_2sm(acc_h, a_hi, b_hi, dah, dbh, idesc, 1u);
                }
                umma_commit_2sm(smem_u32(&empty_bar[s]));
                if (kk == nk - 1) umma_commit_2sm(smem_u32(acc_full));
              }
            }
            __syncwarp();
            s_accum = 1u;
            (void)h_accum;
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ================= epilogue warps 2..9 (both CTAs; 32 rows x 128 columns each) =================
    const int quad = warp & 3;
    const int cgrp = (warp - 2) >> 2;                // which 128 of the tile's 256 columns
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t col_base = (uint32_t)(cgrp * 128);
    float* stage = epi_stage + (warp - 2) * EPI_STAGE_FLOATS;
    const uint32_t acc_empty_leader = mapa_rank(smem_u32(acc_empty), 0);
    const uint32_t s_empty_leader = mapa_rank(smem_u32(s_empty), 0);
    uint32_t gchunk = 0;
    for (int item = pair; item < n_items; item += npairs) {
      const Item2 it = decode_item2(g, item);
      const int mrow0 = it.m0 + (int)rank * 128 + quad * 32;
      const int ncol0 = it.n0 + (int)col_base;
      const int my_row = mrow0 + lane;
      const bool cols_live = ncol0 < g.N;            // warp-uniform: does this warp own any live column?
      uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (EPI == EPI_MASK && g.mask_bits) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kbo = (ncol0 >> 5) + q;
          mw[q] = (my_row < g.M && kbo < g.out_nkb) ? __ldg(g.mask_bits + (size_t)kbo * g.M + my_row) : 0u;
        }
      }
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      const int nchunks = (it.nkb + CHUNK - 1) / CHUNK;
      for (int c = 0; c < nchunks; ++c, ++gchunk) {
        mbar_wait(smem_u32(acc_full), gchunk & 1);
        tc_fence_after();
        if (cols_live) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            uint32_t r0[16], r1[16];
            tmem_ld16_nowait(tmem_base + lane_base + col_base + t * 32, r0);
            tmem_ld16_nowait(tmem_base + lane_base + col_base + t * 32 + 16, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {           // fp32 RN adds
              acc[t * 32 + j] += __uint_as_float(r0[j]);
              acc[t * 32 + 16 + j] += __uint_as_float(r1[j]);
            }
          }
          if (c == nchunks - 1) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint32_t r0[16], r1[16];
              tmem_ld16_nowait(tmem_base + lane_base + 256 + col_base + t * 32, r0);
              tmem_ld16_nowait(tmem_base + lane_base + 256 + col_base + t * 32 + 16, r1);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                acc[t * 32 + j] += __uint_as_float(r0[j]);
                acc[t * 32 + 16 + j] += __uint_as_float(r1[j]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(acc_empty_leader);
          if (c == nchunks - 1) mbar_arrive_cluster(s_empty_leader);
        }
      }
      if (cols_live) {
        float* dense = g.out;
        if (EPI == EPI_PARTIAL) dense += (size_t)it.split * g.M * g.N;
        const bool dense_vec = !OUT_PLANES && ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(dense) & 15) == 0);
        const int rows_ok = min(32, g.M - mrow0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          emit_slice<EPI, 32>(g, OUT_PLANES != 0, &acc[q * 32], mw[q], stage, lane, mrow0, ncol0 + q * 32, rows_ok, dense, dense_vec);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA frees TMEM / exits while the peer may still use its smem or barriers
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------
// dense <-> planes conversion (inputs x / labels-side gradients / weights; everything
// between two GEMMs is written as planes by the producing epilogue instead)
// ---------------------------------------------------------------------------------
// src[rows, cols] row-major -> hi/lo[nkb][rows][32] (zero padded in cols) + sign bits[nkb][rows]
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, uint32_t* __restrict__ bits,
             int rows, int cols, int nkb) {
  const int vec_per_row = nkb * 8;                       // float4 per padded row; 8 consecutive threads = one k-block
  const int64_t nvec = (int64_t)rows * vec_per_row;
  const int64_t nvec_pad = (nvec + 31) & ~(int64_t)31;   // whole warps run the loop (shuffles below)
  const bool vec_src = ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec_pad; i += (int64_t)gridDim.x * blockDim.x) {
    const bool live = i < nvec;
    const int r = live ? (int)(i / vec_per_row) : 0;
    const int c = live ? (int)(i % vec_per_row) * 4 : 0;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      if (vec_src && c + 3 < cols) {
        float4 t = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * cols + c));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (c + q < cols) ? __ldg(src + (size_t)r * cols + c + q) : 0.f;
      }
    }
    float h[4], l[4];
    uint32_t nib = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      split_tf32(v[q], h[q], l[q]);
      nib |= (v[q] > 0.f) ? (1u << q) : 0u;
    }
    uint32_t w = nib << (4 * (threadIdx.x & 7));
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    w |= __shfl_xor_sync(0xffffffffu, w, 4);
    if (live) {
      const size_t dst = ((size_t)(c >> 5) * rows + r) * 32 + (c & 31);
      *reinterpret_cast<float4*>(hi + dst) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(lo + dst) = make_float4(l[0], l[1], l[2], l[3]);
      if ((threadIdx.x & 7) == 0) bits[(size_t)(c >> 5) * rows + r] = w;
    }
  }
}

// hi/lo[nkb][rows][32] -> dst[rows, cols] = hi + lo
__global__ void __launch_bounds__(256)
merge_kernel(const float* __restrict__ hi, const float* __restrict__ lo, float* __restrict__ dst, int rows, int cols,
             int nkb) {
  const int64_t n = (int64_t)rows * nkb * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c32 = (int)(i & 31);
    const int64_t t = i >> 5;
    const int r = (int)(t % rows);
    const int c = (int)(t / rows) * 32 + c32;
    if (c < cols) dst[(size_t)r * cols + c] = hi[i] + lo[i];
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
      rc = fail(ADN_ERR_CUDA, "pl::init: cuTensorMapEncodeTiled entry point unavailable");
      return;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    bool ok = true;
#define ADN_PL_ATTR(E) \
  ok = ok && (cudaFuncSetAttribute(pl_gemm_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess)
    ADN_PL_ATTR(EPI_BIAS_ACT); ADN_PL_ATTR(EPI_MASK); ADN_PL_ATTR(EPI_PARTIAL);
#undef ADN_PL_ATTR
#define ADN_PL_ATTR2(E, P) \
  ok = ok && (cudaFuncSetAttribute(pl_gemm2_kernel<E, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess)
    ADN_PL_ATTR2(EPI_BIAS_ACT, 0); ADN_PL_ATTR2(EPI_BIAS_ACT, 1);
    ADN_PL_ATTR2(EPI_MASK, 0); ADN_PL_ATTR2(EPI_MASK, 1);
    ADN_PL_ATTR2(EPI_PARTIAL, 0);
#undef ADN_PL_ATTR2
    if (!ok) {
      (void)cudaGetLastError();
      rc = fail(ADN_ERR_CUDA, "pl::init: cudaFuncSetAttribute(smem=%d) failed", SMEM_BYTES);
    }
  });
  return rc;
}

int64_t plane_floats(int64_t rows, int64_t cols) { return align_up(rows * ceil_div(cols, BK) * BK, 64); }
int64_t bits_words(int64_t rows, int64_t cols) { return align_up(rows * ceil_div(cols, BK), 64); }
int64_t planes_bytes(int64_t rows, int64_t cols) {
  return (2 * plane_floats(rows, cols) + bits_words(rows, cols)) * (int64_t)sizeof(float);
}
static inline uint32_t* bits_of(float* planes, int64_t rows, int64_t cols) {
  return reinterpret_cast<uint32_t*>(planes + 2 * plane_floats(rows, cols));
}
static inline const uint32_t* bits_of(const float* planes, int64_t rows, int64_t cols) {
  return reinterpret_cast<const uint32_t*>(planes + 2 * plane_floats(rows, cols));
}

// a plane tensor viewed as a GEMM operand
struct Operand {
  const float* hi;
  const float* lo;
  int64_t rows, nkb;
  int mn_major;
};
static Operand operand(const float* planes, int64_t rows, int64_t cols, int mn_major) {
  return Operand{planes, planes + plane_floats(rows, cols), rows, ceil_div(cols, BK), mn_major};
}

static int make_map(CUtensorMap* map, const float* plane, int64_t rows, int64_t nkb, int mn_major) {
  if (!g_encode) return fail(ADN_ERR_CUDA, "pl: adn_init() was not called");
  cuuint64_t gdim[3] = {(cuuint64_t)BK, (cuuint64_t)rows, (cuuint64_t)nkb};
  cuuint64_t gstride[2] = {(cuuint64_t)BK * sizeof(float), (cuuint64_t)rows * BK * sizeof(float)};
  cuuint32_t box_k[3] = {(cuuint32_t)BK, 128u, 1u};
  cuuint32_t box_mn[3] = {(cuuint32_t)BK, 32u, 4u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(plane), gdim, gstride,
                        mn_major ? box_mn : box_k, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ADN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld nkb=%lld mn=%d", (int)r,
                                     (long long)rows, (long long)nkb, mn_major);
  return ADN_OK;
}

// Measured on B200 (profiles/r1d_pair_vs_single.txt): under the ~1 kW power cap the CTA-pair kernel and the
// single-CTA kernel sit at the same ~0.85-0.9 of the clock-limited TF32 rate on the big layers and the pair
// kernel loses on small problems, so the grouped single-CTA kernel is the default; ADN_PL_PAIR=1 routes every
// GEMM through the pair kernel instead (tests check both are bit-identical).
static bool use_pair() {
  static const int env = getenv("ADN_PL_PAIR") ? atoi(getenv("ADN_PL_PAIR")) : 0;
  return env == 1;
}

// one GEMM of a group: operands + epilogue description (tiles / item numbering are filled at launch)
struct GemmDesc {
  Operand a, b;
  GemmParams g;
};

static int encode_maps(const GemmDesc& d, CUtensorMap* a_hi, CUtensorMap* a_lo, CUtensorMap* b_hi, CUtensorMap* b_lo,
                       const char* what) {
  if ((reinterpret_cast<uintptr_t>(d.a.hi) | reinterpret_cast<uintptr_t>(d.a.lo) | reinterpret_cast<uintptr_t>(d.b.hi) |
       reinterpret_cast<uintptr_t>(d.b.lo)) & 127)
    return fail(ADN_ERR_INVALID, "%s: plane buffers must be 256 B aligned", what);
  int rc;
  if ((rc = make_map(a_hi, d.a.hi, d.a.rows, d.a.nkb, d.a.mn_major))) return rc;
  if ((rc = make_map(a_lo, d.a.lo, d.a.rows, d.a.nkb, d.a.mn_major))) return rc;
  if ((rc = make_map(b_hi, d.b.hi, d.b.rows, d.b.nkb, d.b.mn_major))) return rc;
  if ((rc = make_map(b_lo, d.b.lo, d.b.rows, d.b.nkb, d.b.mn_major))) return rc;
  return ADN_OK;
}

template <int EPI>
static int launch_pair(const GemmDesc& d, cudaStream_t st, const char* what) {
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = encode_maps(d, &ma_hi, &ma_lo, &mb_hi, &mb_lo, what))) return rc;
  GemmParams g = d.g;
  g.a_mn = d.a.mn_major;
  g.b_mn = d.b.mn_major;
  g.tiles_m = (int)ceil_div(g.M, BM2);
  g.tiles_n = (int)ceil_div(g.N, BN2);
  const int items = g.tiles_m * g.tiles_n * g.splits;
  const int grid = 2 * std::min(items, sm_count() / 2);
  if (g.out_planes) pl_gemm2_kernel<EPI, 1><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, g);
  else pl_gemm2_kernel<EPI, 0><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, g);
  ADN_CHECK_LAUNCH(what);
  return ADN_OK;
}

// Work-item order of one problem.  Column blocks fastest lets the CTAs that run together share the A tile through
// L2 (A is the big operand of the long-K layers); row blocks fastest makes them write adjacent 16 KB runs of the
// same output k-block slab.  ADN_PL_MFAST: -1 heuristic (default), 0 / 1 force.
static int item_order_m_fastest(const GemmParams& g) {
  static const int env = getenv("ADN_PL_MFAST") ? atoi(getenv("ADN_PL_MFAST")) : -1;
  if (env >= 0) return env;
  return 0;
}

// n independent GEMMs of the same epilogue kind -> ceil(n / MAX_GROUP) persistent launches
template <int EPI>
static int launch_group(const GemmDesc* d, int n, cudaStream_t st, const char* what) {
  if (use_pair()) {
    for (int i = 0; i < n; ++i) {
      int rc = launch_pair<EPI>(d[i], st, what);
      if (rc) return rc;
    }
    return ADN_OK;
  }
  for (int i0 = 0; i0 < n; i0 += MAX_GROUP) {
    const int m = std::min(MAX_GROUP, n - i0);
    Group grp;
    memset(&grp, 0, sizeof(grp));
    int items = 0;
    for (int i = 0; i < m; ++i) {
      Problem& pr = grp.p[i];
      int rc = encode_maps(d[i0 + i], &pr.a_hi, &pr.a_lo, &pr.b_hi, &pr.b_lo, what);
      if (rc) return rc;
      pr.g = d[i0 + i].g;
      pr.g.a_mn = d[i0 + i].a.mn_major;
      pr.g.b_mn = d[i0 + i].b.mn_major;
      pr.g.tiles_m = (int)ceil_div(pr.g.M, BM);
      pr.g.tiles_n = (int)ceil_div(pr.g.N, BN);
      pr.g.m_fastest = item_order_m_fastest(pr.g);
      pr.item0 = items;
      items += pr.g.tiles_m * pr.g.tiles_n * pr.g.splits;
    }
    grp.n = m;
    grp.total_items = items;
    const int grid = std::min(items, sm_count());
    pl_gemm_kernel<EPI><<<grid, NUM_THREADS1, SMEM_BYTES, st>>>(grp);
    ADN_CHECK_LAUNCH(what);
  }
  return ADN_OK;
}

int split(const float* src, int64_t rows, int64_t cols, float* planes, cudaStream_t st) {
  const int64_t nkb = ceil_div(cols, BK);
  const int64_t nvec = rows * nkb * 8;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nvec, 256), (int64_t)sm_count() * 16));
  split_kernel<<<blocks, 256, 0, st>>>(src, planes, planes + plane_floats(rows, cols), bits_of(planes, rows, cols),
                                       (int)rows, (int)cols, (int)nkb);
  ADN_CHECK_LAUNCH("planes split");
  return ADN_OK;
}

int merge(const float* planes, int64_t rows, int64_t cols, float* dst, cudaStream_t st) {
  const int64_t nkb = ceil_div(cols, BK);
  const int64_t n = rows * nkb * 32;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16));
  merge_kernel<<<blocks, 256, 0, st>>>(planes, planes + plane_floats(rows, cols), dst, (int)rows, (int)cols, (int)nkb);
  ADN_CHECK_LAUNCH("planes merge");
  return ADN_OK;
}

// dW split-K: the partial buffer bounds the split count (<= 16M floats, <= 64 splits)
static int max_dw_splits(int64_t in, int64_t out) {
  int64_t s = (16LL << 20) / std::max<int64_t>(1, in * out);
  if (s > MAX_SPLITS) s = MAX_SPLITS;
  if (s < 1) s = 1;
  return (int)s;
}

int64_t dense_bwd_workspace_bytes(int64_t batch, int64_t in, int64_t out) {
  int64_t b = align_up((int64_t)max_dw_splits(in, out) * in * out * (int64_t)sizeof(float), 256);   // dW split-K partials
  b += align_up(ceil_div(batch, 32) * in * (int64_t)sizeof(float), 256);                             // dx column sums per 32 rows
  b += align_up(64 * in * (int64_t)sizeof(float), 256);                                              // their second-level partials
  return b + 512;
}

int dense_fwd_group(const FwdOp* ops, int n, int64_t batch, cudaStream_t st) {
  if (n <= 0) return ADN_OK;
  std::vector<GemmDesc> d((size_t)n);
  for (int i = 0; i < n; ++i) {
    const FwdOp& o = ops[i];
    GemmDesc& e = d[(size_t)i];
    e.a = operand(o.xp, batch, o.in, 0);       // A = x  [M=batch, K=in]  K-major
    e.b = operand(o.wp, o.in, o.out, 1);       // B = w  [K=in, N=out]    MN-major
    GemmParams g{};
    g.M = (int)batch; g.N = (int)o.out;
    g.total_kb = (int)ceil_div(o.in, BK); g.kb_per_split = g.total_kb; g.splits = 1;
    g.bias = o.bias; g.act = o.act;
    if (o.yp) {
      g.out_planes = 1;
      g.out = o.yp; g.out_lo = o.yp + plane_floats(batch, o.out); g.out_nkb = (int)ceil_div(o.out, BK);
      g.out_bits = bits_of(o.yp, batch, o.out);
    } else {
      g.out = o.y; g.ldc = (int)o.out;
    }
    e.g = g;
  }
  return launch_group<EPI_BIAS_ACT>(d.data(), n, st, "pl dense_fwd gemm");
}

int dense_bwd_group(const BwdOp* ops, int n, int64_t batch, cudaStream_t st) {
  if (n <= 0) return ADN_OK;
  const int workers = use_pair() ? sm_count() / 2 : sm_count();
  const int tm = use_pair() ? BM2 : BM, tn = use_pair() ? BN2 : BN;
  struct Carve { float* part; float* cspart; float* cspart2; int splits; };
  std::vector<Carve> cv((size_t)n);
  std::vector<GemmDesc> dwd, dxd;
  // Split-K over the batch for the dW GEMMs of the group.  Items of one launch are dealt round-robin to the
  // CTAs, so they should all cost the same: every problem uses the same k-blocks-per-item `kps`, chosen to
  // minimise (rounds of the whole group) x (kps + per-item overhead) + reduction cost, subject to each
  // problem's partial-buffer bound.
  const int64_t kb_b = ceil_div(batch, BK);
  int64_t best_kps = kb_b;
  {
    double best_t = 1e30;
    for (int s0 = 1; s0 <= MAX_SPLITS && s0 <= kb_b; ++s0) {
      const int64_t kps = ceil_div(kb_b, s0);
      int64_t items = 0, max_s_used = 1;
      for (int i = 0; i < n; ++i) {
        if (!ops[i].dw) continue;
        int64_t k_i = std::max<int64_t>(kps, ceil_div(kb_b, max_dw_splits(ops[i].in, ops[i].out)));
        const int64_t s_i = ceil_div(kb_b, k_i);
        items += ceil_div(ops[i].in, tm) * ceil_div(ops[i].out, tn) * s_i;
        max_s_used = std::max(max_s_used, s_i);
      }
      if (items == 0) break;
      const double t = (double)ceil_div(items, workers) * ((double)kps + 6.0) + 0.75 * (double)max_s_used;
      if (t < best_t) { best_t = t; best_kps = kps; }
    }
  }
  for (int i = 0; i < n; ++i) {
    const BwdOp& o = ops[i];
    if (!o.ws || o.ws_bytes < dense_bwd_workspace_bytes(batch, o.in, o.out))
      return fail(ADN_ERR_WORKSPACE, "pl dense_bwd: op %d workspace %lld < %lld bytes", i, (long long)o.ws_bytes,
                  (long long)dense_bwd_workspace_bytes(batch, o.in, o.out));
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(o.ws) + 255) & ~(uintptr_t)255);
    const int max_s = max_dw_splits(o.in, o.out);
    Carve& c = cv[(size_t)i];
    c.part = reinterpret_cast<float*>(p);
    p += align_up((int64_t)max_s * o.in * o.out * (int64_t)sizeof(float), 256);
    c.cspart = reinterpret_cast<float*>(p);
    p += align_up(ceil_div(batch, 32) * o.in * (int64_t)sizeof(float), 256);
    c.cspart2 = reinterpret_cast<float*>(p);
    c.splits = 1;
    if (o.dw) {
      // ---- dW[in,out] = X^T dZ : A = Xp MN-major (M=in), B = dZp MN-major (N=out), K = batch, split-K ----
      GemmDesc e;
      e.a = operand(o.xp, batch, o.in, 1);
      e.b = operand(o.dzp, batch, o.out, 1);
      GemmParams g{};
      g.M = (int)o.in; g.N = (int)o.out; g.ldc = (int)o.out;
      g.total_kb = (int)kb_b; g.kb_per_split = (int)std::max<int64_t>(best_kps, ceil_div(kb_b, max_s));
      g.splits = (int)ceil_div(kb_b, g.kb_per_split);
      g.out = (g.splits == 1) ? o.dw : c.part;
      c.splits = g.splits;
      e.g = g;
      dwd.push_back(e);
    }
    if (o.dxp || o.dx) {
      // ---- dX[batch,in] = dZ W^T : A = dZp K-major (K=out), B = Wp K-major (N=in, K=out); ReLU mask = sign bits of X ----
      GemmDesc e;
      e.a = operand(o.dzp, batch, o.out, 0);
      e.b = operand(o.wp, o.in, o.out, 0);
      GemmParams g{};
      g.M = (int)batch; g.N = (int)o.in;
      g.total_kb = (int)ceil_div(o.out, BK); g.kb_per_split = g.total_kb; g.splits = 1;
      g.mask_bits = o.x_relu_mask ? bits_of(o.xp, batch, o.in) : nullptr;
      g.out_nkb = (int)ceil_div(o.in, BK);
      g.colsum_part = o.dx_colsum ? c.cspart : nullptr;
      g.colsum_ld = (int)o.in;
      if (o.dxp) {
        g.out_planes = 1;
        g.out = o.dxp; g.out_lo = o.dxp + plane_floats(batch, o.in);
      } else {
        g.out = o.dx; g.ldc = (int)o.in;
      }
      e.g = g;
      dxd.push_back(e);
    }
  }
  int rc;
  if (!dwd.empty()) {
    if ((rc = launch_group<EPI_PARTIAL>(dwd.data(), (int)dwd.size(), st, "pl dW gemm"))) return rc;
    std::vector<simt::ReduceJob> jobs;
    for (int i = 0; i < n; ++i)
      if (ops[i].dw && cv[(size_t)i].splits > 1)
        jobs.push_back(simt::ReduceJob{cv[(size_t)i].part, ops[i].dw, ops[i].in * ops[i].out, cv[(size_t)i].splits,
                                       ops[i].in * ops[i].out});
    if ((rc = simt::reduce_partials_group(jobs.data(), (int)jobs.size(), st))) return rc;
  }
  if (!dxd.empty()) {
    if ((rc = launch_group<EPI_MASK>(dxd.data(), (int)dxd.size(), st, "pl dX gemm"))) return rc;
    // column sums of each [ceil(batch/32), in] partial matrix, fixed order
    std::vector<simt::ColsumJob> jobs;
    for (int i = 0; i < n; ++i)
      if (ops[i].dx_colsum && (ops[i].dxp || ops[i].dx))
        jobs.push_back(simt::ColsumJob{cv[(size_t)i].cspart, ops[i].dx_colsum, ceil_div(batch, 32), ops[i].in,
                                       cv[(size_t)i].cspart2});
    if ((rc = simt::colsum_group(jobs.data(), (int)jobs.size(), st))) return rc;
  }
  return ADN_OK;
}

int dense_fwd(const float* xp, const float* wp, const float* bias, float* yp, float* y, int64_t batch, int64_t in,
              int64_t out, int act, cudaStream_t st) {
  const FwdOp op{xp, wp, bias, yp, y, in, out, act};
  return dense_fwd_group(&op, 1, batch, st);
}

int dense_bwd(const float* xp, const float* wp, const float* dzp, float* dxp, float* dx, float* dx_colsum, float* dw,
              int64_t batch, int64_t in, int64_t out, int x_relu_mask, void* ws, int64_t ws_bytes, cudaStream_t st) {
  const BwdOp op{xp, wp, dzp, dxp, dx, dx_colsum, dw, in, out, x_relu_mask, ws, ws_bytes};
  return dense_bwd_group(&op, 1, batch, st);
}

}  // namespace pl
}  // namespace adn
