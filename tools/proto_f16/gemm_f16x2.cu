// PROTOTYPE, NOT PART OF THE LIBRARY BUILD (tools/proto_f16/run.py compiles and runs it on a B200).
// Status: compiles for sm_100a; NOT yet run on a GPU (round 1 ended without GPU minutes for it).
//
// Question it answers ("fp16 hi/lo planes"): can the dense path keep fp32-level accuracy with
// 2-byte operands and kind::f16 MMAs (twice the kind::tf32 rate, half the operand bytes)?
//
//   x = hi + 2^-11 * lo',  hi = fp16(x),  lo' = fp16((x - hi) * 2^11)         (both 11 significant bits)
//   C = A B^T:   H = sum a_hi b_hi,   S = sum (a_lo' b_hi + a_hi b_lo'),   C ~= H + 2^-11 S      (lo' lo' dropped)
//
// One 128 x 128 output tile per CTA; A [M, K] and B [N, K] are both K-major (row-major with K contiguous), given as
// four fp16 matrices (a_hi, a_lo, b_hi, b_lo) split on the host side by run.py.  Warp 0: TMA producer (3 stages of
// 4 tiles [128 rows x 64 k] = 64 KiB); warp 1: MMA issuer (per stage 4 k-steps of K=16: 1 MMA into H, 2 into S);
// warps 2-5: epilogue (tcgen05.ld both accumulators, H + 2^-11 S, fp32 row-major store).
// H is accumulated in TMEM for `chunk_kb` k-blocks at a time and added into registers in between (two-level
// accumulation, as planes.cu does for TF32: the tensor core truncates its fp32 accumulator on every add);
// chunk_kb = 0 keeps one TMEM chain over the whole K so the two can be compared.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 64;          // BK fp16 = 128 B = one SWIZZLE_128B row
constexpr int TILE_BYTES = BM * BK * 2;             // 16 KiB
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // a_hi a_lo b_hi b_lo
constexpr int THREADS = 64 + 128;                   // TMA warp, MMA warp, 4 epilogue warps
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 256;                      // H at column 0, S at column 128

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
    if ((it & 1023u) == 1023u) {      // bounded spin: trap instead of hanging the box
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major SWIZZLE_128B descriptor: start >> 4 | LBO(unused) 1 << 16 ; hi word: SBO 1024 >> 4, version 1, layout type 2
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi() { return (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29); }
// kind::f16 instruction descriptor: D = f32 (1 << 4), A = B = f16 (format 0 at [7,10) and [10,13)), K-major both,
// N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t da, uint32_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(da), "r"(db), "r"(desc_hi()), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

struct Maps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

// chunk_kb: k-blocks (of 64) accumulated in TMEM before the epilogue adds them into registers (0 = all of K)
__global__ void __launch_bounds__(THREADS, 1)
gemm_f16x2_kernel(const __grid_constant__ Maps maps, float* __restrict__ out, int M, int N, int K, int chunk_kb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;        // [STAGES] MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;     // MMA -> epilogue (one chunk done)
  uint64_t* acc_empty = bars + 2 * STAGES + 1;  // epilogue -> MMA (accumulators drained), count 4 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / BK;
  const int ck = chunk_kb > 0 ? chunk_kb : nkb;
  const int nchunks = (nkb + ck - 1) / ck;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      mbar_init(smem_u32(acc_full), 1);
      mbar_init(smem_u32(acc_empty), 4);
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
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(smem_u32(&empty_bar[s]), ((kb / STAGES) & 1) ^ 1);
        const uint32_t bar = smem_u32(&full_bar[s]);
        const uint32_t dst = smem_u32(base + s * STAGE_BYTES);
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_2d(&maps.a_hi, bar, dst + 0 * TILE_BYTES, kb * BK, m0);
        tma_load_2d(&maps.a_lo, bar, dst + 1 * TILE_BYTES, kb * BK, m0);
        tma_load_2d(&maps.b_hi, bar, dst + 2 * TILE_BYTES, kb * BK, n0);
        tma_load_2d(&maps.b_lo, bar, dst + 3 * TILE_BYTES, kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN);
      const uint32_t tH = tmem_base, tS = tmem_base + 128;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const int in_chunk = kb % ck;
        if (in_chunk == 0 && kb > 0) {
          // the previous chunk was committed below; wait until the epilogue drained H before overwriting it
          mbar_wait(smem_u32(acc_empty), ((kb / ck - 1) & 1));
          tc_fence_after();
        }
        mbar_wait(smem_u32(&full_bar[s]), (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(base + s * STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {           // K = 16 per kind::f16 MMA, +32 B per step
          const uint32_t a_hi = desc_lo(st + 0 * TILE_BYTES + ks * 32), a_lo = desc_lo(st + 1 * TILE_BYTES + ks * 32);
          const uint32_t b_hi = desc_lo(st + 2 * TILE_BYTES + ks * 32), b_lo = desc_lo(st + 3 * TILE_BYTES + ks * 32);
          umma_f16(tH, a_hi, b_hi, idesc, (in_chunk | ks) != 0);                 // H restarts with every chunk
          umma_f16(tS, a_lo, b_hi, idesc, (kb | ks) != 0);                        // S runs over the whole K
          umma_f16(tS, a_hi, b_lo, idesc, 1u);
        }
        umma_commit(smem_u32(&empty_bar[s]));          // stage free when these MMAs retire
        if (in_chunk == ck - 1 || kb == nkb - 1) umma_commit(smem_u32(acc_full));
      }
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =================
    const int quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int row = m0 + quad * 32 + lane;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(smem_u32(acc_full), c & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float t[16];
        tmem_ld16(tmem_base + lane_base + (uint32_t)c0, t);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[c0 + e] += t[e];          // RN add of the chunk's H
      }
      if (c == nchunks - 1) {
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float t[16];
          tmem_ld16(tmem_base + lane_base + (uint32_t)(128 + c0), t);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[c0 + e] = fmaf(t[e], 1.0f / 2048.0f, acc[c0 + e]);   // + 2^-11 S
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_empty));
    }
    if (row < M) {
#pragma unroll
      for (int j = 0; j < BN; j += 4)
        if (n0 + j < N)
          *reinterpret_cast<float4*>(out + (size_t)row * N + n0 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
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

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(EncodeFn enc, CUtensorMap* map, const void* ptr, int rows, int K) {
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace

// a_hi, a_lo: fp16 [M, K]; b_hi, b_lo: fp16 [N, K]; out: fp32 [M, N].  M, N multiples of 128, K of 64.
extern "C" int proto_gemm_f16x2(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* out, int M,
                                int N, int K, int chunk_kb, void* stream) {
  if (M % BM || N % BN || K % BK) return -22;
  static EncodeFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -5;
    enc = reinterpret_cast<EncodeFn>(fn);
    if (cudaFuncSetAttribute(gemm_f16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return -5;
  }
  Maps maps;
  int rc;
  if ((rc = make_map(enc, &maps.a_hi, a_hi, M, K))) return rc;
  if ((rc = make_map(enc, &maps.a_lo, a_lo, M, K))) return rc;
  if ((rc = make_map(enc, &maps.b_hi, b_hi, N, K))) return rc;
  if ((rc = make_map(enc, &maps.b_lo, b_lo, N, K))) return rc;
  dim3 grid(N / BN, M / BM);
  gemm_f16x2_kernel<<<grid, THREADS, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(maps, out, M, N, K, chunk_kb);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
