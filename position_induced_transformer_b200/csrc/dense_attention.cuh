// K2: dense (global, locality >= 1) position-attention on the 5th-generation tensor cores.
//
// A global stage is a genuine dense contraction  O[i, c] = sum_j P[i, j] V[j, c]  with
// P[i, j] = exp(s_h (v_min_i - d2_ij)) generated on the fly from coordinates (flash style: P never
// exists in HBM) and V the value features of every sample side by side (shared meshes) or of one
// sample (per-sample meshes).  The tile loop is a warp-specialised producer/consumer pipeline:
//
//   warps 0-7  "generators": two threads per tile row; per K block each evaluates 16 of the 32 weights from the
//              coordinates (exp2 with the row's known shift, so no online rescaling), splits each into a
//              TF32 high part and an fp32 residual, and writes both as UMMA operand A (K-major,
//              128-byte swizzle) into shared memory; it also keeps the fp32 row sum.
//   warps 8-15 "stagers": stream the [32 x NV] value block from global memory with 128-bit loads, split
//              hi/lo the same way, and store it as UMMA operand B in MN-major form.  For 32-bit operands the
//              MN-major canonical layout is SWIZZLE_128B_BASE32B (atoms of 4 K-rows x 128 bytes, 32-byte chunks
//              XOR-ed with the K row); with the plain SWIZZLE_128B descriptor a tf32 MN-major operand reads back
//              as zeros (measured: scripts/probe/umma_probe.cu, modes 0/1 vs 5).  No transpose is needed.
//   warp 16    one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=NV, K=8): hi*hi + lo*hi + hi*lo,
//              i.e. 3xTF32 with fp32 accumulation in TMEM (error ~2^-21, inside the 1e-5 parity budget),
//              then tcgen05.commit's the stage back to the producers through an mbarrier.
//   warps 0-7  epilogue: tcgen05.ld the accumulator rows out of TMEM (the two threads of a row take half the columns each),
//              normalise by the row sum, store.
//
// Operands cannot come through TMA here: both need the in-register hi/lo split (and A is computed, not
// loaded), so the stage hand-off is mbarrier + fence.proxy.async rather than a TMA transaction count.
//
// The same pipeline serves the backward of a dense stage:
//   DENSE_DSCALE   two A operands (P and P*d2) against the same V block -> accumulators O and W in TMEM;
//                  epilogue forms -sum_e dO_e (W_e - (m/l) O_e) / l per row and reduces it per head;
//   DENSE_DVALUES  roles swapped: tile rows are columns j, the K loop runs over rows i of every head,
//                  A = P^T / l_i, B = dO[i, (b,d)]; epilogue adds the concat pass-through and stores dU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "geometry.cuh"

namespace pit {

constexpr int DENSE_ROWS = 128;  // UMMA M
constexpr int DENSE_KB = 32;     // reduction entries per K block = one 128-byte swizzle row of tf32
// Two generator threads per tile row (16 of the 32 weights of a K block each) and twice as many stager threads: with one warp of
// each kind per SM sub-partition every MUFU / LDS / STS latency of the producers was exposed and the kernel ran at their pace
// whatever the MMA count (ncu: 3xTF32 and single-product modes took the same time, issue slots 34 % busy).
constexpr int DENSE_GEN_THREADS = 256;
constexpr int DENSE_STAGE_THREADS = 256;
constexpr int DENSE_GEN_WARPS = DENSE_GEN_THREADS / 32, DENSE_STAGE_WARPS = DENSE_STAGE_THREADS / 32;
constexpr int DENSE_MMA_WARP = DENSE_GEN_WARPS + DENSE_STAGE_WARPS;
constexpr int DENSE_THREADS = DENSE_GEN_THREADS + DENSE_STAGE_THREADS + 32;
constexpr int DENSE_STAGES = 2;

enum : int { DENSE_FWD = 0, DENSE_DSCALE = 1, DENSE_DVALUES = 2 };

struct DenseParams {
  // geometry: "owner" points index the tile rows, "reduced" points index the K loop
  const float* mesh_own;  // [(B),n_own,sd]
  const float* mesh_red;  // [(B),n_red,sd]
  const float* period;
  const float* scale;   // [H]
  const float* v_min;   // [(B),N] indexed by attention row i
  const float* rowsum;  // [(B),H,N] (backward modes)
  int n_own, n_red;     // FWD/DSCALE: N, M ; DVALUES: M, N
  int N, M, B, H, D, sd, mesh_batched;
  int width;            // value columns in total: mesh_batched ? D : B*D
  // operand B source: element (k, n) at  b_src + b_off + (n / D) * b_bstride + n % D + k * b_kstride (+ h * b_hstride)
  const float* b_src;
  int64_t b_kstride, b_bstride, b_hstride, b_off;
  // FWD
  float* out;
  int64_t ld_out, col_off;
  float* rowsum_out;
  // DSCALE
  const float* d_out;
  float* d_scale;  // [H] zero-initialised
  // DVALUES
  float* d_values;  // [B,M,D]
  int add_concat;
  int split_heads;  // DVALUES: one CTA per head (grid.z = H x samples), partial sums meet in d_values (zero-initialised) by RED
  int operand_prec; // DENSE_PREC_*: how both operands reach the tensor pipe
};

// Operand precision of the products (the accumulation is fp32 in TMEM in every case):
//   FP32   3xTF32: hi*hi + lo*hi + hi*lo, fp32 parity (~1e-6)                                   -- the default
//   TF32   one product on operands rounded to TF32: what the reference's einsum runs as under its own
//          torch.set_float32_matmul_precision('high') (pit.py:2); bound 1e-3
//   BF16   one product on operands rounded to BF16 (8 significant bits): numerically the BF16-operand / fp32-accumulate
//          product of SURVEY 8c (bound 5e-3; a BF16 value is a TF32 value, so the kind::tf32 instruction reproduces
//          kind::f16 exactly) -- issued on the TF32 pipe from 32-bit tiles, i.e. without the 2x rate and the halved
//          shared-memory traffic that 16-bit operand tiles would add.
enum : int { DENSE_PREC_FP32 = 0, DENSE_PREC_TF32 = 1, DENSE_PREC_BF16 = 2 };

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread for the CTA.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when every tcgen05 operation issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread l of warp w reads TMEM lane 32*(w%4)+l.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), SWIZZLE_128B, version 1.
//   K-major operand (A and B): rows of 128 bytes (32 tf32 along K), 8-row groups 1024 bytes apart (SBO); LBO unused (1).
//   One MMA consumes K = 8 tf32 = 32 bytes of every row: the k-th step starts 32*k bytes into the tile.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor (InstrDescriptor): D=F32, A=B=TF32, A K-major, B MN-major, M=128, N=NV.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(DENSE_ROWS >> 4) << 24);
}

// High part of the 3xTF32 split: x rounded to nearest at tf32 precision (10 explicit mantissa bits), so that the
// residual x - hi fits the next tf32 with half the truncation error of a plain mask.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// x rounded to the nearest BF16 value (ties to even), kept in an fp32 register
__device__ __forceinline__ float bf16_round(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float((u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u);
}
// high part of an operand under the precision mode `prec`
__device__ __forceinline__ float operand_hi(float x, int prec) { return prec == DENSE_PREC_BF16 ? bf16_round(x) : tf32_hi(x); }

// Byte offset of the 16-byte chunk (row r, chunk c of 8) inside a K-major 128B-swizzled tile of 128 rows.
__device__ __forceinline__ uint32_t a_chunk_offset(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }
// Operand B, MN-major SWIZZLE_128B_BASE32B: atoms of 4 K-rows x 128 bytes (32 value columns); inside an atom the
// 32-byte chunk index is XOR-ed with the K row.  Atoms of one 32-column group are contiguous along K (SBO = 512
// bytes per 4 K-rows), column groups are DENSE_KB/4 atoms apart (LBO = 4096 bytes).  One MMA (K = 8) spans two atoms.
constexpr uint32_t B_SBO = 512, B_LBO = (DENSE_KB / 4) * 512, B_KSTEP = 1024;
// Byte offset of the 16-byte piece holding columns 4*n4 .. 4*n4+3 of K-row k.
__device__ __forceinline__ uint32_t b_chunk_offset(int k, int n4) {
  const int kr = k & 3;
  return (uint32_t)((n4 >> 3) * B_LBO + (k >> 2) * B_SBO + kr * 128 + (((((n4 & 7) >> 1) ^ kr)) << 5) + (n4 & 1) * 16);
}
__device__ __forceinline__ uint64_t umma_desc_b(uint32_t smem_addr) {
  uint64_t d = umma_desc(smem_addr, B_LBO, B_SBO);
  d &= ~((uint64_t)7 << 61);
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

// Squared distance for the dense path: no mask is decided on it, so plain (FMA-contracted) arithmetic is fine.
template <int GEO>
__device__ __forceinline__ float dist2_fast(const Point<GEO>& a, float bx, float by, float period) {
  if (GEO == GEO_EUCLID1) {
    const float dx = a.x - bx;
    return dx * dx;
  } else if (GEO == GEO_EUCLID2) {
    const float dx = a.x - bx, dy = a.y - by;
    return fmaf(dy, dy, dx * dx);
  } else if (GEO == GEO_PERIODIC1) {
    float m = fabsf(a.x - bx);
    m = fminf(m, period - m);
    return m * m;
  } else {
    float mx = fabsf(a.x - bx), my = fabsf(a.y - by);
    mx = fminf(mx, period - mx);
    my = fminf(my, period - my);
    return fmaf(my, my, mx * mx);
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MODE, int NV>
struct DenseSmem {
  static constexpr int A_TILES = (MODE == DENSE_DSCALE) ? 4 : 2;  // hi, lo (and the d2-weighted pair)
  static constexpr int A_BYTES = DENSE_ROWS * DENSE_KB * 4;        // 16 KB
  static constexpr int B_BYTES = DENSE_KB * NV * 4;
  static constexpr int STAGE_BYTES = A_TILES * A_BYTES + 2 * B_BYTES;
  static constexpr int TOTAL = DENSE_STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + DENSE_GEN_WARPS * 32 * 16 /*reduced-point tables*/ + 256;
};

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
// The body takes its block coordinates as an argument so that two modes can share one launch (dense_bwd_pair_kernel).
template <int GEO, int MODE, int NV>
__device__ __forceinline__ void dense_attention_body(const DenseParams& P, const dim3 bid) {
  using L = DenseSmem<MODE, NV>;
  constexpr int ACC_COLS = (MODE == DENSE_DSCALE) ? 2 * NV : NV;
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  extern __shared__ unsigned char dense_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[DENSE_STAGES], empty_bar[DENSE_STAGES], done_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red[DENSE_GEN_THREADS / 32];
  __shared__ float half_sum[2][2][DENSE_ROWS];   // [l | m][half][row]: the two generator threads of a row exchange their partial sums

  const uint32_t raw = smem_u32(dense_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;  // swizzle atoms need 1024-byte alignment
  unsigned char* tiles_ptr = dense_smem_raw + (tiles - raw);
  float4* red_pts = reinterpret_cast<float4*>(tiles_ptr + DENSE_STAGES * L::STAGE_BYTES);  // [generator warps][32] (x, y, vmin*, post)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = bid.z;
  const int h_fixed = z % P.H;                       // FWD/DSCALE: the head of this CTA
  const int bm = P.mesh_batched ? z / P.H : 0;       // sample (per-sample meshes)
  const int dv_heads = (MODE == DENSE_DVALUES && P.split_heads) ? P.H : 1;
  const int h_dv = z % dv_heads;                     // DVALUES with split heads: the head of this CTA
  const int bm_dv = P.mesh_batched ? z / dv_heads : 0;  // DVALUES: grid.z = samples (x heads when split)
  const int own0 = bid.x * DENSE_ROWS;
  const int n0 = bid.y * NV;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const int sample = (MODE == DENSE_DVALUES) ? bm_dv : bm;
  // CTA-uniform.  The scale gradient is a difference of large sums (W - (m/l) O): operand rounding shows up amplified there
  // (measured 10 % with BF16 operands), so that mode always multiplies 3xTF32 whatever the setting.
  const int prec = (MODE == DENSE_DSCALE) ? DENSE_PREC_FP32 : P.operand_prec;
  const bool split = prec == DENSE_PREC_FP32;      // 3xTF32: the residual tiles are written and multiplied as well

  if (tid == 0) {
    for (int s = 0; s < DENSE_STAGES; ++s) {
      mbar_init(&full_bar[s], DENSE_GEN_WARPS + DENSE_STAGE_WARPS);   // one arrival per producer WARP (512 per-thread arrivals
                                                                      // on one mbarrier serialise in the shared-memory atomic unit)
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == DENSE_MMA_WARP) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int heads_in_k = (MODE == DENSE_DVALUES && !P.split_heads) ? P.H : 1;
  const int kb_per_head = (P.n_red + DENSE_KB - 1) / DENSE_KB;
  const int n_kb = kb_per_head * heads_in_k;

  if (warp < DENSE_GEN_WARPS) {
    // =========================== generators (and epilogue) ===========================
    const int r = tid & (DENSE_ROWS - 1);  // tile row
    const int half = tid / DENSE_ROWS;     // which 16 of the 32 weights of a K block (and which half of the epilogue columns)
    const int own = own0 + r;
    const bool own_ok = own < P.n_own;
    const float* mesh_own = P.mesh_own + (int64_t)sample * P.n_own * P.sd;
    const float* mesh_red = P.mesh_red + (int64_t)sample * P.n_red * P.sd;
    const Point<GEO> me = load_point<GEO>(mesh_own, own_ok ? own : 0, P.sd);
    float4* my_pts = red_pts + warp * 32;
    const float LOG2E = 1.4426950408889634f;
    // FWD/DSCALE: the row's own shift; DVALUES: per reduced row, read from the table
    float own_vmin = 0.f;
    if (MODE != DENSE_DVALUES) own_vmin = __ldg(P.v_min + (int64_t)sample * P.N + (own_ok ? own : 0));
    float lsum = 0.f, msum = 0.f;

    // The reduced point of the NEXT K block is fetched while the current block is computed, so the global-load
    // latency (the longest dependency of a K block at small problem sizes) is off the critical path.
    auto fetch = [&](int kb) -> float4 {
      const int h = (MODE == DENSE_DVALUES) ? h_dv + kb / kb_per_head : h_fixed;
      const int k = (kb % kb_per_head) * DENSE_KB + lane;
      const bool ok = kb < n_kb && k < P.n_red;
      const Point<GEO> q = load_point<GEO>(mesh_red, ok ? k : 0, P.sd);
      float bias = 0.f, post = 1.f;
      if (MODE == DENSE_DVALUES && ok) {
        bias = __ldg(P.v_min + (int64_t)sample * P.N + k) * (__ldg(P.scale + h) * LOG2E);
        post = 1.f / __ldg(P.rowsum + ((int64_t)sample * P.H + h) * P.N + k);
      }
      return make_float4(ok ? q.x : INFINITY, q.y, bias, post);  // points past the end sit at infinity: weight exp2(-inf) = 0
    };
    float4 next_pt = fetch(0);

    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % DENSE_STAGES;
      const uint32_t use = kb / DENSE_STAGES;
      const int h = (MODE == DENSE_DVALUES) ? h_dv + kb / kb_per_head : h_fixed;
      const float sc2 = __ldg(P.scale + h) * LOG2E;
      // the 32 reduced points of this block (per-warp copy: no block-level sync needed): (x, y, shift*sc2, post)
      __syncwarp();
      my_pts[lane] = next_pt;
      __syncwarp();
      next_pt = fetch(kb + 1);
      const float my_bias = own_vmin * sc2;
      mbar_wait(&empty_bar[s], (use & 1) ^ 1);
      unsigned char* stage = tiles_ptr + s * L::STAGE_BYTES;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = half * 4 + cc;
        float hi[4], lo[4], hid[4], lod[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 q = my_pts[c * 4 + e];
          const float d2 = dist2_fast<GEO>(me, q.x, q.y, period);
          float p = fast_exp2(fmaf(-d2, sc2, (MODE == DENSE_DVALUES) ? q.z : my_bias));
          if (MODE == DENSE_DVALUES) p *= q.w;
          lsum += p;
          hi[e] = operand_hi(p, prec);
          lo[e] = p - hi[e];
          if (MODE == DENSE_DSCALE) {
            const float pd = (d2 < INFINITY) ? p * d2 : 0.f;
            msum += pd;
            hid[e] = operand_hi(pd, prec);
            lod[e] = pd - hid[e];
          }
        }
        const uint32_t off = a_chunk_offset(r, c);
        *reinterpret_cast<float4*>(stage + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        if (split) *reinterpret_cast<float4*>(stage + L::A_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        if (MODE == DENSE_DSCALE) {
          *reinterpret_cast<float4*>(stage + 2 * L::A_BYTES + off) = make_float4(hid[0], hid[1], hid[2], hid[3]);
          if (split) *reinterpret_cast<float4*>(stage + 3 * L::A_BYTES + off) = make_float4(lod[0], lod[1], lod[2], lod[3]);
        }
      }
      fence_async_shared();      // this thread's tile writes are visible to the async proxy (the tensor pipe) ...
      __syncwarp();              // ... for every lane of the warp ...
      if (lane == 0) mbar_arrive(&full_bar[s]);   // ... before the warp's single arrival
    }

    // ------------------------------- epilogue -------------------------------
    half_sum[0][half][r] = lsum;
    half_sum[1][half][r] = msum;
    asm volatile("bar.sync 1, %0;" ::"n"(DENSE_GEN_THREADS));  // generator warps only
    lsum = half_sum[0][0][r] + half_sum[0][1][r];
    msum = half_sum[1][0][r] + half_sum[1][1][r];
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);   // TMEM lanes 32 (warp % 4) .. +31 = tile rows
    // the two threads of a row split the accumulator columns
    constexpr int EPI_COLS = NV / 2;
    const int c_lo = half * EPI_COLS;
    if (MODE == DENSE_FWD) {
      const float inv_l = 1.f / lsum;
      if (own_ok && bid.y == 0 && half == 0) P.rowsum_out[((int64_t)sample * P.H + h_fixed) * P.N + own] = lsum;
      for (int c0 = c_lo; c0 < c_lo + EPI_COLS; c0 += 32) {
        float v[32];
        tmem_ld32(lane_addr + c0, v);  // warp-collective: every lane takes part
        if (own_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const int n = n0 + c0 + e;
            if (n < P.width) {
              const int b = P.mesh_batched ? sample : n / P.D;
              const int d = P.mesh_batched ? n : n - b * P.D;
              float* dst = P.out + ((int64_t)b * P.N + own) * P.ld_out + P.col_off + (int64_t)h_fixed * P.D + d;
              *reinterpret_cast<float4*>(dst) = make_float4(v[e] * inv_l, v[e + 1] * inv_l, v[e + 2] * inv_l, v[e + 3] * inv_l);
            }
          }
        }
      }
    } else if (MODE == DENSE_DSCALE) {
      // -sum_e dO_e (W_e - (m/l) O_e) / l for this row, this block of value columns
      const float l = __ldg(P.rowsum + ((int64_t)sample * P.H + h_fixed) * P.N + (own_ok ? own : 0));
      const float ratio = msum / lsum;  // m/l (lsum == l up to rounding)
      float dot = 0.f;
      for (int c0 = c_lo; c0 < c_lo + EPI_COLS; c0 += 32) {
        float o[32], w[32];
        tmem_ld32(lane_addr + c0, o);
        tmem_ld32(lane_addr + NV + c0, w);
        if (own_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const int n = n0 + c0 + e;
            if (n < P.width) {
              const int b = P.mesh_batched ? sample : n / P.D;
              const int d = P.mesh_batched ? n : n - b * P.D;
              const float4 g = __ldg(reinterpret_cast<const float4*>(P.d_out + ((int64_t)b * P.N + own) * P.ld_out + P.col_off +
                                                                    (int64_t)h_fixed * P.D + d));
              dot = fmaf(g.x, w[e] - ratio * o[e], dot);
              dot = fmaf(g.y, w[e + 1] - ratio * o[e + 1], dot);
              dot = fmaf(g.z, w[e + 2] - ratio * o[e + 2], dot);
              dot = fmaf(g.w, w[e + 3] - ratio * o[e + 3], dot);
            }
          }
        }
      }
      float v = own_ok ? -dot / l : 0.f;
      v = warp_sum(v);
      if (lane == 0) red[warp] = v;
      asm volatile("bar.sync 1, %0;" ::"n"(DENSE_GEN_THREADS));  // generator warps only
      if (tid == 0) {
        float total = 0.f;
#pragma unroll
        for (int w = 0; w < DENSE_GEN_WARPS; ++w) total += red[w];
        atomicAdd(P.d_scale + h_fixed, total);
      }
    } else {
      // dU[b, j, :] (+ concat pass-through)
      for (int c0 = c_lo; c0 < c_lo + EPI_COLS; c0 += 32) {
        float v[32];
        tmem_ld32(lane_addr + c0, v);
        if (own_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const int n = n0 + c0 + e;
            if (n < P.width) {
              const int b = P.mesh_batched ? sample : n / P.D;
              const int d = P.mesh_batched ? n : n - b * P.D;
              float4 acc = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
              if (P.add_concat && h_dv == 0) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(P.d_out + ((int64_t)b * P.N + own) * P.ld_out + d));
                acc.x += g.x;
                acc.y += g.y;
                acc.z += g.z;
                acc.w += g.w;
              }
              float4* dst = reinterpret_cast<float4*>(P.d_values + ((int64_t)b * P.M + own) * P.D + d);
              if (P.split_heads)
                atomicAdd(dst, acc);  // H partial sums onto zeros: the result does not depend on their order for H = 2
              else
                *dst = acc;
            }
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp < DENSE_MMA_WARP) {
    // =========================== stagers: operand B ===========================
    // 128-bit loads along the value columns (coalesced), hi/lo split, 128-bit stores into the MN-major tile: a
    // quarter warp writes the eight 16-byte pieces of one 128-byte atom row, so the stores are conflict-free.
    const int t = tid - DENSE_GEN_THREADS;
    constexpr int N4 = NV / 4;                                               // 16-byte pieces per K-row
    constexpr int T_PER_ROW = N4 < DENSE_STAGE_THREADS ? N4 : DENSE_STAGE_THREADS;
    constexpr int ROWS_PER_PASS = DENSE_STAGE_THREADS / T_PER_ROW;
    constexpr int PASSES = DENSE_KB / ROWS_PER_PASS;                         // 16 (NV=256), 8 (128), 4 (64)
    static_assert(N4 <= DENSE_STAGE_THREADS, "NV must be <= 512");
    const int n4 = t % T_PER_ROW;
    const int krow0 = t / T_PER_ROW;
    const int n = n0 + n4 * 4;
    const bool n_ok = n < P.width;
    const int b = P.mesh_batched ? sample : (n_ok ? n / P.D : 0);
    const int d = P.mesh_batched ? n : n - b * P.D;
    const float* col_base = P.b_src + P.b_off + (int64_t)b * P.b_bstride + (n_ok ? d : 0);
    // The value block of K block kb+1 is requested (into a second register set) BEFORE block kb is split and stored: an L2 / HBM
    // round trip (the 1 MB value block of a sample streams through once per CTA) takes longer than one K block of MMAs, and
    // with a single register set every K block paid it in full -- the kernel ran at the load latency, not at the tensor pipe
    // (ncu, elasticity shape: the top stall was the first use of the loaded registers; 3xTF32 and single-product modes took
    // the same time).
    auto load_block = [&](int kb, float4 (&v)[PASSES]) {
      const int h = (MODE == DENSE_DVALUES) ? h_dv + kb / kb_per_head : 0;
      const int k0 = (kb % kb_per_head) * DENSE_KB;
#pragma unroll
      for (int it = 0; it < PASSES; ++it) {
        const int k = k0 + krow0 + it * ROWS_PER_PASS;
        v[it] = (kb < n_kb && n_ok && k < P.n_red)
                    ? __ldg(reinterpret_cast<const float4*>(col_base + (int64_t)k * P.b_kstride + (int64_t)h * P.b_hstride))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_block = [&](int kb, const float4 (&v)[PASSES]) {
      const int s = kb % DENSE_STAGES;
      const uint32_t use = kb / DENSE_STAGES;
      mbar_wait(&empty_bar[s], (use & 1) ^ 1);
      unsigned char* stage = tiles_ptr + s * L::STAGE_BYTES + L::A_TILES * L::A_BYTES;
#pragma unroll
      for (int it = 0; it < PASSES; ++it) {
        const float4 hi = make_float4(operand_hi(v[it].x, prec), operand_hi(v[it].y, prec), operand_hi(v[it].z, prec), operand_hi(v[it].w, prec));
        const uint32_t off = b_chunk_offset(krow0 + it * ROWS_PER_PASS, n4);
        *reinterpret_cast<float4*>(stage + off) = hi;
        if (split) *reinterpret_cast<float4*>(stage + L::B_BYTES + off) = make_float4(v[it].x - hi.x, v[it].y - hi.y, v[it].z - hi.z, v[it].w - hi.w);
      }
      fence_async_shared();      // this thread's tile writes are visible to the async proxy (the tensor pipe) ...
      __syncwarp();              // ... for every lane of the warp ...
      if (lane == 0) mbar_arrive(&full_bar[s]);   // ... before the warp's single arrival
    };
    float4 va[PASSES], vb[PASSES];
    load_block(0, va);
    for (int kb = 0; kb < n_kb; kb += 2) {
      load_block(kb + 1, vb);
      store_block(kb, va);
      if (kb + 1 < n_kb) {
        load_block(kb + 2, va);
        store_block(kb + 1, vb);
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    constexpr uint32_t IDESC = umma_idesc_tf32(NV);
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % DENSE_STAGES;
      const uint32_t use = kb / DENSE_STAGES;
      mbar_wait(&full_bar[s], use & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_base = tiles + s * L::STAGE_BYTES;
        const uint32_t b_base = a_base + L::A_TILES * L::A_BYTES;
#pragma unroll
        for (int kg = 0; kg < DENSE_KB / 8; ++kg) {
          const uint32_t acc = (kb > 0 || kg > 0) ? 1u : 0u;
          const uint64_t a_hi = umma_desc(a_base + kg * 32, 16, 1024);
          const uint64_t a_lo = umma_desc(a_base + L::A_BYTES + kg * 32, 16, 1024);
          const uint64_t b_hi = umma_desc_b(b_base + kg * B_KSTEP);
          const uint64_t b_lo = umma_desc_b(b_base + L::B_BYTES + kg * B_KSTEP);
          umma_tf32(tmem_base, a_hi, b_hi, IDESC, acc);
          if (split) {
            umma_tf32(tmem_base, a_lo, b_hi, IDESC, 1u);
            umma_tf32(tmem_base, a_hi, b_lo, IDESC, 1u);
          }
          if (MODE == DENSE_DSCALE) {
            const uint64_t ad_hi = umma_desc(a_base + 2 * L::A_BYTES + kg * 32, 16, 1024);
            const uint64_t ad_lo = umma_desc(a_base + 3 * L::A_BYTES + kg * 32, 16, 1024);
            umma_tf32(tmem_base + NV, ad_hi, b_hi, IDESC, acc);
            if (split) {
              umma_tf32(tmem_base + NV, ad_lo, b_hi, IDESC, 1u);
              umma_tf32(tmem_base + NV, ad_hi, b_lo, IDESC, 1u);
            }
          }
        }
        umma_commit(&empty_bar[s]);  // the stage may be overwritten once these MMAs have read it
        if (kb == n_kb - 1) umma_commit(&done_bar);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == DENSE_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int GEO, int MODE, int NV>
__global__ void __launch_bounds__(DENSE_THREADS, 1) dense_attention_kernel(const DenseParams P) {
  dense_attention_body<GEO, MODE, NV>(P, blockIdx);
}

// Backward of a small global stage in ONE launch: the scale-gradient CTAs and the value-gradient CTAs are independent
// (both only read dO), and each set alone covers a fraction of the chip -- side by side they overlap instead of queueing.
// Blocks [0, n_scale) run the DSCALE body on grid gs, the rest the DVALUES body on grid gv (both linearised x-fastest).
template <int GEO, int NV>
__global__ void __launch_bounds__(DENSE_THREADS, 1) dense_bwd_pair_kernel(const DenseParams Ps, const DenseParams Pv, const int n_scale,
                                                                        const dim3 gs, const dim3 gv) {
  const bool scale_part = (int)blockIdx.x < n_scale;  // CTA-uniform
  const dim3 g = scale_part ? gs : gv;
  const unsigned b = scale_part ? blockIdx.x : blockIdx.x - n_scale;
  const dim3 bid(b % g.x, (b / g.x) % g.y, b / (g.x * g.y));
  if (scale_part)
    dense_attention_body<GEO, DENSE_DSCALE, NV>(Ps, bid);
  else
    dense_attention_body<GEO, DENSE_DVALUES, NV>(Pv, bid);
}

template <int NV>
struct DensePairSmem {
  static constexpr int TOTAL = DenseSmem<DENSE_DSCALE, NV>::TOTAL > DenseSmem<DENSE_DVALUES, NV>::TOTAL ? DenseSmem<DENSE_DSCALE, NV>::TOTAL
                                                                                                       : DenseSmem<DENSE_DVALUES, NV>::TOTAL;
};

}  // namespace pit
